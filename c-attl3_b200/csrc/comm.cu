// comm.cu -- gradient all-reduce for data-parallel training: NCCL over NVLink 5 / NVSwitch, one
// process per GPU.  The reference has no distributed training at all; this is the
// exchange step of the data-parallel batch loop (cattle/optimizer/SGDOptimizer.hpp).
//
// libnccl is opened at run time (dlopen) so that single-GPU users, and the CPU-side symbol checks,
// do not need it.  Rendezvous needs no MPI: rank 0 publishes the ncclUniqueId in a file that the
// other ranks of the same host poll (one host, 1/2/4/8 GPUs -- BASELINE.json).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <string>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct NcclApi {
	void* lib = nullptr;
	int (*GetUniqueId)(NcclUniqueId*) = nullptr;
	int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
	int (*CommDestroy)(NcclComm) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*CommSplit)(NcclComm, int, int, NcclComm*, void*) = nullptr;   // NCCL >= 2.18; optional
	int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;   // optional (peer-memory set-up)
	const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl() {
	static NcclApi api;
	static bool tried = false;
	if (!tried) {
		tried = true;
		const char* names[] = { getenv("CATTL3_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
		for (const char* n : names) {
			if (n && (api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
				break;
		}
		if (api.lib) {
			api.GetUniqueId = (int (*)(NcclUniqueId*)) dlsym(api.lib, "ncclGetUniqueId");
			api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int)) dlsym(api.lib, "ncclCommInitRank");
			api.CommDestroy = (int (*)(NcclComm)) dlsym(api.lib, "ncclCommDestroy");
			api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)) dlsym(api.lib, "ncclAllReduce");
			api.GroupStart = (int (*)()) dlsym(api.lib, "ncclGroupStart");
			api.GroupEnd = (int (*)()) dlsym(api.lib, "ncclGroupEnd");
			api.GetErrorString = (const char* (*)(int)) dlsym(api.lib, "ncclGetErrorString");
			api.CommSplit = (int (*)(NcclComm, int, int, NcclComm*, void*)) dlsym(api.lib, "ncclCommSplit");
			api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)) dlsym(api.lib, "ncclAllGather");
			if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce || !api.GroupStart ||
					!api.GroupEnd) {
				dlclose(api.lib);
				api.lib = nullptr;
			}
		}
	}
	return api.lib ? &api : nullptr;
}

int nccl_fail(int rc, const char* what) {
	NcclApi* n = nccl();
	cattl3::set_error("NCCL error %d (%s) in %s", rc, n && n->GetErrorString ? n->GetErrorString(rc) : "?", what);
	return CATTL3_ERR_CUDA;
}

#define CATTL3_NCCL(call, what) do { int rc__ = (call); if (rc__ != NCCL_SUCCESS) return nccl_fail(rc__, what); } while (0)

// ---- one-shot all-reduce of short vectors over NVLink peer memory ---------------------------------------------------
// The statistics of a synchronised BatchNorm layer are 2 C + 1 doubles per direction, dozens of times per step: far below
// the size where NCCL's ring / tree protocols pay, and each ncclAllReduce costs its launch plus a multi-hop handshake.
// Every rank owns a mailbox in its HBM that its peers map (cudaIpc; NVSwitch makes every peer one hop away): a rank
// WRITES its vector into its lane of every mailbox, raises a flag there, waits until its own mailbox has every lane's
// flag and adds the lanes in rank order -- the same order on every rank, so the results are bit-identical everywhere.
// One kernel of one CTA, no intermediate hops.  Two slots alternate: a rank can run at most one call ahead of its
// slowest peer (it needs that peer's flag of the current call to finish it), so the slot it writes next is never the
// one a peer still reads.  The call counter lives in the mailbox and is advanced by the kernel itself, so a captured
// step graph replays correctly.
constexpr int PEER_MAX_RANKS = 8, PEER_MAX_BYTES = 16384, PEER_SLOTS = 2;
struct PeerMailbox {
	unsigned long long epoch;                                           // calls completed by the owner
	unsigned long long pad[15];
	unsigned long long flags[PEER_SLOTS][PEER_MAX_RANKS * 16];          // [slot][lane * 16]: the call number the lane's data belong to (a line each)
	unsigned char data[PEER_SLOTS][PEER_MAX_RANKS][PEER_MAX_BYTES];
};
struct PeerTable { PeerMailbox* box[PEER_MAX_RANKS]; };

template<typename T>
__global__ void __launch_bounds__(256) peer_allreduce_kernel(T* __restrict__ buf, int count, int rank, int world, PeerTable peers) {
	PeerMailbox* mine = peers.box[rank];
	const unsigned long long call = mine->epoch + 1;
	const int slot = (int) (call & 1);
	for (int r = 0; r < world; ++r) {
		T* lane = reinterpret_cast<T*>(peers.box[r]->data[slot][rank]);
		for (int i = threadIdx.x; i < count; i += 256) lane[i] = buf[i];
	}
	__threadfence_system();
	__syncthreads();
	if ((int) threadIdx.x < world)
		*reinterpret_cast<volatile unsigned long long*>(&peers.box[threadIdx.x]->flags[slot][rank * 16]) = call;
	if ((int) threadIdx.x < world) {
		const volatile unsigned long long* flag = &mine->flags[slot][threadIdx.x * 16];
		while (*flag != call) { }
		__threadfence_system();
	}
	__syncthreads();
	for (int i = threadIdx.x; i < count; i += 256) {
		T sum = reinterpret_cast<const volatile T*>(mine->data[slot][0])[i];
		for (int r = 1; r < world; ++r) sum += reinterpret_cast<const volatile T*>(mine->data[slot][r])[i];
		buf[i] = sum;
	}
	__syncthreads();
	if (threadIdx.x == 0) mine->epoch = call;
}

} // namespace

struct cattl3_comm {
	cattl3_ctx* ctx = nullptr;
	NcclComm comm = nullptr;
	// a duplicate communicator for the side-stream exchanges (ncclCommSplit): operations on ONE communicator serialise
	// in issue order whatever their streams, which would chain the main stream's BatchNorm-statistics all-reduces
	// behind the gradient buckets travelling beside them; null = share `comm`
	NcclComm comm_async = nullptr;
	int world = 1, rank = 0;
	// the side stream of the _async all-reduces: forked from the context's stream by an event per call, joined by
	// cattl3_comm_wait (so the exchange of one layer's gradients overlaps the kernels of the layers behind it)
	cudaStream_t side = nullptr;
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	bool pending = false;
	// peer-memory mailboxes for short vectors (null: not set up, every exchange goes through NCCL)
	PeerMailbox* mailbox = nullptr;
	PeerTable peers = {};
	bool peer_ok = false;
};

// Maps every rank's mailbox into this process.  Any failure leaves peer_ok false (NCCL carries everything).
static void peer_setup(cattl3_comm* c) {
	NcclApi* n = nccl();
	if (!n || !n->AllGather || c->world > PEER_MAX_RANKS || getenv("CATTL3_NO_PEER_REDUCE"))
		return;
	// (every rank reaches the two collectives below whatever fails locally: env and world size are the same everywhere)
	bool ok = cudaMalloc((void**) &c->mailbox, sizeof(PeerMailbox)) == cudaSuccess;
	if (!ok) { cudaGetLastError(); c->mailbox = nullptr; }
	cudaIpcMemHandle_t mine;
	if (ok) {
		cudaMemset(c->mailbox, 0, sizeof(PeerMailbox));
		ok = cudaIpcGetMemHandle(&mine, c->mailbox) == cudaSuccess;
	}
	// every rank takes part in the gather whatever happened locally (a rank that failed sends a zero handle)
	unsigned char* dev = nullptr;
	const size_t hb = sizeof(cudaIpcMemHandle_t) + 8;
	unsigned char host[(sizeof(cudaIpcMemHandle_t) + 8) * PEER_MAX_RANKS] = {};
	unsigned char rec[sizeof(cudaIpcMemHandle_t) + 8] = {};
	if (ok) { memcpy(rec, &mine, sizeof(mine)); rec[sizeof(mine)] = 1; }
	if (cudaMalloc((void**) &dev, hb * (c->world + 1)) != cudaSuccess) { cudaGetLastError(); return; }
	cudaMemcpy(dev, rec, hb, cudaMemcpyHostToDevice);
	const int rc = n->AllGather(dev, dev + hb, hb, /* ncclChar */ 0, c->comm, c->ctx->stream);
	if (rc != NCCL_SUCCESS || cudaStreamSynchronize(c->ctx->stream) != cudaSuccess) { cudaGetLastError(); cudaFree(dev); return; }
	cudaMemcpy(host, dev + hb, hb * c->world, cudaMemcpyDeviceToHost);
	cudaFree(dev);
	for (int r = 0; r < c->world; ++r) ok = ok && host[r * hb + sizeof(cudaIpcMemHandle_t)] == 1;
	for (int r = 0; ok && r < c->world; ++r) {
		if (r == c->rank) { c->peers.box[r] = c->mailbox; continue; }
		cudaIpcMemHandle_t h;
		memcpy(&h, host + r * hb, sizeof(h));
		void* p = nullptr;
		if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
		c->peers.box[r] = (PeerMailbox*) p;
	}
	// all or nothing: a rank that could not map a peer keeps everybody on NCCL
	float* agree = nullptr;
	if (cudaMalloc((void**) &agree, sizeof(float)) != cudaSuccess) { cudaGetLastError(); return; }
	const float vote = ok ? 1.f : 0.f;
	float votes = 0.f;
	cudaMemcpy(agree, &vote, sizeof(float), cudaMemcpyHostToDevice);
	if (n->AllReduce(agree, agree, 1, NCCL_FLOAT32, NCCL_SUM, c->comm, c->ctx->stream) == NCCL_SUCCESS &&
			cudaStreamSynchronize(c->ctx->stream) == cudaSuccess)
		cudaMemcpy(&votes, agree, sizeof(float), cudaMemcpyDeviceToHost);
	else
		cudaGetLastError();
	cudaFree(agree);
	c->peer_ok = ok && votes == (float) c->world;
}

using namespace cattl3;

extern "C" {

int cattl3_comm_unique_id(void* id128) {
	CATTL3_REQUIRE(id128, "comm_unique_id: null buffer");
	NcclApi* n = nccl();
	if (!n) {
		set_error("libnccl.so.2 could not be loaded (set CATTL3_NCCL_LIB)");
		return CATTL3_ERR_UNSUPPORTED;
	}
	NcclUniqueId id;
	CATTL3_NCCL(n->GetUniqueId(&id), "ncclGetUniqueId");
	memcpy(id128, &id, sizeof(id));
	return CATTL3_OK;
}

int cattl3_comm_create(cattl3_comm** out, cattl3_ctx* ctx, int world_size, int rank, const void* id128) {
	CATTL3_REQUIRE(out, "comm_create: null out pointer");
	*out = nullptr;
	CATTL3_CHECK(check_ctx(ctx));
	CATTL3_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "comm_create: rank %d outside world of %d", rank, world_size);
	cattl3_comm* c = new cattl3_comm();
	c->ctx = ctx; c->world = world_size; c->rank = rank;
	if (world_size > 1) {
		NcclApi* n = nccl();
		if (!n) {
			delete c;
			set_error("libnccl.so.2 could not be loaded (set CATTL3_NCCL_LIB)");
			return CATTL3_ERR_UNSUPPORTED;
		}
		if (!id128) {
			delete c;
			set_error("comm_create: a unique id is required for world_size > 1");
			return CATTL3_ERR_INVALID;
		}
		NcclUniqueId id;
		memcpy(&id, id128, sizeof(id));
		int rc = n->CommInitRank(&c->comm, world_size, id, rank);
		if (rc != NCCL_SUCCESS) {
			delete c;
			return nccl_fail(rc, "ncclCommInitRank");
		}
		if (n->CommSplit && !getenv("CATTL3_COMM_SINGLE")) {
			if (n->CommSplit(c->comm, 0, rank, &c->comm_async, nullptr) != NCCL_SUCCESS)
				c->comm_async = nullptr;
		}
		peer_setup(c);
		int lo = 0, hi = 0;
		cudaDeviceGetStreamPriorityRange(&lo, &hi);   // the exchange is short and latency bound: highest priority
		if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi) != cudaSuccess ||
				cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
				cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
			cattl3_comm_destroy(c);
			set_error("comm_create: side stream / events");
			return CATTL3_ERR_CUDA;
		}
	}
	*out = c;
	return CATTL3_OK;
}

// WORLD_SIZE / RANK from the environment (the torchrun convention); the unique id travels through the
// file CATTL3_COMM_ID_FILE (default /tmp/cattl3_nccl_id.<MASTER_PORT or 0>), written by rank 0 with an
// atomic rename and removed by rank 0 when the communicator is destroyed.
int cattl3_comm_create_from_env(cattl3_comm** out, cattl3_ctx* ctx) {
	CATTL3_REQUIRE(out, "comm_create_from_env: null out pointer");
	const char* ws = getenv("WORLD_SIZE");
	const char* rk = getenv("RANK");
	const int world = ws ? atoi(ws) : 1, rank = rk ? atoi(rk) : 0;
	if (world <= 1)
		return cattl3_comm_create(out, ctx, 1, 0, nullptr);
	// The file is tied to the job: the ranks of one launch share their parent (the torchrun agent, or the test that
	// spawned them), so a file left behind by a crashed run on the same port is never read by the next one.
	std::string path;
	if (const char* p = getenv("CATTL3_COMM_ID_FILE")) {
		path = p;
	} else {
		const char* port = getenv("MASTER_PORT");
		const char* run = getenv("TORCHELASTIC_RUN_ID");
		static int sequence = 0;   // communicators are created in the same order on every rank
		path = std::string("/tmp/cattl3_nccl_id.") + (port ? port : "0") + "." + std::to_string((long long) getppid()) +
				(run ? std::string(".") + run : std::string()) + "." + std::to_string(sequence++);
		for (char& ch : path) if (ch == ' ' || ch == ':') ch = '_';
	}
	NcclUniqueId id;
	if (rank == 0) {
		CATTL3_CHECK(cattl3_comm_unique_id(&id));
		const std::string tmp = path + ".tmp";
		unlink(tmp.c_str());
		FILE* f = fopen(tmp.c_str(), "wbx");   // exclusive: never follows a planted link
		CATTL3_REQUIRE(f, "cannot write %s", tmp.c_str());
		fwrite(&id, sizeof(id), 1, f);
		fclose(f);
		CATTL3_REQUIRE(rename(tmp.c_str(), path.c_str()) == 0, "cannot publish %s", path.c_str());
	} else {
		bool got = false;
		for (int i = 0; i < 6000 && !got; ++i) {  // up to 60 s
			FILE* f = fopen(path.c_str(), "rb");
			if (f) {
				got = fread(&id, sizeof(id), 1, f) == 1;
				fclose(f);
			}
			if (!got) usleep(10000);
		}
		CATTL3_REQUIRE(got, "timed out waiting for the NCCL id file %s", path.c_str());
	}
	int rc = cattl3_comm_create(out, ctx, world, rank, &id);
	if (rank == 0 && rc == CATTL3_OK)
		unlink(path.c_str());  // every rank has joined once ncclCommInitRank returns
	return rc;
}

int cattl3_comm_destroy(cattl3_comm* c) {
	if (!c) return CATTL3_OK;
	if (c->comm) {
		cudaSetDevice(c->ctx->device);
		if (c->side) cudaStreamSynchronize(c->side);
		cudaStreamSynchronize(c->ctx->stream);
		if (NcclApi* n = nccl()) {
			if (c->comm_async) n->CommDestroy(c->comm_async);
			n->CommDestroy(c->comm);
		}
	}
	if (c->peer_ok) {
		for (int r = 0; r < c->world; ++r)
			if (r != c->rank && c->peers.box[r]) cudaIpcCloseMemHandle(c->peers.box[r]);
	}
	if (c->mailbox) cudaFree(c->mailbox);
	if (c->ev_fork) cudaEventDestroy(c->ev_fork);
	if (c->ev_join) cudaEventDestroy(c->ev_join);
	if (c->side) cudaStreamDestroy(c->side);
	delete c;
	return CATTL3_OK;
}

int cattl3_comm_world_size(const cattl3_comm* c) { return c ? c->world : 1; }
int cattl3_comm_rank(const cattl3_comm* c) { return c ? c->rank : 0; }

int cattl3_comm_group_start(cattl3_comm* c) {
	CATTL3_REQUIRE(c, "null communicator");
	if (c->world > 1) CATTL3_NCCL(nccl()->GroupStart(), "ncclGroupStart");
	return CATTL3_OK;
}
int cattl3_comm_group_end(cattl3_comm* c) {
	CATTL3_REQUIRE(c, "null communicator");
	if (c->world > 1) CATTL3_NCCL(nccl()->GroupEnd(), "ncclGroupEnd");
	return CATTL3_OK;
}

static int allreduce(cattl3_comm* c, void* buf, int64_t count, int dtype) {
	CATTL3_REQUIRE(c && buf && count > 0, "comm_allreduce: bad arguments");
	CATTL3_CHECK(check_ctx(c->ctx));
	if (c->world == 1)
		return CATTL3_OK;
	const size_t bytes = (size_t) count * (dtype == NCCL_FLOAT64 ? 8 : 4);
	if (c->peer_ok && bytes <= (size_t) PEER_MAX_BYTES) {
		// short vectors (BatchNorm statistics, scalars): one-shot exchange through the ranks' mailboxes
		if (dtype == NCCL_FLOAT64)
			peer_allreduce_kernel<double><<<1, 256, 0, c->ctx->stream>>>((double*) buf, (int) count, c->rank, c->world, c->peers);
		else
			peer_allreduce_kernel<float><<<1, 256, 0, c->ctx->stream>>>((float*) buf, (int) count, c->rank, c->world, c->peers);
		CATTL3_LAUNCHED(c->ctx);
		return CATTL3_OK;
	}
	CATTL3_NCCL(nccl()->AllReduce(buf, buf, (size_t) count, dtype, NCCL_SUM, c->comm, c->ctx->stream), "ncclAllReduce");
	return CATTL3_OK;
}
// The same exchange on the communicator's side stream: it starts when everything enqueued on the context's stream so far
// has finished (the weight gradients it sums) and runs beside whatever is enqueued next (the input gradient, the
// layers further back); cattl3_comm_wait makes the context's stream wait for every exchange started this way.
static int allreduce_async(cattl3_comm* c, void* buf, int64_t count, int dtype) {
	CATTL3_REQUIRE(c && buf && count > 0, "comm_allreduce_async: bad arguments");
	CATTL3_CHECK(check_ctx(c->ctx));
	if (c->world == 1)
		return CATTL3_OK;
	CATTL3_REQUIRE(!c->ctx->capturing, "comm_allreduce_async: not inside a step graph");
	CATTL3_CUDA(cudaEventRecord(c->ev_fork, c->ctx->stream));
	CATTL3_CUDA(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
	CATTL3_NCCL(nccl()->AllReduce(buf, buf, (size_t) count, dtype, NCCL_SUM, c->comm_async ? c->comm_async : c->comm, c->side),
			"ncclAllReduce");
	c->pending = true;
	return CATTL3_OK;
}
int cattl3_comm_allreduce_sum_async_f32(cattl3_comm* c, float* buf, int64_t count) { return allreduce_async(c, buf, count, NCCL_FLOAT32); }
int cattl3_comm_allreduce_sum_async_f64(cattl3_comm* c, double* buf, int64_t count) { return allreduce_async(c, buf, count, NCCL_FLOAT64); }
int cattl3_comm_wait(cattl3_comm* c) {
	CATTL3_REQUIRE(c, "null communicator");
	if (c->world == 1 || !c->pending)
		return CATTL3_OK;
	CATTL3_CUDA(cudaEventRecord(c->ev_join, c->side));
	CATTL3_CUDA(cudaStreamWaitEvent(c->ctx->stream, c->ev_join, 0));
	c->pending = false;
	return CATTL3_OK;
}
int cattl3_comm_allreduce_sum_f32(cattl3_comm* c, float* buf, int64_t count) { return allreduce(c, buf, count, NCCL_FLOAT32); }
int cattl3_comm_allreduce_sum_f64(cattl3_comm* c, double* buf, int64_t count) { return allreduce(c, buf, count, NCCL_FLOAT64); }

}
