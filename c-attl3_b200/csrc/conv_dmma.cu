// conv_dmma.cu -- the double instantiation of the kernel layers on the FP64 TENSOR pipe: warp-level
// mma.sync.m8n8k4.f64 (DMMA).  Measured on B200 (scripts/dmma_peak.cu): 37.0 TFLOP/s, against 33.7 for plain DFMA
// (scripts/dfma_peak.cu) -- but what matters more is what each instruction carries: one DMMA is 256 FMAs for two
// 8-byte fragment loads, where the FMA kernels (conv_dfma.cu) issue 64 DFMAs per eight 16-byte shared-memory loads
// and stall on exactly those loads (profiles/README.md, r1e: FP64 pipe 51 % active, short-scoreboard stalls).
// tcgen05 has no fp64 kind, so this is the tensor-core path there is for double.
//
// Two families.  Layers whose grid fills the device run the PRODUCER-WARP kernels further down (dmma2_*: a cp.async ring behind
// mbarriers, eight warps that only read fragments and issue DMMAs; config 2: 0.88 / 0.78 / 0.88 of the DMMA rate).  Smaller
// layers keep the first family, directly below: same implicit GEMM and tiling as conv_dfma.cu -- 128 x 64 (weight gradient:
// also 128 x 128) block tiles, 256 threads = 8 warps as 2 (rows) x 4 (columns), a warp owns 64 x 16 (64 x 32) as 8 x 2 (8 x 4)
// m8n8k4 tiles, accumulators in registers; tiles staged global -> shared with cp.async, two buffers, one barrier per k-block
// of 8, two CTAs per SM; no im2col buffer.
// Fragments (PTX ISA, mma.m8n8k4 .f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4],
// C[row = lane / 4][col = 2 * (lane % 4) + {0, 1}].  Shared-memory rows are k (the reduction index) with a pitch of
// tile + 4 doubles: conflict free for the fragment reads (the weight gradient's transposing 8-byte writes are 2-way
// conflicts -- 301 M of them at config 2, one of the reasons for the second family's [row][8 m] layout).
#include <stdlib.h>

#include "activations.cuh"

namespace cattl3 {

namespace {

constexpr int DM_BM = 128, DM_BK = 8, DM_THREADS = 256, DM_PAD = 4;

__device__ __forceinline__ void cp_async_f64(double* dst, const double* src, bool pred) {
	const uint32_t d = (uint32_t) __cvta_generic_to_shared(dst);
	const int bytes = pred ? 8 : 0;
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_f64() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_f64() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
			: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// One k-block (8 reduction steps) of the warp's 64 x (8 * NI) tile: As / Bs point at [k][row] / [k][col] of the
// current buffer, row0 / col0 are the warp's offsets in the block tile.
template<int NI, int PA, int PB>
__device__ __forceinline__ void dmma_block(const double* __restrict__ As, const double* __restrict__ Bs, int row0, int col0,
		int lane, double (&acc)[8][NI][2]) {
	const int kq = lane & 3, g = lane >> 2;
	#pragma unroll
	for (int k4 = 0; k4 < DM_BK; k4 += 4) {
		double a[8], b[NI];
		#pragma unroll
		for (int mi = 0; mi < 8; ++mi) a[mi] = As[(k4 + kq) * PA + row0 + 8 * mi + g];
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(k4 + kq) * PB + col0 + 8 * ni + g];
		#pragma unroll
		for (int mi = 0; mi < 8; ++mi)
			#pragma unroll
			for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
	}
}

// Epilogue of the gather GEMMs in three straight-line passes (bias, activation with the switch over the kind OUTSIDE the
// element loop, stores): with the switch inside, the elements of a thread were as many jumps through hundreds of KB of
// code -- a fifth of the forward kernel's stall samples at config 2 were instruction fetches of this part
// (profiles/README.md, r2k).  The lane holds rows 8 mi + lane / 4, columns 8 ni + 2 (lane % 4) + {0, 1} of the warp tile
// whose first row / column are mw / jw.
template<int NI>
__device__ __forceinline__ void dmma_epilogue(double (&acc)[8][NI][2], const GatherGeom& gg, long long M, int J, long long mw, int jw,
		int lane, const double* __restrict__ bias, int bias_mode, double* __restrict__ out, int act_kind, double act_param,
		double* __restrict__ act_out) {
	const long long P = (long long) gg.OH * gg.OW;
	const int g = lane >> 2, kq = lane & 3;
	if (bias_mode != 0) {
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) {
			#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int j = jw + 8 * ni + 2 * kq + e;
				if (j >= J) continue;
				if (bias_mode == 1) {
					const double bj = __ldg(bias + j);
					#pragma unroll
					for (int mi = 0; mi < 8; ++mi) acc[mi][ni][e] += bj;
				} else {
					#pragma unroll
					for (int mi = 0; mi < 8; ++mi) {
						const long long m = mw + 8 * mi + g;
						if (m < M) acc[mi][ni][e] += __ldg(bias + m / gg.N + P * j);
					}
				}
			}
		}
	}
	auto store = [&](double* __restrict__ dst) {
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) {
			#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int j = jw + 8 * ni + 2 * kq + e;
				if (j >= J) continue;
				#pragma unroll
				for (int mi = 0; mi < 8; ++mi) {
					const long long m = mw + 8 * mi + g;
					if (m < M) dst[m + M * j] = acc[mi][ni][e];
				}
			}
		}
	};
	if (out) store(out);
	if (act_out) {
		act_fwd_rt_n<double, 8 * NI * 2>(act_kind, &acc[0][0][0], act_param);
		store(act_out);
	}
}

// Gather GEMM: out[m + M*j] = bias + sum_{tap, r} src(m, tap, r) * w(tap, r, j) (+ fused activation).
template<int NI>
__global__ void __launch_bounds__(DM_THREADS, NI == 2 ? 2 : 1) dmma_gather_gemm_kernel(GatherGeom gg,
		const double* __restrict__ src, const double* __restrict__ w, const double* __restrict__ bias, int bias_mode,
		double* __restrict__ out, int act_kind, double act_param, double* __restrict__ act_out) {
	constexpr int BN = 32 * NI;
	constexpr int PA = DM_BM + DM_PAD, PB = BN + DM_PAD;
	constexpr int B_ITERS = DM_BK * BN / DM_THREADS;   // 4 (BN = 128) or 2 (BN = 64)
	constexpr int B_KSTEP = DM_THREADS / BN;           // 2 or 4
	__shared__ __align__(16) double As[2][DM_BK * PA];
	__shared__ __align__(16) double Bs[2][DM_BK * PB];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int row0 = (warp & 1) * 64, col0 = (warp >> 1) * (8 * NI);
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long m0 = (long long) blockIdx.x * DM_BM;
	const int j0 = blockIdx.y * BN;
	const int R = gg.SC, J = gg.J, T = gg.RH * gg.RW;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;
	const bool unit_den = gg.denh == 1 && gg.denw == 1;

	// loader: this thread always fetches row a_ml of the A tile and column b_j of the B tile
	const int a_ml = tid & (DM_BM - 1), a_k0 = tid >> 7;
	const long long am = m0 + a_ml;
	const bool m_ok = am < M;
	const int an = (int) (am % gg.N);
	const long long apix = am / gg.N;
	const int aoh = (int) (apix % gg.OH), aow = (int) (apix / gg.OH);
	const int b_j = tid & (BN - 1), b_k0 = tid / BN;
	const bool j_ok = j0 + b_j < J;
	const long long wj = (long long) (j0 + b_j) * gg.w_sj;

	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

	// taps outer, channel blocks of 8 inner (a partial last block is zero filled); coordinates once per tap
	auto tap_src = [&](int tap) -> const double* {
		const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
		const int th = aoh * gg.ah + rh * gg.bh + gg.ch;
		const int tw = aow * gg.aw + rw * gg.bw + gg.cw;
		if (!m_ok || th < 0 || tw < 0) return nullptr;
		int ih = th, iw = tw;
		if (!unit_den) {
			if (th % gg.denh != 0 || tw % gg.denw != 0) return nullptr;
			ih = th / gg.denh; iw = tw / gg.denw;
		}
		if (ih >= gg.SH || iw >= gg.SW) return nullptr;
		return src + an + (long long) gg.N * (ih + (long long) gg.SH * iw);
	};
	const int rblocks = (R + DM_BK - 1) / DM_BK;
	const int ksteps = T * rblocks;
	int f_tap = 0, f_r0 = 0;
	const double* f_src = tap_src(0);
	const double* f_w = w + wj;
	auto fetch = [&](int buf) {
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int r = f_r0 + a_k0 + 2 * i;
			const bool ok = f_src && r < R;
			cp_async_f64(&As[buf][(a_k0 + 2 * i) * PA + a_ml], ok ? f_src + r * plane : src, ok);
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int r = f_r0 + b_k0 + B_KSTEP * i;
			const bool ok = j_ok && r < R;
			cp_async_f64(&Bs[buf][(b_k0 + B_KSTEP * i) * PB + b_j], ok ? f_w + r * gg.w_sr : w, ok);
		}
		cp_async_commit_f64();
		f_r0 += DM_BK;
		if (f_r0 >= R) {
			f_r0 = 0;
			if (++f_tap < T) {
				f_src = tap_src(f_tap);
				f_w = w + wj + f_tap * gg.w_stap;
			}
		}
	};
	fetch(0);
	cp_async_wait_f64();
	__syncthreads();
	for (int ks = 0; ks < ksteps; ++ks) {
		const int buf = ks & 1;
		if (ks + 1 < ksteps) fetch(buf ^ 1);
		dmma_block<NI, PA, PB>(As[buf], Bs[buf], row0, col0, lane, acc);
		cp_async_wait_f64();
		__syncthreads();
	}

	dmma_epilogue<NI>(acc, gg, M, J, m0 + row0, j0 + col0, lane, bias, bias_mode, out, act_kind, act_param, act_out);
}

// Weight gradient: dw(tap, r, j) += sum_m src(m, tap, r) * plain[m + M*j]; rows of the output tile are the flattened
// (tap, r), the reduction runs over m in blocks of 8 (both operands are m-contiguous in HBM: 64-byte runs, transposed
// on the way into shared memory); split over m across the grid, per-split partials, deterministic reduce.
template<int NI>
__global__ void __launch_bounds__(DM_THREADS, NI == 2 ? 2 : 1) dmma_wgrad_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ plain, double* __restrict__ partial, long long m_per_split, long long dw_elems) {
	constexpr int BN = 32 * NI;
	constexpr int PA = DM_BM + DM_PAD, PB = BN + DM_PAD;
	constexpr int B_ITERS = BN / 32;
	__shared__ __align__(16) double As[2][DM_BK * PA];
	__shared__ __align__(16) double Bs[2][DM_BK * PB];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int row0 = (warp & 1) * 64, col0 = (warp >> 1) * (8 * NI);
	const int R = gg.SC, J = gg.J;
	const int Ktot = gg.RH * gg.RW * R;
	const int k0 = blockIdx.x * DM_BM, j0 = blockIdx.y * BN;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.z * m_per_split;
	const long long me = ms + m_per_split < M ? ms + m_per_split : M;
	const long long plane = (long long) gg.N * gg.SH * gg.SW;

	// loader: 8 consecutive m per (tap, r) row / filter column; this thread owns m offset l_mm and rows l_r0 + 32 i
	const int l_mm = tid & 7, l_r0 = tid >> 3;
	int row_dh[4], row_dw[4], row_r[4];
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const int k = k0 + l_r0 + 32 * i;
		if (k < Ktot) {
			const int tap = k / R, r = k - tap * R;
			const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
			row_dh[i] = rh * gg.bh + gg.ch; row_dw[i] = rw * gg.bw + gg.cw; row_r[i] = r;
		} else {
			row_dh[i] = 0; row_dw[i] = 0; row_r[i] = -1;
		}
	}

	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

	auto fetch = [&](long long mc, int buf) {
		const long long m = mc + l_mm;
		const bool ok = m < me;
		const unsigned mu = (unsigned) m, pixu = mu / (unsigned) gg.N;   // M < 2^31
		const int n = (int) (mu - pixu * (unsigned) gg.N);
		const int ow = (int) (pixu / (unsigned) gg.OH), oh = (int) (pixu - (unsigned) ow * (unsigned) gg.OH);
		const int bh = oh * gg.ah, bw = ow * gg.aw;
		#pragma unroll
		for (int i = 0; i < 4; ++i) {
			const int th = bh + row_dh[i], tw = bw + row_dw[i];
			const bool aok = ok && row_r[i] >= 0 && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW;   // denh = denw = 1
			cp_async_f64(&As[buf][l_mm * PA + l_r0 + 32 * i],
					aok ? src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + row_r[i] * plane : src, aok);
		}
		#pragma unroll
		for (int i = 0; i < B_ITERS; ++i) {
			const int j = j0 + l_r0 + 32 * i;
			const bool bok = ok && j < J;
			cp_async_f64(&Bs[buf][l_mm * PB + l_r0 + 32 * i], bok ? plain + m + M * j : plain, bok);
		}
		cp_async_commit_f64();
	};

	const long long steps = me > ms ? (me - ms + DM_BK - 1) / DM_BK : 0;
	if (steps > 0) {
		fetch(ms, 0);
		cp_async_wait_f64();
	}
	__syncthreads();
	for (long long st = 0; st < steps; ++st) {
		const int buf = (int) (st & 1);
		if (st + 1 < steps) fetch(ms + (st + 1) * DM_BK, buf ^ 1);
		dmma_block<NI, PA, PB>(As[buf], Bs[buf], row0, col0, lane, acc);
		cp_async_wait_f64();
		__syncthreads();
	}

	double* dst = partial + (long long) blockIdx.z * dw_elems;
	const int g = lane >> 2, kq = lane & 3;
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi) {
		const int k = k0 + row0 + 8 * mi + g;
		if (k >= Ktot) continue;
		const int tap = k / R, r = k - tap * R;
		const long long base = tap * gg.w_stap + r * gg.w_sr;
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) {
			#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int j = j0 + col0 + 8 * ni + 2 * kq + e;
				if (j < J) dst[base + j * gg.w_sj] = acc[mi][ni][e];
			}
		}
	}
}

__global__ void __launch_bounds__(256) dmma_wgrad_reduce_kernel(const double* __restrict__ partial, int splits, long long elems,
		double* __restrict__ dw) {
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < elems; i += (long long) gridDim.x * 256) {
		double s = 0;
		for (int z = 0; z < splits; ++z) s += partial[(long long) z * elems + i];
		dw[i] += s;
	}
}


// ---- version 2: producer warps + a ring of stages ---------------------------------------------------------------------
// What held the kernels above at 0.52 - 0.74 of the DMMA rate (profiles/README.md, r2a): ptxas spaces a warp's DMMAs 16
// clocks apart, which IS the rate of the pipe (one m8n8k4 per 16 clocks per scheduler) -- one warp per scheduler would
// fill it if it issued nothing else.  But every warp also ran the loader (pixel decode, bounds, 64-bit addresses: ~240
// of its ~370 instructions per k-block), all warps at the same time right behind the barrier, with one k-block (~1 us)
// of lookahead for loads that come from HBM.  Here eight warps only read fragments and issue DMMAs; four producer warps
// run the loader four k-blocks ahead (cp.async, completion signalled through mbarriers with
// cp.async.mbarrier.arrive.noinc), and a stage is handed back by one arrival per consumer warp.  No CTA-wide barrier
// in the main loop.  Gather GEMM: the reduction runs channel block outer, taps inner, so the nine shifted reads of a
// channel block follow each other (the input gradient of config 2 read 7.8 GB of HBM for 2.1 GB with taps outer), and
// the batch-contiguous A rows travel as 16-byte copies.  Weight gradient: tiles stay [row][8 m] in shared memory the way
// they lie in HBM (16-byte copies, no 2-way conflicts of transposing 8-byte writes), with the 16-byte chunks of a row
// XOR-swizzled so that the fragment reads are conflict free; 64 x 256 tiles where 128 rows would not divide
// taps x channels (config 2: 576 = 9 x 64 against 5 x 128 = 640).
constexpr int D2_BK = 8, D2_STAGES = 4, D2_CONSUMERS = 256, D2_PRODUCERS = 128, D2_THREADS = D2_CONSUMERS + D2_PRODUCERS;

__device__ __forceinline__ uint32_t d2_smem(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void d2_mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(d2_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void d2_mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(d2_smem(bar)) : "memory");
}
__device__ __forceinline__ void d2_mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"D2_WAIT_LOOP:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra D2_WAIT_DONE;\n\t"
		"bra D2_WAIT_LOOP;\n\t"
		"D2_WAIT_DONE:\n\t}"
		:: "r"(d2_smem(bar)), "r"(parity) : "memory");
}
// the executing thread's earlier cp.async copies, once complete, count as ONE arrival (already part of the init count)
__device__ __forceinline__ void d2_cp_async_arrive(uint64_t* bar) {
	asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(d2_smem(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_f64x2(double* dst, const double* src, bool pred) {
	const uint32_t d = d2_smem(dst);
	const int bytes = pred ? 16 : 0;
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all_f64() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Twelve warps = three per scheduler = 168 registers each; the accumulators (128) and two sets of fragments want more, the
// loader needs far less: the producer warp group hands registers over (per scheduler 80 + 2 x 208 <= 3 x 168).
constexpr int D2_PRODUCER_REGS = 80, D2_CONSUMER_REGS = 208;
__device__ __forceinline__ void d2_regs_producer() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(D2_PRODUCER_REGS)); }
__device__ __forceinline__ void d2_regs_consumer() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(D2_CONSUMER_REGS)); }

// ring bookkeeping shared by both kernels: stage s of k-block kb, and the parity of its barriers' current phase
__device__ __forceinline__ void d2_ring_init(uint64_t* full, uint64_t* empty, int tid) {
	if (tid == 0) {
		#pragma unroll
		for (int s = 0; s < D2_STAGES; ++s) {
			d2_mbar_init(&full[s], D2_PRODUCERS);          // one cp.async arrival per producer thread
			d2_mbar_init(&empty[s], D2_CONSUMERS / 32);    // one arrival per consumer warp
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
}

// Fragments of one group of four reduction steps (k4 = 0 or 4) of the warp's 64 x (8 NI) tile.
// Gather GEMM tiles are [k][row] with a pitch of tile + 4 (the layout of the kernels above).
template<int NI, int PA, int PB>
__device__ __forceinline__ void d2_load_frags(const double* __restrict__ As, const double* __restrict__ Bs, int row0, int col0,
		int lane, int k4, double (&a)[8], double (&b)[NI]) {
	const int kq = lane & 3, g = lane >> 2;
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi) a[mi] = As[(k4 + kq) * PA + row0 + 8 * mi + g];
	#pragma unroll
	for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(k4 + kq) * PB + col0 + 8 * ni + g];
}
// Weight-gradient tiles are [row][8 m], the 16-byte chunks of a row XOR-swizzled: element (row, kk) lies at
// row * 8 + 2 * ((kk >> 1) ^ (row & 2)) + (kk & 1); lanes g = 0..3 of a half warp read rows whose bit 1 differs in pairs,
// so the 16 lanes hit 16 different 8-byte banks.
template<int NI>
__device__ __forceinline__ void d2_load_frags_sw(const double* __restrict__ As, const double* __restrict__ Bs, int row0, int col0,
		int lane, int k4, double (&a)[8], double (&b)[NI]) {
	const int kq = lane & 3, g = lane >> 2;
	const int off = 2 * (((k4 + kq) >> 1) ^ (g & 2)) + (kq & 1);   // row0, col0 and 8 * mi are multiples of 8: (row & 2) == (g & 2)
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi) a[mi] = As[(row0 + 8 * mi + g) * D2_BK + off];
	#pragma unroll
	for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(col0 + 8 * ni + g) * D2_BK + off];
}
template<int NI>
__device__ __forceinline__ void d2_mma_frags(double (&acc)[8][NI][2], const double (&a)[8], const double (&b)[NI]) {
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
}

// The weights of a gather GEMM in the order the kernel consumes them: [column tile][k-step][8 reduce rows][BN columns],
// zero filled, so that a k-block's B tile is one contiguous 8 x BN block (16-byte copies; gathered straight from the
// reference's layout every lane of a copy hit a different sector).  k-step = channel block * taps + tap.
__global__ void __launch_bounds__(256) dmma2_pack_weights_kernel(GatherGeom gg, int BN, int j_tiles, int ksteps,
		const double* __restrict__ w, double* __restrict__ wp) {
	const int T = gg.RH * gg.RW;
	const long long total = (long long) j_tiles * ksteps * D2_BK * BN;
	for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long) gridDim.x * 256) {
		const int col = (int) (i % BN);
		const long long t = i / BN;
		const int k = (int) (t % D2_BK);
		const long long u = t / D2_BK;
		const int ks = (int) (u % ksteps), jt = (int) (u / ksteps);
		const int rb = ks / T, tap = ks - rb * T;
		const int r = rb * D2_BK + k, j = jt * BN + col;
		wp[i] = r < gg.SC && j < gg.J ? w[tap * gg.w_stap + r * gg.w_sr + j * gg.w_sj] : 0.0;
	}
}

// Gather GEMM, consumer warps as WR (rows) x WC (columns), block tile (64 WR) x (8 NI WC).  VEC: the batch is even, so
// rows 2 p and 2 p + 1 of a tile are neighbours in HBM at a 16-byte aligned address.
template<int WR, int WC, int NI, bool VEC>
__global__ void __launch_bounds__(D2_THREADS, 1) dmma2_gather_gemm_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ wp, const double* __restrict__ bias, int bias_mode, double* __restrict__ out, int act_kind,
		double act_param, double* __restrict__ act_out, int j_tiles) {
	static_assert(WR * WC * 32 == D2_CONSUMERS, "eight consumer warps");
	constexpr int BM = 64 * WR, BN = 8 * NI * WC;
	constexpr int PA = BM + DM_PAD, PB = BN + DM_PAD;
	constexpr int STAGE = D2_BK * (PA + PB);
	extern __shared__ __align__(16) unsigned char d2_raw[];
	double* tiles = reinterpret_cast<double*>(d2_raw);
	uint64_t* full = reinterpret_cast<uint64_t*>(tiles + D2_STAGES * STAGE);
	uint64_t* empty = full + D2_STAGES;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	// column tiles fastest: the CTAs that gather the same rows run side by side
	const long long m0 = (long long) (blockIdx.x / j_tiles) * BM;
	const int j0 = (int) (blockIdx.x % j_tiles) * BN;
	const int R = gg.SC, J = gg.J;
	const int rblocks = (R + D2_BK - 1) / D2_BK;
	const int ksteps = gg.RH * gg.RW * rblocks;

	d2_ring_init(full, empty, tid);

	if (warp >= D2_CONSUMERS / 32) {
		// ---- producers ----
		d2_regs_producer();
		const int pt = tid - D2_CONSUMERS;
		// VEC: a thread copies ONE pair of rows (one pixel decode) over 8 / KSPLIT reduce rows; otherwise BM / 128 single rows over all 8
		constexpr int PAIRS = BM / 2, KSPLIT = VEC ? D2_PRODUCERS / PAIRS : 1, KPER = D2_BK / KSPLIT;
		constexpr int UNITS = VEC ? 1 : BM / D2_PRODUCERS;
		constexpr int BCOPIES = D2_BK * BN / 2 / D2_PRODUCERS;
		const int k_first = VEC ? (pt / PAIRS) * KPER : 0;
		const long long plane = (long long) gg.N * gg.SH * gg.SW;
		const bool unit_den = gg.denh == 1 && gg.denw == 1;
		int u_row[UNITS], u_n[UNITS], u_oh[UNITS], u_ow[UNITS];
		#pragma unroll
		for (int u = 0; u < UNITS; ++u) {
			u_row[u] = VEC ? 2 * (pt % PAIRS) : pt + D2_PRODUCERS * u;
			const long long m = m0 + u_row[u];
			if (m < M) {
				const long long pix = m / gg.N;
				u_n[u] = (int) (m - pix * gg.N);
				u_ow[u] = (int) (pix / gg.OH);
				u_oh[u] = (int) (pix - (long long) u_ow[u] * gg.OH);
			} else {
				u_n[u] = -1; u_oh[u] = 0; u_ow[u] = 0;
			}
		}
		// this column tile's packed weights: 8 x BN doubles per k-step, copied as BN / 2 pairs per reduce row
		const double* wp_tile = wp + (long long) (blockIdx.x % j_tiles) * ksteps * (D2_BK * BN);
		int rh = 0, rw = 0, r0 = 0;
		for (int ks = 0; ks < ksteps; ++ks) {
			const int s = ks & (D2_STAGES - 1);
			d2_mbar_wait(&empty[s], ((ks / D2_STAGES) & 1) ^ 1);
			double* As = tiles + s * STAGE;
			double* Bs = As + D2_BK * PA;
			#pragma unroll
			for (int u = 0; u < UNITS; ++u) {
				const int th = u_oh[u] * gg.ah + rh * gg.bh + gg.ch;
				const int tw = u_ow[u] * gg.aw + rw * gg.bw + gg.cw;
				bool ok = u_n[u] >= 0 && th >= 0 && tw >= 0;
				int ih = th, iw = tw;
				if (!unit_den) {
					ok = ok && th % gg.denh == 0 && tw % gg.denw == 0;
					ih = th / gg.denh; iw = tw / gg.denw;
				}
				ok = ok && ih < gg.SH && iw < gg.SW;
				const double* p = src + (ok ? u_n[u] + (long long) gg.N * (ih + (long long) gg.SH * iw) : 0);
				#pragma unroll
				for (int kk = 0; kk < KPER; ++kk) {
					const int k = k_first + kk;
					const bool okk = ok && r0 + k < R;
					const double* q = okk ? p + (long long) (r0 + k) * plane : src;
					if (VEC) cp_async_f64x2(&As[k * PA + u_row[u]], q, okk);
					else cp_async_f64(&As[k * PA + u_row[u]], q, okk);
				}
			}
			const double* wk = wp_tile + (long long) ks * (D2_BK * BN);
			#pragma unroll
			for (int i = 0; i < BCOPIES; ++i) {
				const int idx = pt + D2_PRODUCERS * i, k = idx / (BN / 2), cp = idx - k * (BN / 2);
				cp_async_f64x2(&Bs[k * PB + 2 * cp], wk + k * BN + 2 * cp, true);
			}
			d2_cp_async_arrive(&full[s]);
			if (++rh == gg.RH) {
				rh = 0;
				if (++rw == gg.RW) { rw = 0; r0 += D2_BK; }
			}
		}
		cp_async_wait_all_f64();
		return;
	}

	// ---- consumers ----
	d2_regs_consumer();
	const int row0 = (warp % WR) * 64, col0 = (warp / WR) * (8 * NI);
	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
	// the first fragments of k-block ks + 1 are read before the second half of k-block ks is issued: a warp never sits
	// on the barrier + fragment latency with nothing in the pipe
	double a0[8], b0[NI], a1[8], b1[NI];
	d2_mbar_wait(&full[0], 0);
	d2_load_frags<NI, PA, PB>(tiles, tiles + D2_BK * PA, row0, col0, lane, 0, a0, b0);
	for (int ks = 0; ks < ksteps; ++ks) {
		const int s = ks & (D2_STAGES - 1);
		const double* As = tiles + s * STAGE;
		d2_load_frags<NI, PA, PB>(As, As + D2_BK * PA, row0, col0, lane, 4, a1, b1);
		d2_mma_frags<NI>(acc, a0, b0);
		if (ks + 1 < ksteps) {
			const int sn = (ks + 1) & (D2_STAGES - 1);
			d2_mbar_wait(&full[sn], ((ks + 1) / D2_STAGES) & 1);
			const double* An = tiles + sn * STAGE;
			d2_load_frags<NI, PA, PB>(An, An + D2_BK * PA, row0, col0, lane, 0, a0, b0);
		}
		d2_mma_frags<NI>(acc, a1, b1);
		__syncwarp();
		if (lane == 0) d2_mbar_arrive(&empty[s]);
	}

	dmma_epilogue<NI>(acc, gg, M, J, m0 + row0, j0 + col0, lane, bias, bias_mode, out, act_kind, act_param, act_out);
}

// Weight gradient (even batch, M even): tile rows are (tap, r), columns j, the reduction runs over m in blocks of 8.
template<int WR, int WC, int NI>
__global__ void __launch_bounds__(D2_THREADS, 1) dmma2_wgrad_kernel(GatherGeom gg, const double* __restrict__ src,
		const double* __restrict__ plain, double* __restrict__ partial, long long m_per_split, long long dw_elems) {
	static_assert(WR * WC * 32 == D2_CONSUMERS, "eight consumer warps");
	constexpr int BMK = 64 * WR, BN = 8 * NI * WC;
	constexpr int STAGE = D2_BK * (BMK + BN);
	constexpr int AROWS = BMK / 32, BCOLS = BN / 32;   // per producer thread: (chunk = pt & 3, rows / columns pt >> 2 + 32 i)
	extern __shared__ __align__(16) unsigned char d2_raw[];
	double* tiles = reinterpret_cast<double*>(d2_raw);
	uint64_t* full = reinterpret_cast<uint64_t*>(tiles + D2_STAGES * STAGE);
	uint64_t* empty = full + D2_STAGES;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int R = gg.SC, J = gg.J;
	const int Ktot = gg.RH * gg.RW * R;
	const int k0 = blockIdx.x * BMK, j0 = blockIdx.y * BN;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long ms = (long long) blockIdx.z * m_per_split;
	const long long me = ms + m_per_split < M ? ms + m_per_split : M;
	const long long steps = me > ms ? (me - ms + D2_BK - 1) / D2_BK : 0;

	d2_ring_init(full, empty, tid);

	if (warp >= D2_CONSUMERS / 32) {
		// ---- producers ----
		d2_regs_producer();
		const int pt = tid - D2_CONSUMERS;
		const int ch = pt & 3, rg = pt >> 2;
		const long long plane = (long long) gg.N * gg.SH * gg.SW;
		int row_dh[AROWS], row_dw[AROWS], row_r[AROWS];
		#pragma unroll
		for (int i = 0; i < AROWS; ++i) {
			const int k = k0 + rg + 32 * i;
			if (k < Ktot) {
				const int tap = k / R, r = k - tap * R;
				const int rw = tap / gg.RH, rh = tap - rw * gg.RH;
				row_dh[i] = rh * gg.bh + gg.ch; row_dw[i] = rw * gg.bw + gg.cw; row_r[i] = r;
			} else {
				row_dh[i] = 0; row_dw[i] = 0; row_r[i] = -1;
			}
		}
		// the chunk's place in a row of the tile: rows rg + 32 i have (row & 2) == (rg & 2)
		const int dst_off = 2 * (ch ^ (rg & 2));
		for (long long st = 0; st < steps; ++st) {
			const int s = (int) (st & (D2_STAGES - 1));
			d2_mbar_wait(&empty[s], (uint32_t) (((st / D2_STAGES) & 1) ^ 1));
			double* As = tiles + s * STAGE;
			double* Bs = As + D2_BK * BMK;
			const long long m = ms + st * D2_BK + 2 * ch;
			const bool ok = m < me;   // me is even: m + 1 < me as well
			const unsigned mu = (unsigned) m, pixu = mu / (unsigned) gg.N;   // M < 2^31
			const int n = (int) (mu - pixu * (unsigned) gg.N);
			const int ow = (int) (pixu / (unsigned) gg.OH), oh = (int) (pixu - (unsigned) ow * (unsigned) gg.OH);
			const int bh = oh * gg.ah, bw = ow * gg.aw;
			#pragma unroll
			for (int i = 0; i < AROWS; ++i) {
				const int th = bh + row_dh[i], tw = bw + row_dw[i];
				const bool aok = ok && row_r[i] >= 0 && th >= 0 && tw >= 0 && th < gg.SH && tw < gg.SW;   // denh = denw = 1
				cp_async_f64x2(&As[(rg + 32 * i) * D2_BK + dst_off],
						aok ? src + n + (long long) gg.N * (th + (long long) gg.SH * tw) + row_r[i] * plane : src, aok);
			}
			#pragma unroll
			for (int i = 0; i < BCOLS; ++i) {
				const int j = j0 + rg + 32 * i;
				const bool bok = ok && j < J;
				cp_async_f64x2(&Bs[(rg + 32 * i) * D2_BK + dst_off], bok ? plain + m + M * j : plain, bok);
			}
			d2_cp_async_arrive(&full[s]);
		}
		cp_async_wait_all_f64();
		return;
	}

	// ---- consumers ----
	d2_regs_consumer();
	const int row0 = (warp % WR) * 64, col0 = (warp / WR) * (8 * NI);
	double acc[8][NI][2];
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi)
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
	double a0[8], b0[NI], a1[8], b1[NI];
	if (steps > 0) {
		d2_mbar_wait(&full[0], 0);
		d2_load_frags_sw<NI>(tiles, tiles + D2_BK * BMK, row0, col0, lane, 0, a0, b0);
	}
	for (long long st = 0; st < steps; ++st) {
		const int s = (int) (st & (D2_STAGES - 1));
		const double* As = tiles + s * STAGE;
		d2_load_frags_sw<NI>(As, As + D2_BK * BMK, row0, col0, lane, 4, a1, b1);
		d2_mma_frags<NI>(acc, a0, b0);
		if (st + 1 < steps) {
			const int sn = (int) ((st + 1) & (D2_STAGES - 1));
			d2_mbar_wait(&full[sn], (uint32_t) (((st + 1) / D2_STAGES) & 1));
			const double* An = tiles + sn * STAGE;
			d2_load_frags_sw<NI>(An, An + D2_BK * BMK, row0, col0, lane, 0, a0, b0);
		}
		d2_mma_frags<NI>(acc, a1, b1);
		__syncwarp();
		if (lane == 0) d2_mbar_arrive(&empty[s]);
	}

	double* dst = partial + (long long) blockIdx.z * dw_elems;
	const int g = lane >> 2, kq = lane & 3;
	#pragma unroll
	for (int mi = 0; mi < 8; ++mi) {
		const int k = k0 + row0 + 8 * mi + g;
		if (k >= Ktot) continue;
		const int tap = k / R, r = k - tap * R;
		const long long base = tap * gg.w_stap + r * gg.w_sr;
		#pragma unroll
		for (int ni = 0; ni < NI; ++ni) {
			#pragma unroll
			for (int e = 0; e < 2; ++e) {
				const int j = j0 + col0 + 8 * ni + 2 * kq + e;
				if (j < J) dst[base + j * gg.w_sj] = acc[mi][ni][e];
			}
		}
	}
}

// CATTL3_DMMA_V1=1 brings the barrier-per-k-block kernels back (A/B runs)
bool dmma_v2_enabled() { return getenv("CATTL3_DMMA_V1") == nullptr; }
// CATTL3_DMMA_TRACE=1 names the kernel of every launch on stderr (tests check which version ran)
void dmma_trace(const char* what, int wr, int wc, int vec) {
	if (getenv("CATTL3_DMMA_TRACE")) fprintf(stderr, "dmma2 %s tile %d x %d%s\n", what, 64 * wr, 32 * wc, vec ? " vec" : "");
}

template<int WR, int WC, int NI, bool VEC>
int launch_gather2(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* w, const double* bias, int bias_mode,
		double* out, int act_kind, double act_param, double* act_out) {
	constexpr int BM = 64 * WR, BN = 8 * NI * WC;
	constexpr size_t smem = sizeof(double) * D2_STAGES * D2_BK * (BM + BN + 2 * DM_PAD) + 2 * D2_STAGES * sizeof(uint64_t);
	auto kern = dmma2_gather_gemm_kernel<WR, WC, NI, VEC>;
	static bool configured = false;
	if (!configured) {
		CATTL3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured = true;
	}
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const int j_tiles = (int) ceil_div(gg.J, BN);
	const long long ctas = ceil_div(M, BM) * j_tiles;
	CATTL3_REQUIRE(ctas < (1ll << 31), "gather GEMM: grid too large");
	const int ksteps = gg.RH * gg.RW * (int) ceil_div(gg.SC, D2_BK);
	const long long wp_elems = (long long) j_tiles * ksteps * D2_BK * BN;
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->tc_w, &ctx->tc_w_bytes, (size_t) wp_elems * sizeof(double)));
	dmma2_pack_weights_kernel<<<ew_grid(ctx, wp_elems, 256), 256, 0, ctx->stream>>>(gg, BN, j_tiles, ksteps, w, (double*) ctx->tc_w);
	CATTL3_LAUNCHED(ctx);
	dmma_trace("gather", WR, WC, VEC);
	kern<<<(unsigned) ctas, D2_THREADS, smem, ctx->stream>>>(gg, src, (const double*) ctx->tc_w, bias, bias_mode, out, act_kind, act_param,
			act_out, j_tiles);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

template<int WR, int WC, int NI>
int launch_wgrad2(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* plain, double* dw) {
	constexpr int BMK = 64 * WR, BN = 8 * NI * WC;
	constexpr size_t smem = sizeof(double) * D2_STAGES * D2_BK * (BMK + BN) + 2 * D2_STAGES * sizeof(uint64_t);
	auto kern = dmma2_wgrad_kernel<WR, WC, NI>;
	static bool configured = false;
	if (!configured) {
		CATTL3_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		configured = true;
	}
	const int Ktot = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) Ktot * gg.J;
	const long long gx = ceil_div(Ktot, BMK), gy = ceil_div(gg.J, BN);
	// one CTA per SM: as many splits of the reduction as make the grid one wave
	long long splits = ctx->sm_count / (gx * gy);
	if (splits < 1) splits = 1;
	const long long max_splits = ceil_div(M, 1024);
	if (splits > max_splits) splits = max_splits;
	if (splits > 65535) splits = 65535;
	const long long m_per_split = ceil_div(ceil_div(M, splits), D2_BK) * D2_BK;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(double)));
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	dmma_trace("wgrad", WR, WC, 1);
	kern<<<grid, D2_THREADS, smem, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	dmma_wgrad_reduce_kernel<<<ew_grid(ctx, elems, 256), 256, 0, ctx->stream>>>((const double*) ctx->ws, (int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace

bool dmma_gather_gemm_supported(const GatherGeom& gg) {
	return gg.J > 32 && gg.SC >= 6 && (long long) gg.RH * gg.RW * gg.SC >= 32;
}

int dmma_gather_gemm(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* w, const double* bias,
		int bias_mode, double* out, const EpilogueArgs* ep) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const bool act = ep && ep->act_kind != CATTL3_ACT_NONE;
	CATTL3_REQUIRE(out || act, "gather GEMM: no output tensor");
	double* act_out = act ? (double*) ep->act_out : nullptr;
	const int act_kind = act ? ep->act_kind : CATTL3_ACT_NONE;
	const double act_param = act ? ep->act_param : 0.0;
	if (dmma_v2_enabled()) {
		// version 2 (producer warps): 128 x 128 tiles, or 256 x 64 where 64-wide column tiles waste less; only where the
		// grid fills the device (small layers keep the two-CTAs-per-SM kernel below)
		const bool narrow = ceil_div(gg.J, 64) * 64 < ceil_div(gg.J, 128) * 128;
		const long long ctas = narrow ? ceil_div(M, 256) * ceil_div(gg.J, 64) : ceil_div(M, 128) * ceil_div(gg.J, 128);
		if (ctas >= ctx->sm_count) {
			const bool vec = gg.N % 2 == 0;
			if (narrow)
				return vec ? launch_gather2<4, 2, 4, true>(ctx, gg, src, w, bias, bias_mode, out, act_kind, act_param, act_out)
						: launch_gather2<4, 2, 4, false>(ctx, gg, src, w, bias, bias_mode, out, act_kind, act_param, act_out);
			return vec ? launch_gather2<2, 4, 4, true>(ctx, gg, src, w, bias, bias_mode, out, act_kind, act_param, act_out)
					: launch_gather2<2, 4, 4, false>(ctx, gg, src, w, bias, bias_mode, out, act_kind, act_param, act_out);
		}
	}
	// 128 x 64 tiles whatever the filter count: 122 registers -> two CTAs (16 warps) per SM hide the fragment-load
	// latency; measured at config 2 (256 filters): 23.3 TFLOP/s against 18.5 with 128 x 128 tiles and one CTA per SM
	dim3 grid((unsigned) ceil_div(M, DM_BM), (unsigned) ceil_div(gg.J, 64));
	dmma_gather_gemm_kernel<2><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, w, bias, bias_mode, out, act_kind, act_param,
			act_out);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

bool dmma_wgrad_supported(const GatherGeom& gg) {
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	return gg.J > 32 && (long long) gg.RH * gg.RW * gg.SC >= 64 && M >= 1024 && M < (1ll << 31) && gg.denh == 1 && gg.denw == 1;
}

int dmma_wgrad(cattl3_ctx* ctx, const GatherGeom& gg, const double* src, const double* plain, double* dw) {
	const int Ktot = gg.RH * gg.RW * gg.SC;
	const long long M = (long long) gg.N * gg.OH * gg.OW;
	const long long elems = (long long) Ktot * gg.J;
	if (dmma_v2_enabled() && gg.N % 2 == 0 && M >= 16384) {
		// version 2 (producer warps, 16-byte copies of batch pairs): the tile shape that pads taps x channels and filters least
		auto padded = [&](int bmk, int bn) { return ceil_div(Ktot, bmk) * bmk * ceil_div(gg.J, bn) * bn; };
		const long long c128 = padded(128, 128), c64 = padded(64, 256), c256 = padded(256, 64);
		if (c128 <= c64 && c128 <= c256) return launch_wgrad2<2, 4, 4>(ctx, gg, src, plain, dw);
		if (c64 <= c256) return launch_wgrad2<1, 8, 4>(ctx, gg, src, plain, dw);
		return launch_wgrad2<4, 2, 4>(ctx, gg, src, plain, dw);
	}
	// here the wide tile wins (19.3 against 17.0 TFLOP/s at 256 filters; a 512-thread 128 x 128 variant with 64 x 16 warp
	// tiles measured 17.2): half as many passes over the gathered rows
	const int BN = gg.J > 64 ? 128 : 64;
	const long long gx = ceil_div(Ktot, DM_BM), gy = ceil_div(gg.J, BN);
	long long splits = (long long) (BN == 128 ? 1 : 2) * ctx->sm_count / (gx * gy);
	if (splits < 1) splits = 1;
	const long long max_splits = ceil_div(M, 1024);
	if (splits > max_splits) splits = max_splits;
	long long m_per_split = ceil_div(ceil_div(M, splits), DM_BK) * DM_BK;
	splits = ceil_div(M, m_per_split);
	CATTL3_CHECK(ensure_buffer(ctx, &ctx->ws, &ctx->ws_bytes, (size_t) (splits * elems) * sizeof(double)));
	dim3 grid((unsigned) gx, (unsigned) gy, (unsigned) splits);
	if (BN == 128)
		dmma_wgrad_kernel<4><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	else
		dmma_wgrad_kernel<2><<<grid, DM_THREADS, 0, ctx->stream>>>(gg, src, plain, (double*) ctx->ws, m_per_split, elems);
	CATTL3_LAUNCHED(ctx);
	dmma_wgrad_reduce_kernel<<<ew_grid(ctx, elems, 256), 256, 0, ctx->stream>>>((const double*) ctx->ws, (int) splits, elems, dw);
	CATTL3_LAUNCHED(ctx);
	return CATTL3_OK;
}

} // namespace cattl3
