#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json configs[1]:

  train samples/s (fwd + bwd + optimizer step) of a single ConvKernelLayer 3x3 (pad 1, stride 1),
  56x56x64 -> 256 filters, batch 256 per GPU, float, synthetic data; conv TFLOP/s vs the tensor peak.

One step = ConvKernelLayer::pass_forward + pass_back (weight, bias and input gradients) + one fused
Nadam update of (W, b) (the reference examples' optimizer), all through the C ABI of
include/cattl3_b200.h.  For N > 1 (torchrun, one rank per GPU) every rank runs its own batch-256 shard
(weak scaling) and the gradient arena [dW | db] is all-reduced with NCCL before the optimizer step.

  value : samples/s with the inputs resident in HBM (x 205 MB + dY 822 MB per rank > the 126 MB L2)
  e2e   : the same step through the host-buffer entry points (cattl3_conv_*_host_f32): x and dY come
          from pinned host memory and y and dX go back to the host inside the timed region
  roofline / kernels : per-GPU algorithmic TFLOP/s (2*M*K*F per GEMM pass) of each device kernel,
          timed alone with CUDA events on the launching stream, against the TF32 tensor peak
          (= measured bf16 peak / 2, MEASURED_PEAKS.json); 3xTF32 issues 3 MMAs per algorithmic MAC,
          so tensor_pipe_frac = 3 * frac
  cpu_baseline : the unmodified reference (oracle/_ref, Eigen + OpenMP) on the box's host cores on a
          bounded sample of the same workload (rank 0, N=1 only)

`--impl reference` times that CPU reference alone (rank 0; other ranks exit).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BATCH, H, W, C, F, R = 256, 56, 56, 64, 256, 3
GEOM = (N_BATCH, H, W, C, F, R, R, 1, 1, 1, 1, 0, 0)
M = N_BATCH * H * W
K = R * R * C
FLOP_PER_PASS = 2.0 * M * K * F           # 236.76 GFLOP: fwd, wgrad, dgrad each
NADAM = (1e-3, 1e-1, 1e-3, 1e-5)          # NadamOptimizer defaults, NadamOptimizer.hpp:39-42
WORKLOAD = "ConvKernelLayer 3x3 pad1 stride1, 56x56x64->256, batch 256/GPU, float32, fwd+bwd+Nadam step"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # NVML unavailable: report that instead of inventing numbers
            self.reasons.add("nvml_unavailable: %s" % type(e).__name__)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_sample(n_sample, reps, keep=None):
    """Times the unmodified reference (or, if its shim is absent, the C restatement) on config 2 at
    batch n_sample: returns (samples/s, cores, kind, description)."""
    import numpy as np
    from oracle import binding
    kind = "reference" if binding.have_ref() else "port"
    lib = binding.Oracle("ref" if kind == "reference" else "orc")
    g = binding.Geom(n_sample, *GEOM[1:])
    rng = np.random.default_rng(2001)
    x = np.asfortranarray(rng.uniform(-1, 1, (n_sample, H, W, C)).astype(np.float32))
    w = np.asfortranarray((rng.standard_normal((K, F)) * (2.0 / K) ** 0.5).astype(np.float32))
    b = np.zeros((1, F), np.float32, order="F")
    dy = np.asfortranarray(rng.uniform(-1, 1, (n_sample, H, W, F)).astype(np.float32))
    best = None
    for i in range(reps + 1):
        t0 = time.perf_counter()
        r = lib.conv(g, x, w, b, dy)
        lib.optimizer(8, NADAM, 0.0, w, [r["dw"]], 1)
        dt = time.perf_counter() - t0
        if i > 0:  # first pass is warm-up
            best = dt if best is None else min(best, dt)
    desc = "batch %d of 256 (same layer), fwd+bwd+Nadam, best of %d after 1 warm-up, %s" % (
        n_sample, reps, "C-ATTL3 Eigen path, -O3 -mavx2 -mfma -fopenmp" if kind == "reference"
        else "C restatement of the reference (reference shim not built)")
    if keep is not None:
        keep.update(x=x, w=w, b=b, dy=dy, out=r)
    return n_sample / best, lib.num_threads(), kind, desc


def parity_against(ctx, pkg, keep, n_sample):
    """The CPU baseline's own outputs against the CUDA path on the same inputs (the sample batch of the same layer):
    norm-relative max|a - b| / max|b| per tensor, the metric of tests/cases.py."""
    import numpy as np
    import torch
    g = pkg.ConvGeom(n_sample, *GEOM[1:])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda()
    xd, wd, bd, dyd = dev(keep["x"]), dev(keep["w"]), dev(keep["b"]), dev(keep["dy"])
    yd = torch.empty(keep["dy"].size, device="cuda")
    dxd = torch.empty(keep["x"].size, device="cuda")
    dwd, dbd = torch.zeros(keep["w"].size, device="cuda"), torch.zeros(keep["b"].size, device="cuda")
    ctx.conv_forward(g, xd, wd, bd, yd)
    ctx.conv_backward(g, xd, wd, dyd, dwd, dbd, dxd)
    torch.cuda.synchronize()
    out = {}
    for k, t in (("y", yd), ("dx", dxd), ("dw", dwd), ("db", dbd)):
        ref = keep["out"][k].ravel(order="F").astype(np.float64)
        out[k] = float(np.max(np.abs(t.cpu().numpy().astype(np.float64) - ref)) / np.max(np.abs(ref)))
    out["tolerance"] = 1e-4
    out["against"] = "cpu_baseline outputs (same inputs, batch %d), kernel path %s" % (n_sample, ctx.last_path)
    return out


def measured_tf32_peak():
    """scripts/tf32_peak (built by __graft_entry__.build()): dense tcgen05.mma kind::tf32 with A from tensor memory, the
    instruction the kernels issue, verified and timed on this GPU right now.  None if the probe is not there."""
    import subprocess
    exe = os.path.join(ROOT, "scripts", "tf32_peak")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "0.5", "1sm_ts"], capture_output=True, text=True, timeout=60).stdout
        d = json.loads(out.strip().splitlines()[-1])["1sm_ts"]
        if d.get("verify_max_abs_err") != 0:
            return None
        return {"burst": d["burst_tflops"], "sustained": d["sustained_tflops"]}
    except Exception:
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    n_sample = 16
    t0 = time.perf_counter()
    vals = []
    for _ in range(max(1, args.warmup) + max(1, args.steps)):
        v, cores, kind, desc = reference_sample(n_sample, 1)
        vals.append(v)
        if time.perf_counter() - t0 > 150:
            break
    vals = vals[max(1, args.warmup):] or vals
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": "train samples/s (fwd+bwd+step)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * n_sample / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": desc},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def pin_to_gpu_numa_node(index):
    """Runs this rank on the CPU cores next to its GPU (NVML's ideal affinity) BEFORE any pinned host memory is allocated:
    first touch then places the staging buffers of the host-buffer path on the GPU's own NUMA node, so that eight ranks
    streaming 2 GB per step each do not all cross the socket interconnect.  Returns the number of cores, or None."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cores = [64 * w + b for w in range(words) for b in range(64) if (mask[w] >> b) & 1]
        allowed = sorted(set(cores) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return None


def double_extras(ctx, pkg, torch, dev, stream):
    """BASELINE.json configs[1] names float AND double: the same layer at the same size in double (FP64 tensor cores,
    mma.sync m8n8k4 -- tcgen05 has no fp64 kind), each pass timed alone, against the DMMA rate scripts/dmma_peak measures now."""
    import subprocess
    g = pkg.ConvGeom(*GEOM)
    gen = torch.Generator(device=dev).manual_seed(4001)
    x = torch.rand(M * C, device=dev, generator=gen, dtype=torch.float64) * 2 - 1
    dy = torch.rand(M * F, device=dev, generator=gen, dtype=torch.float64) * 2 - 1
    w = torch.randn(K * F, device=dev, generator=gen, dtype=torch.float64) * (2.0 / K) ** 0.5
    b = torch.zeros(F, device=dev, dtype=torch.float64)
    y = torch.empty(M * F, device=dev, dtype=torch.float64)
    dx = torch.empty(M * C, device=dev, dtype=torch.float64)
    dw, db = torch.zeros(K * F, device=dev, dtype=torch.float64), torch.zeros(F, device=dev, dtype=torch.float64)

    def time_alone(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        bb.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(bb) / reps

    t = {"forward": time_alone(lambda: ctx.conv_forward(g, x, w, b, y))}
    path = ctx.last_path
    t["weight+bias gradient"] = time_alone(lambda: ctx.conv_backward(g, x, w, dy, dw, db, None))
    t["input gradient"] = time_alone(lambda: ctx.conv_backward(g, x, w, dy, None, None, dx))
    del x, dy, y, dx
    torch.cuda.empty_cache()
    peak = None
    exe = os.path.join(ROOT, "scripts", "dmma_peak")
    if os.path.exists(exe):
        try:
            peak = json.loads(subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1])[
                "dmma_m8n8k4_tflops"]
        except Exception:
            peak = None
    out = {"dtype": "f64", "kernel_path": path, "peak_tflops": peak,
           "peak_source": "scripts/dmma_peak (mma.sync.m8n8k4.f64 chains on every SM) run now" if peak else "not measured",
           "kernels": []}
    for name, ms in t.items():
        ach = FLOP_PER_PASS / (ms * 1e-3) / 1e12
        out["kernels"].append({"pass": name, "ms": round(ms, 3), "achieved_tflops": round(ach, 2),
                               "frac": round(ach / peak, 4) if peak else None})
    out["samples_per_s"] = round(N_BATCH / (sum(t.values()) * 1e-3), 1)
    return out


def network_extras(world, dist=None):
    """Configs 4 and 5 through the C++ batch loop (scripts/bench_networks.py), each in a child process of this rank with a
    time limit: a secondary number must never be able to take the headline line down with it.  The ranks' children find each
    other through an explicit NCCL id file (their parents differ, so the default name, which is tied to the parent, would not match)."""
    import subprocess
    out = {}
    steps = 8 if world <= 2 else 4
    script = os.path.join(ROOT, "scripts", "bench_networks.py")
    nonce = [os.getpid()]
    if dist is not None:
        dist.broadcast_object_list(nonce, src=0)   # rank 0's pid names the id files of this launch: a leftover file never matches
    for cfg in (4, 5):
        env = dict(os.environ)
        env["CATTL3_COMM_ID_FILE"] = "/tmp/cattl3_nccl_id.bench.%s.%d.%d" % (os.environ.get("MASTER_PORT", "0"), nonce[0], cfg)
        try:
            r = subprocess.run([sys.executable, script, "--config", str(cfg), "--steps", str(steps), "--epochs", "3"], env=env,
                               capture_output=True, text=True, timeout=240)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0:
                out["config%d" % cfg] = {"error": "exit code %d: %s" % (r.returncode, r.stderr.strip()[-200:])}
            elif lines:
                d = json.loads(lines[-1])
                out["config%d" % cfg] = {k: d[k] for k in ("value", "unit", "n_gpus", "ms_per_step", "epoch_ms", "scaling", "config")}
        except subprocess.TimeoutExpired:
            out["config%d" % cfg] = {"error": "no result within 240 s"}
        except Exception as e:   # the shim is built by __graft_entry__.build(); say why if it cannot run
            out["config%d" % cfg] = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 auto (tcgen05), 1 SIMT only")
    ap.add_argument("--no-networks", action="store_true", help="skip the network-level extras (configs 4 and 5)")
    ap.add_argument("--no-double", action="store_true", help="skip the double-precision extras (N = 1 only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    all_cores = os.sched_getaffinity(0)
    numa_cores = pin_to_gpu_numa_node(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(local_rank, stream.cuda_stream)
    ctx.set_conv_path(args.path)
    g = pkg.ConvGeom(*GEOM)

    gen = torch.Generator(device=dev).manual_seed(2001 + rank)
    x = torch.rand(M * C, device=dev, generator=gen) * 2 - 1              # x uniform[-1,1), seed 2001
    dy = torch.rand(M * F, device=dev, generator=gen) * 2 - 1             # dY uniform[-1,1)
    wgen = torch.Generator(device=dev).manual_seed(7)                     # identical weights on every rank
    arena = torch.zeros(K * F + F, device=dev)                            # parameters [W | b]
    arena[:K * F] = torch.randn(K * F, device=dev, generator=wgen) * (2.0 / K) ** 0.5   # He init
    grads = torch.zeros_like(arena)                                       # gradient arena [dW | db]
    m_state, v_state = torch.zeros_like(arena), torch.zeros_like(arena)
    w, b = arena[:K * F], arena[K * F:]
    dw, db = grads[:K * F], grads[K * F:]
    y = torch.empty(M * F, device=dev)
    dx = torch.empty(M * C, device=dev)
    tstep = [0]
    # the product's communicator (cattl3_comm_*: NCCL behind the C ABI); torch.distributed only synchronises the ranks
    # around the timed region and takes the max of their times
    comm = pkg.Comm(ctx) if world > 1 else None

    def step():
        ctx.conv_forward(g, x, w, b, y)
        ctx.conv_backward(g, x, w, dy, dw, db, None)        # weight + bias gradient
        if comm is not None:
            # sum over ranks (the loss gradient is divided by the GLOBAL batch upstream), on the communicator's side
            # stream: the exchange overlaps the input gradient, which does not depend on it
            comm.allreduce(grads, asynchronous=True)
        ctx.conv_backward(g, x, w, dy, None, None, dx)      # input gradient
        if comm is not None:
            comm.wait()
        st = pkg.make_opt_step(pkg.OPT["nadam"], NADAM, tstep[0], 0, 0.0, True)
        ctx.optimizer_step(st, arena.numel(), arena, grads, m_state, v_state)
        tstep[0] += 1

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    path_used = ctx.last_path
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    launches = ctx.launches - l0
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * N_BATCH / (ms_per_step / 1000.0)

    # ---- e2e: host buffers through the *_host entry points ------------------------------------------------
    e2e_steps = max(2, min(args.steps, 5))
    xh = torch.empty(M * C, pin_memory=True).copy_(x.cpu())
    dyh = torch.empty(M * F, pin_memory=True).copy_(dy.cpu())
    yh = torch.empty(M * F, pin_memory=True)
    dxh = torch.empty(M * C, pin_memory=True)
    xkeep = torch.empty(M * C, device=dev)

    def e2e_step():
        # enqueue the whole step (uploads, kernels, downloads on the library's three streams), then wait for the host
        # tensors: y and dX are complete in pinned host memory when the step ends
        ctx.conv_forward_host_async(g, xh, w, b, yh, xkeep)
        ctx.conv_backward_host_async(g, xkeep, w, dyh, dw, db, dxh)
        if comm is not None:
            comm.allreduce(grads)
        st = pkg.make_opt_step(pkg.OPT["nadam"], NADAM, tstep[0], 0, 0.0, True)
        ctx.optimizer_step(st, arena.numel(), arena, grads, m_state, v_state)
        tstep[0] += 1
        ctx.host_wait()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * N_BATCH * e2e_steps / e2e_s
    h2d = 4 * (M * C + M * F)
    d2h = 4 * (M * F + M * C)

    # ---- per-kernel rooflines: each device pass timed alone on the launching stream ---------------------
    pk = peaks()
    torch.cuda.synchronize()
    # rank 0 measures the denominator on its own GPU at every world size (the other ranks wait: their lines are not printed,
    # and a probe on a busy box would not be a peak)
    probe = measured_tf32_peak() if rank == 0 else None
    barrier()
    if probe:
        tf32_peak = probe["burst"]
        peak_source = ("dense tcgen05.mma kind::tf32 (M128 N256 K8, A from tensor memory) measured by scripts/tf32_peak in this "
                       "run: %.1f TFLOP/s burst (kernels timed alone), %.1f sustained" % (probe["burst"], probe["sustained"]))
    else:
        tf32_peak = pk["bf16_burst"] / 2.0
        peak_source = "TF32 dense = bf16 burst %.1f / 2, %s (scripts/tf32_peak not run)" % (pk["bf16_burst"], pk["source"])

    def time_alone(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        bb.record(stream)
        torch.cuda.synchronize()
        return a.elapsed_time(bb) / reps

    t_fwd = time_alone(lambda: ctx.conv_forward(g, x, w, b, y))
    t_bwd_nodx = time_alone(lambda: ctx.conv_backward(g, x, w, dy, dw, db, None))
    grads.zero_()
    st = pkg.make_opt_step(pkg.OPT["nadam"], NADAM, tstep[0], 0, 0.0, True)
    t_opt = time_alone(lambda: ctx.optimizer_step(st, arena.numel(), arena, grads, m_state, v_state))
    kernels = []
    t_dgrad = time_alone(lambda: ctx.conv_backward(g, x, w, dy, None, None, dx))
    for name, t, flop in (("conv forward (pack_weights + tc_gather_gemm_kernel)", t_fwd, FLOP_PER_PASS),
                          ("weight+bias gradient (tc_wgrad_kernel + wgrad_reduce_tc_kernel)", t_bwd_nodx, FLOP_PER_PASS),
                          ("input gradient (pack_weights + tc_gather_gemm_kernel)", t_dgrad, FLOP_PER_PASS)):
        ach = flop / (t * 1e-3) / 1e12
        kernels.append({"pass": name, "ms": round(t, 4), "achieved_tflops": round(ach, 2),
                        "frac": round(ach / tf32_peak, 4), "tensor_pipe_frac": round(3 * ach / tf32_peak, 4)})
    opt_bytes = 7 * 4 * arena.numel()
    kernels.append({"pass": "fused Nadam step (opt_step_kernel)", "ms": round(t_opt, 4),
                    "achieved_gbs": round(opt_bytes / (t_opt * 1e-3) / 1e9, 2), "bound": "launch latency (0.6 MB)"})
    dom = max(kernels[:3], key=lambda k: k["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(dom["pass"].split(" (")[0])
    roofline = {"bound": "tensor", "achieved": dom["achieved_tflops"], "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                "frac": dom["frac"], "traffic": traffic, "kernel": dom["pass"],
                "tensor_pipe_frac": dom["tensor_pipe_frac"],
                "peak_source": peak_source + "; 3xTF32 issues 3 MMAs per algorithmic MAC, so frac <= 1/3 and "
                               "tensor_pipe_frac = 3*frac"}

    line = {
        "metric": "train samples/s (fwd+bwd+step)", "value": round(value, 1), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * N_BATCH, "parallelism": "dp%d" % world,
                   "kernel_path": path_used, "l2_policy": "inputs larger than L2 (x 205 MB + dY 822 MB per rank)",
                   "conv_tflops_per_gpu": round(3 * FLOP_PER_PASS / (ms_per_step * 1e-3) / 1e12, 2)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "host_cores_near_gpu": numa_cores, "api": "cattl3_conv_forward_host_async_f32 + cattl3_conv_backward_host_async_f32 + optimizer step + "
                       "cattl3_host_wait per step (uploads / kernels / downloads pipelined over filter chunks on three streams)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "kernels": kernels,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        keep = {}
        os.sched_setaffinity(0, all_cores)   # the CPU baseline gets every host core again, not just the ones next to the GPU
        v, cores, kind, desc = reference_sample(32, 2, keep)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": "samples/s", "cores": cores, "kind": kind, "sample": desc}
        line["parity"] = parity_against(ctx, pkg, keep, 32)
    if rank == 0 and world == 1 and not args.no_double:
        line["double"] = double_extras(ctx, pkg, torch, dev, stream)
    if comm is not None:
        comm.destroy()
    if not args.no_networks:
        # BASELINE.json configs[3] and [4] through the product's own data-parallel batch loop (cattle::SGDOptimizer::_train
        # sharding + cattl3_comm_* exchange overlapped with the backward pass + synchronised BatchNorm; C++ host side,
        # scripts/bench_networks.py), same launch, weak scaling at 64 samples per GPU: secondary numbers beside the headline
        line["networks"] = network_extras(world, dist)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
