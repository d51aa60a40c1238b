#!/usr/bin/env python
"""Per-op timings on one B200 through the C ABI: the config-2 sweep of the kernel layers
(3x3 / 1x1, stride {1,2}, dilation {0,1}; float on tcgen05, double on DFMA) as algorithmic TFLOP/s, and
the memory-bound layers (activations, pooling, batch normalisation, residual add, optimizer step) as
algorithmic GB/s against the measured HBM peak.  Each op is timed alone with CUDA events on the launching
stream after a warm-up; tensors are larger than L2.  Prints one JSON object per line.

  python scripts/bench_ops.py [--only conv|mem] [--double] > gpurun_out/ops.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
from bench import peaks  # noqa: E402


def timed(stream, fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="all", choices=["all", "conv", "small", "mem", "fused"])
    ap.add_argument("--double", action="store_true", help="also run the double (DFMA) kernel-layer sweep")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 SIMT, 3 big-tile FMA (no tensor cores)")
    args = ap.parse_args()
    pkg = load_package()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(0, stream.cuda_stream)
    ctx.set_conv_path(args.path)
    pk = peaks()
    out = lambda d: print(json.dumps(d), flush=True)

    if args.only in ("all", "conv", "small"):
        sweeps = [("3x3 s1 d0 (config 2)", (256, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)),
                  ("3x3 s2 d0", (256, 56, 56, 64, 256, 3, 3, 1, 1, 2, 2, 0, 0)),
                  ("3x3 s1 d1", (256, 56, 56, 64, 256, 3, 3, 2, 2, 1, 1, 1, 1)),
                  ("1x1 s1", (256, 56, 56, 64, 256, 1, 1, 0, 0, 1, 1, 0, 0)),
                  ("3x3 s1 d0 256->256 28x28", (256, 28, 28, 256, 256, 3, 3, 1, 1, 1, 1, 0, 0))]
        if args.only == "small":
            # the kernel layers of the network configs (4: ResNet modules at batch 64; 5: conv lanes over 512 frames and the
            # LSTM gate kernels at batch 64): small GEMMs where fixed costs decide; run with --reps 50 and --path 0 / 1 / 3
            sweeps = [("config 5 LSTM input kernel 16x16x64->16", (64, 16, 16, 64, 16, 3, 3, 1, 1, 1, 1, 0, 0)),
                      ("config 5 LSTM state kernel 16x16x16->16", (64, 16, 16, 16, 16, 3, 3, 1, 1, 1, 1, 0, 0)),
                      ("config 5 dense module 32x32x32->16", (512, 32, 32, 32, 16, 3, 3, 1, 1, 1, 1, 0, 0)),
                      ("config 5 lane 32x32x3->16", (512, 32, 32, 3, 16, 3, 3, 1, 1, 1, 1, 0, 0)),
                      ("config 4 module 56x56x64->64 n64", (64, 56, 56, 64, 64, 3, 3, 1, 1, 1, 1, 0, 0))]
        for dtype in [torch.float32] + ([torch.float64] if args.double else []):
            for name, geom in sweeps:
                g = pkg.ConvGeom(*geom)
                n, h, w, c, f, rh, rw, ph, pw, sh, sw, dh, dw_ = geom
                oh = (h - rh - (rh - 1) * dh + 2 * ph) // sh + 1
                ow = (w - rw - (rw - 1) * dw_ + 2 * pw) // sw + 1
                m, k = n * oh * ow, rh * rw * c
                flop = 2.0 * m * k * f
                x = torch.rand(n * h * w * c, device=dev, dtype=dtype) * 2 - 1
                dy = torch.rand(m * f, device=dev, dtype=dtype) * 2 - 1
                wt = torch.randn(k * f, device=dev, dtype=dtype) * (2.0 / k) ** 0.5
                b = torch.zeros(f, device=dev, dtype=dtype)
                y = torch.empty(m * f, device=dev, dtype=dtype)
                dx = torch.empty(n * h * w * c, device=dev, dtype=dtype)
                dwt, db = torch.zeros_like(wt), torch.zeros_like(b)
                reps = args.reps if dtype == torch.float32 else 2
                t_f = timed(stream, lambda: ctx.conv_forward(g, x, wt, b, y), reps)
                path_f = ctx.last_path
                t_w = timed(stream, lambda: ctx.conv_backward(g, x, wt, dy, dwt, db, None), reps)
                t_b = timed(stream, lambda: ctx.conv_backward(g, x, wt, dy, dwt, db, dx), reps)
                path_b = ctx.last_path
                tf = lambda t: round(flop / (t * 1e-3) / 1e12, 2)
                out({"op": "ConvKernelLayer " + name, "dtype": str(dtype).split(".")[-1], "M": m, "K": k, "F": f,
                     "gflop_per_pass": round(flop / 1e9, 2), "path_fwd": path_f, "path_bwd": path_b,
                     "fwd_ms": round(t_f, 4), "wgrad_ms": round(t_w, 4), "dgrad_ms": round(t_b - t_w, 4),
                     "fwd_tflops": tf(t_f), "wgrad_tflops": tf(t_w), "dgrad_tflops": tf(t_b - t_w)})
                del x, dy, wt, y, dx

    if args.only in ("all", "fused"):
        # The forward chains of a network, fused epilogue against layer by layer (config 2 geometry, float, tcgen05):
        #   ConvKernelLayer -> ReLU                and   ConvKernelLayer -> BatchNormLayer (training) -> ReLU.
        for geom in ((256, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0), (256, 56, 56, 64, 64, 3, 3, 1, 1, 1, 1, 0, 0)):
            g = pkg.ConvGeom(*geom)
            n, h, w, c, f = geom[:5]
            m, k = n * h * w, 9 * c
            dt = torch.float32
            x = torch.rand(m * c, device=dev, dtype=dt) * 2 - 1
            wt = torch.randn(k * f, device=dev, dtype=dt) * (2.0 / k) ** 0.5
            b = torch.rand(f, device=dev, dtype=dt)
            y, a, z = (torch.empty(m * f, device=dev, dtype=dt) for _ in range(3))
            gm, bt = torch.ones(f, device=dev, dtype=dt), torch.zeros(f, device=dev, dtype=dt)
            rm, rs, sm, ss = (torch.zeros(f, device=dev, dtype=dt) for _ in range(4))
            stats = torch.zeros(2 * f, device=dev, dtype=torch.float64)
            E = m * f * 4
            relu = pkg.ACT["relu"]
            cases = [
                ("conv -> ReLU, layer by layer", lambda: (ctx.conv_forward(g, x, wt, b, y),
                                                          ctx.activation_forward(relu, 0.0, n, h * w * f, y, a)), 3 * E),
                ("conv -> ReLU, fused epilogue (pre-activation kept for backward)",
                 lambda: ctx.conv_forward_fused(g, x, wt, b, y, act_kind=relu, act_out=a), 2 * E),
                ("conv -> ReLU, fused epilogue (inference: activated output only)",
                 lambda: ctx.conv_forward_fused(g, x, wt, b, None, act_kind=relu, act_out=a), E),
                ("conv -> BatchNorm(train) -> ReLU, layer by layer",
                 lambda: (ctx.conv_forward(g, x, wt, b, y),
                          ctx.batchnorm_forward(1, n, h, w, f, 1, 1, 0.1, 1e-5, y, gm, bt, rm, rs, sm, ss, z),
                          ctx.activation_forward(relu, 0.0, n, h * w * f, z, a)), 6 * E),
                ("conv (statistics in the epilogue) -> BatchNorm apply + ReLU in one pass",
                 lambda: (ctx.conv_forward_fused(g, x, wt, b, y, col_stats=stats),
                          ctx.batchnorm_forward_stats(1, n, h, w, f, 1, 0.1, 1e-5, y, stats, b, gm, bt, rm, rs, sm, ss, z,
                                                      act_kind=relu, act_out=a)), 4 * E),
            ]
            for name, fn, act_bytes in cases:
                t = timed(stream, fn, args.reps)
                out({"op": name, "shape": "N=%d %dx%dx%d -> %d, 3x3" % (n, h, w, c, f), "dtype": "float32", "ms": round(t, 4), "path": ctx.last_path,
                     "activation_tensor_bytes_moved": act_bytes,
                     "conv_gflop": round(2.0 * m * k * f / 1e9, 2)})
            del x, wt, y, a, z

    if args.only in ("all", "mem"):
        hbm = pk["hbm"]
        n, h, w, c = 256, 56, 56, 256
        e = n * h * w * c   # 205.5 M elements: y of config 2
        for dtype in (torch.float32, torch.float64):
            s = 4 if dtype == torch.float32 else 8
            ee = e if s == 4 else e // 2
            cc = c if s == 4 else c // 2
            x = torch.rand(ee, device=dev, dtype=dtype) * 2 - 1
            y = torch.empty_like(x)
            g_ = torch.rand(ee, device=dev, dtype=dtype) * 2 - 1
            dx = torch.empty_like(x)
            dn = str(dtype).split(".")[-1]

            def rec(name, t, bytes_):
                gbs = bytes_ / (t * 1e-3) / 1e9
                out({"op": name, "dtype": dn, "elems": ee, "ms": round(t, 4), "algorithmic_bytes": bytes_,
                     "gbs": round(gbs, 1), "hbm_peak_gbs": hbm, "frac": round(gbs / hbm, 3)})
            for kind in ("relu", "leaky_relu", "elu", "swish", "sigmoid", "tanh", "softplus"):
                a = pkg.ACT[kind]
                t = timed(stream, lambda: ctx.activation_forward(a, 0.1, n, ee // n, x, y), args.reps)
                rec("%s forward" % kind, t, 2 * s * ee)
                t = timed(stream, lambda: ctx.activation_backward(a, 0.1, n, ee // n, x, y, g_, dx), args.reps)
                streams = {"relu": 3, "leaky_relu": 3, "elu": 4, "swish": 3, "sigmoid": 3, "tanh": 3, "softplus": 3}[kind]
                rec("%s backward" % kind, t, streams * s * ee)
            # softmax over 10 classes, many rows
            rows = ee // 10
            t = timed(stream, lambda: ctx.activation_forward(pkg.ACT["softmax"], 1e-8, rows, 10, x, y), args.reps)
            rec("softmax forward (vol 10)", t, 2 * s * rows * 10)
            t = timed(stream, lambda: ctx.activation_backward(pkg.ACT["softmax"], 1e-8, rows, 10, x, y, g_, dx), args.reps)
            rec("softmax backward (vol 10)", t, 3 * s * rows * 10)
            # pooling 2x2 stride 2
            pg = pkg.PoolGeom(n, h, w, cc, 2, 2, 2, 2)
            eo = n * (h // 2) * (w // 2) * cc
            am = torch.empty(eo, device=dev, dtype=torch.uint8)
            for kind in ("max", "mean"):
                kk = pkg.POOL[kind]
                t = timed(stream, lambda: ctx.pool_forward(kk, pg, x, y, am), args.reps)
                rec("%s pool 2x2/2 forward" % kind, t, s * (ee + eo) + (eo if kind == "max" else 0))
                t = timed(stream, lambda: ctx.pool_backward(kk, pg, g_, am, dx), args.reps)
                rec("%s pool 2x2/2 backward" % kind, t, s * (eo + ee) + (eo if kind == "max" else 0))
            # batch normalisation, per channel, training
            gamma = torch.ones(cc, device=dev, dtype=dtype)
            beta = torch.zeros(cc, device=dev, dtype=dtype)
            rm, ri, sm, si = (torch.zeros(cc, device=dev, dtype=dtype) for _ in range(4))
            dg, dbt = torch.zeros(cc, device=dev, dtype=dtype), torch.zeros(cc, device=dev, dtype=dtype)
            t = timed(stream, lambda: ctx.batchnorm_forward(1, n, h, w, cc, 1, 1, 0.1, 1e-5, x, gamma, beta, rm, ri, sm, si, y),
                      args.reps)
            rec("batch-norm (per channel) training forward", t, 3 * s * ee)
            t = timed(stream, lambda: ctx.batchnorm_forward(1, n, h, w, cc, 0, 1, 0.1, 1e-5, x, gamma, beta, rm, ri, sm, si, y),
                      args.reps)
            rec("batch-norm (per channel) inference forward", t, 2 * s * ee)
            t = timed(stream, lambda: ctx.batchnorm_backward(1, n, h, w, cc, x, gamma, sm, si, g_, dg, dbt, dx), args.reps)
            rec("batch-norm (per channel) backward", t, 5 * s * ee)
            # residual add, optimizer step on a large arena
            t = timed(stream, lambda: ctx.add_inplace(ee, y, x), args.reps)
            rec("residual add (y += x)", t, 3 * s * ee)
            p_ = 64 * 1024 * 1024 if s == 4 else 32 * 1024 * 1024
            par = torch.randn(p_, device=dev, dtype=dtype)
            gr = torch.randn(p_, device=dev, dtype=dtype)
            m1, m2 = torch.zeros_like(par), torch.zeros_like(par)
            # streams: SURVEY 8d's count (read p, g, state; write p, state) + the write of the gradient reset
            for kind, streams in (("sgd", 4), ("momentum", 6), ("adam", 8), ("nadam", 8)):
                st = pkg.make_opt_step(pkg.OPT[kind], (1e-3, 0.1, 1e-3, 1e-5), 3, 0, 0.0, True, dtype=dn)
                t = timed(stream, lambda: ctx.optimizer_step(st, p_, par, gr, m1, m2), args.reps)
                d = {"op": "optimizer step %s (+ gradient reset)" % kind, "dtype": dn, "elems": p_, "ms": round(t, 4),
                     "algorithmic_bytes": streams * s * p_}
                d["gbs"] = round(d["algorithmic_bytes"] / (t * 1e-3) / 1e9, 1)
                d["hbm_peak_gbs"], d["frac"] = hbm, round(d["gbs"] / hbm, 3)
                out(d)
            del x, y, g_, dx, par, gr, m1, m2


if __name__ == "__main__":
    main()
