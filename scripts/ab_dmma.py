#!/usr/bin/env python
"""A/B of the double kernel layers: the producer-warp DMMA kernels (conv_dmma.cu, version 2) against the
barrier-per-k-block ones (CATTL3_DMMA_V1=1) on the same inputs -- first the results (a set of ragged / strided / dilated /
transposed shapes, then config 2 at full size), then each pass timed alone.  One JSON line per case.

  python scripts/ab_dmma.py [--quick] > gpurun_out/ab_dmma.jsonl
"""
import argparse
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def load_package():
    spec = importlib.util.spec_from_file_location("cattl3_b200", os.path.join(ROOT, "c-attl3_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["cattl3_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def out_dims(geom, transposed):
    n, h, w, c, f, rh, rw, ph, pw, sh, sw, dh, dw = geom
    if transposed:
        return (h - 1) * sh + rh + (rh - 1) * dh - 2 * ph, (w - 1) * sw + rw + (rw - 1) * dw - 2 * pw
    return (h - rh - (rh - 1) * dh + 2 * ph) // sh + 1, (w - rw - (rw - 1) * dw + 2 * pw) // sw + 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="results only, no timing")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--config2", action="store_true", help="config 2 only (the ncu capture runs this with --quick)")
    args = ap.parse_args()
    pkg = load_package()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    ctx = pkg.Context(0, stream.cuda_stream)
    dt = torch.float64
    cases = [  # (name, geom, transposed, timed)
        ("ragged 3x3", (64, 29, 27, 36, 40, 3, 3, 1, 1, 1, 1, 0, 0), False, False),
        ("ragged 3x3, 9 channels -> 96", (34, 30, 28, 9, 96, 3, 3, 1, 1, 1, 1, 0, 0), False, False),
        ("odd batch", (65, 28, 28, 40, 72, 3, 3, 1, 1, 1, 1, 0, 0), False, False),
        ("3x3 stride 2", (32, 57, 55, 40, 130, 3, 3, 1, 1, 2, 2, 0, 0), False, False),
        ("3x3 dilated", (64, 30, 30, 48, 64, 3, 3, 2, 2, 1, 1, 1, 1), False, False),
        ("5x5 pad 2", (64, 28, 28, 9, 48, 5, 5, 2, 2, 1, 1, 0, 0), False, False),
        ("1x1", (64, 28, 28, 64, 256, 1, 1, 0, 0, 1, 1, 0, 0), False, False),
        ("transposed 3x3 stride 2", (128, 20, 20, 40, 40, 3, 3, 1, 1, 2, 2, 0, 0), True, False),
        ("transposed 2x2 stride 1", (64, 28, 28, 40, 40, 2, 2, 0, 0, 1, 1, 0, 0), True, False),
        ("config 2 (256 x 56 x 56 x 64 -> 256, 3x3)", (256, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0), False, True),
    ]
    if args.config2:
        cases = cases[-1:]
    worst = 0.0
    for name, geom, transposed, timed in cases:
        g = pkg.ConvGeom(*geom)
        n, h, w, c, f, rh, rw = geom[:7]
        oh, ow = out_dims(geom, transposed)
        m_in, m_out = n * h * w, n * oh * ow
        k = rh * rw * (f if transposed else c)
        kf = rh * rw * c * f
        gen = torch.Generator(device=dev).manual_seed(1234)
        x = torch.rand(m_in * c, device=dev, generator=gen, dtype=dt) * 2 - 1
        dy = torch.rand(m_out * f, device=dev, generator=gen, dtype=dt) * 2 - 1
        wt = torch.randn(kf, device=dev, generator=gen, dtype=dt) * (2.0 / k) ** 0.5
        nb = (oh * ow * f) if transposed else f
        b = torch.rand(nb, device=dev, generator=gen, dtype=dt) - 0.5
        res, paths, times = {}, {}, {}
        for ver in ("v1", "v2"):
            if ver == "v1":
                os.environ["CATTL3_DMMA_V1"] = "1"
            else:
                os.environ.pop("CATTL3_DMMA_V1", None)
            y = torch.empty(m_out * f, device=dev, dtype=dt)
            dx = torch.empty(m_in * c, device=dev, dtype=dt)
            dwt = torch.full((kf,), 0.25, device=dev, dtype=dt)   # the gradient accumulates on top of what is there
            db = torch.zeros(nb, device=dev, dtype=dt)
            ctx.conv_forward(g, x, wt, b, y, transposed=transposed)
            pf = ctx.last_path
            ctx.conv_backward(g, x, wt, dy, dwt, db, dx, transposed=transposed)
            torch.cuda.synchronize()
            res[ver] = (y, dx, dwt)
            paths[ver] = pf
            if timed and not args.quick:
                def t(fn):
                    fn()
                    torch.cuda.synchronize()
                    a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    for _ in range(args.reps):
                        fn()
                    bb.record(stream)
                    torch.cuda.synchronize()
                    return round(a.elapsed_time(bb) / args.reps, 3)
                y2, dx2, dw2, db2 = torch.empty_like(y), torch.empty_like(dx), torch.zeros_like(dwt), torch.zeros_like(db)
                times[ver] = {"forward_ms": t(lambda: ctx.conv_forward(g, x, wt, b, y2, transposed=transposed)),
                              "wgrad_ms": t(lambda: ctx.conv_backward(g, x, wt, dy, dw2, db2, None, transposed=transposed)),
                              "dgrad_ms": t(lambda: ctx.conv_backward(g, x, wt, dy, None, None, dx2, transposed=transposed))}
                del y2, dx2, dw2, db2
        rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
        errs = {key: rel(res["v2"][i], res["v1"][i]) for i, key in enumerate(("y", "dx", "dw"))}
        worst = max(worst, *errs.values())
        line = {"case": name, "transposed": transposed, "path": paths, "v2_vs_v1_rel": errs}
        if times:
            flop = 2.0 * m_out * k * f
            line["ms"] = times
            line["tflops_v2"] = {p: round(flop / (ms * 1e-3) / 1e12, 2) for p, ms in times["v2"].items()}
            line["tflops_v1"] = {p: round(flop / (ms * 1e-3) / 1e12, 2) for p, ms in times["v1"].items()}
        print(json.dumps(line), flush=True)
        del x, dy, wt, res
        torch.cuda.empty_cache()
    print(json.dumps({"worst_rel": worst, "ok": worst < 1e-12}), flush=True)
    sys.exit(0 if worst < 1e-12 else 1)


if __name__ == "__main__":
    main()
