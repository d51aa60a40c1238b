"""Development timings: config-2 double (DFMA path) and config-1 training through the C++ batch loop."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import load_package
import cases as C
from oracle import binding
pkg = load_package()
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if "double" in sys.argv:
    N = 256
    g = pkg.ConvGeom(N, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    dt = torch.float64
    x = torch.rand(N * 56 * 56 * 64, device="cuda", dtype=dt) * 2 - 1
    w = torch.randn(576 * 256, device="cuda", dtype=dt) * 0.06
    b = torch.zeros(256, device="cuda", dtype=dt)
    y = torch.empty(N * 56 * 56 * 256, device="cuda", dtype=dt)
    dy = torch.rand(N * 56 * 56 * 256, device="cuda", dtype=dt) * 2 - 1
    dx = torch.empty_like(x); dw = torch.zeros_like(w); db = torch.zeros_like(b)
    flop = 2.0 * N * 56 * 56 * 576 * 256
    t = timeit(lambda: ctx.conv_forward(g, x, w, b, y))
    print("f64 fwd  %.2f ms  %.1f TFLOP/s" % (t, flop / t / 1e9))
    t2 = timeit(lambda: ctx.conv_backward(g, x, w, dy, dw, db, None))
    print("f64 wgrad+bgrad %.2f ms  %.1f TFLOP/s" % (t2, flop / t2 / 1e9))
    t3 = timeit(lambda: ctx.conv_backward(g, x, w, dy, dw, db, dx))
    print("f64 dgrad %.2f ms  %.1f TFLOP/s" % (t3 - t2, flop / (t3 - t2) / 1e9))
if "cifar" in sys.argv:
    shim = binding.Oracle("ref", path=os.path.join(ROOT, "tests", "cpp", "_build", "libcattle_b200_shim.so"))
    ref = binding.Oracle("ref") if binding.have_ref() else None
    for dt in (np.float32,):
        rng = np.random.default_rng(1001)
        n = 1280
        x = C.rand(rng, (n, 32, 32, 3), dt)
        obj = np.zeros((n, 1, 1, 10), dtype=dt, order="F"); obj[np.arange(n), 0, 0, np.arange(n) % 10] = 1
        shim.train_cifar(x, obj, 64, 1)
        l0 = ctx.launches
        p, loss, ms = shim.train_cifar(x, obj, 64, 1)
        print("config 1 b200: %d samples, batch 64, 1 epoch: %.1f ms -> %.0f samples/s (loss %.4f)" % (n, ms, n / ms * 1e3, loss))
        if ref:
            p, loss, ms = ref.train_cifar(x, obj, 64, 1)
            print("config 1 reference CPU (%d threads): %.1f ms -> %.0f samples/s (loss %.4f)" % (ref.num_threads(), ms, n / ms * 1e3, loss))
