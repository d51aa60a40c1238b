"""Quick device timing of the config-2 convolution passes (development probe, not the bench)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
path = int(sys.argv[2]) if len(sys.argv) > 2 else 0
g = pkg.ConvGeom(N, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
ctx = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
ctx.set_conv_path(path)
x = torch.rand(N * 56 * 56 * 64, device="cuda") * 2 - 1
w = torch.randn(576 * 256, device="cuda") * 0.06
b = torch.zeros(256, device="cuda")
y = torch.empty(N * 56 * 56 * 256, device="cuda")
dy = torch.rand(N * 56 * 56 * 256, device="cuda") * 2 - 1
dx = torch.empty_like(x); dw = torch.zeros_like(w); db = torch.zeros_like(b)
flop = 2.0 * N * 56 * 56 * 576 * 256
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
t = timeit(lambda: ctx.conv_forward(g, x, w, b, y))
print("fwd  %.3f ms  %.1f TFLOP/s  path=%s" % (t, flop / t / 1e9, ctx.last_path))
t = timeit(lambda: ctx.conv_backward(g, x, w, dy, dw, db, dx))
print("bwd  %.3f ms  %.1f TFLOP/s (2 GEMMs) path=%s" % (t, 2 * flop / t / 1e9, ctx.last_path))
t2 = timeit(lambda: ctx.conv_backward(g, x, w, dy, dw, db, None))
print("bwd(no dx: wgrad+bgrad) %.3f ms -> dgrad(+split) %.3f ms" % (t2, t - t2))
print("launches", ctx.launches)
