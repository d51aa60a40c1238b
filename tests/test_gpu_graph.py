"""Step graphs and the LSTM gate kernel through the C ABI (include/cattl3_b200.h: cattl3_graph_*, cattl3_malloc / _free
under capture, cattl3_optimizer_step_indirect, cattl3_muladd).  A captured step must compute what the same calls
compute eagerly, on fresh inputs and with fresh optimizer scalars at every replay."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import cases as C  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


class Raw:
    """A cattl3_malloc block dressed as a tensor for the ctypes wrappers (dtype + data_ptr)."""

    def __init__(self, ptr, dtype):
        self.ptr, self.dtype = ptr, dtype

    def data_ptr(self):
        return self.ptr


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_muladd_matches_separately_rounded_expression(U, dt):
    """out = [out +] a * b [+ c * d] with every product and sum rounded on its own (no FMA contraction), as the
    reference's Eigen expressions are evaluated (LSTMNeuralNetwork.hpp:290-296): bit-exact against numpy."""
    c = U.ctx()
    rng = np.random.default_rng(61)
    for count in (1, 7, 1000, 70001):
        a, b, cc, d, o = (rng.uniform(-2, 2, count).astype(dt) for _ in range(5))
        ad, bd, cd, dd = U.dev(a), U.dev(b), U.dev(cc), U.dev(d)
        od = U.dev(o)
        c.muladd(count, False, ad, bd, None, None, od)
        assert np.array_equal(U.host(od, (count,)), a * b)
        c.muladd(count, False, ad, bd, cd, dd, od)
        assert np.array_equal(U.host(od, (count,)), a * b + cc * d)
        od = U.dev(o)
        c.muladd(count, True, ad, bd, cd, dd, od)
        assert np.array_equal(U.host(od, (count,)), o + (a * b + cc * d))
        c.muladd(count, False, od, bd, None, None, od)   # in place: state_grad *= forget_filter
        assert np.array_equal(U.host(od, (count,)), (o + (a * b + cc * d)) * b)
    with pytest.raises(U.pkg.Cattl3Error) as e:
        c.muladd(4, False, ad, bd, cd, None, od)
    assert e.value.code == U.pkg.ERR_INVALID


def _opt_step_bytes(U, kind, hyper, timestep, dt):
    st = U.pkg.make_opt_step(kind, hyper, timestep, 0, 0.0, True, np.dtype(dt).name)
    return st, bytes(st)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_captured_step_replays_like_eager_calls(U, dt):
    """conv forward -> ReLU -> conv backward -> fused Nadam update, captured once and replayed on three different inputs
    with three different sets of optimizer scalars, against the same calls issued eagerly.  The activations of the
    captured step come from cattl3_malloc (the graph's arena) and are released inside the capture."""
    import torch
    pkg = U.pkg
    torch.cuda.synchronize()
    ctx = pkg.Context(0, None)   # its own stream: the legacy default stream cannot be captured
    case = C.CONV_CASES["c2_small"]
    _, x0, w0, b0, dy0 = C.conv_inputs(case, dt, 71)
    g = pkg.ConvGeom(*case)
    tdt = torch.float32 if dt == np.float32 else torch.float64
    nx, ny, nw, nb = x0.size, dy0.size, w0.size, b0.size
    esize = np.dtype(dt).itemsize
    kind, hyper = C.OPT_CASES["nadam"]
    rng = np.random.default_rng(72)
    xs = [np.asfortranarray(rng.uniform(-1, 1, x0.shape).astype(dt)) for _ in range(3)]

    def make_state():
        p = torch.cat([U.dev(w0), U.dev(b0)])            # [W | b] as one parameter array
        return dict(p=p, g=torch.zeros_like(p), s1=torch.zeros_like(p), s2=torch.zeros_like(p), s3=torch.zeros_like(p))

    def step_calls(st, x, dyt, out, y_ptr, a_ptr, update):
        """The step: y = conv(x), a = relu(y), (dW, db) += conv_backward(x, dY = a - dyt-free surrogate), update."""
        W, B = st["p"][:nw], st["p"][nw:]
        ctx.conv_forward(g, x, W, B, y_ptr)
        ctx.activation_forward(pkg.ACT["relu"], 0.0, case[0], ny // case[0], y_ptr, a_ptr)
        ctx.conv_backward(g, x, W, a_ptr, st["g"][:nw], st["g"][nw:], None)
        ctx.scale(ny, 1.0, a_ptr, out)
        update(st)

    # eager run: three steps, scalars by value
    eager = make_state()
    xd = U.dev(xs[0])
    out_e = torch.empty(ny, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    a0 = ctx.allocated_bytes()
    outs_e = []
    for t in range(3):
        xd.copy_(U.dev(xs[t]))
        torch.cuda.synchronize()
        y_ptr, a_ptr = ctx.malloc(ny * esize), ctx.malloc(ny * esize)
        step, _ = _opt_step_bytes(U, kind, hyper, t, dt)
        step_calls(eager, xd, None, out_e, Raw(y_ptr, tdt), Raw(a_ptr, tdt),
                   lambda st: ctx.optimizer_step(step, nw + nb, st["p"], st["g"], st["s1"], st["s2"], st["s3"]))
        ctx.free(y_ptr)
        ctx.free(a_ptr)
        ctx.synchronize()
        outs_e.append(out_e.cpu().numpy().copy())
        if t == 0:
            per_step = ctx.allocated_bytes() - a0
    assert per_step >= 2 * ny * esize

    # captured run: same start, scalars through device memory
    cap = make_state()
    out_c = torch.empty(ny, dtype=tdt, device="cuda")
    dev_step = torch.zeros(ctypes.sizeof(pkg.OptStep), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    with pytest.raises(pkg.Cattl3Error):
        ctx.graph_begin(0)
    ctx.graph_begin(per_step + 4096)
    y_ptr, a_ptr = ctx.malloc(ny * esize), ctx.malloc(ny * esize)
    step_calls(cap, xd, None, out_c, Raw(y_ptr, tdt), Raw(a_ptr, tdt),
               lambda st: ctx.optimizer_step_indirect(kind, dev_step, nw + nb, st["p"], st["g"], st["s1"], st["s2"], st["s3"]))
    ctx.free(y_ptr)
    again = ctx.malloc(ny * esize)
    assert again == y_ptr, "a block released inside the capture is recycled"
    with pytest.raises(pkg.Cattl3Error) as e:
        ctx.malloc(1 << 30)                               # more than the arena holds
    assert e.value.code == pkg.ERR_UNSUPPORTED
    ctx.free(again)
    ctx.free(a_ptr)
    graph = ctx.graph_end()
    ctx.synchronize()
    assert float(cap["p"].sub(torch.cat([U.dev(w0), U.dev(b0)])).abs().max()) == 0.0, "capturing must not execute anything"
    for t in range(3):
        xd.copy_(U.dev(xs[t]))
        _, raw = _opt_step_bytes(U, kind, hyper, t, dt)
        dev_step.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        torch.cuda.synchronize()
        ctx.graph_launch(graph)
        ctx.synchronize()
        assert np.array_equal(out_c.cpu().numpy(), outs_e[t]), t
    ctx.synchronize()
    assert torch.equal(cap["p"], eager["p"]), "three replayed updates equal three eager updates bit for bit"
    assert float(cap["g"].abs().max()) == 0.0
    ctx.free(y_ptr)                                       # arena memory after the capture: a no-op
    ctx.graph_destroy(graph)
    ctx.close()


def test_graph_launch_refuses_a_graph_whose_scratch_moved(U):
    """Library scratch (here: the repacked weights of the tcgen05 path) is addressed by the graph; when a larger problem
    makes it grow, cattl3_graph_launch must refuse the older graph instead of replaying onto freed memory."""
    import torch
    pkg = U.pkg
    torch.cuda.synchronize()
    ctx = pkg.Context(0, None)
    small, big = C.CONV_CASES["c2_small"], C.CONV_CASES["c2_small_f256"]
    _, x, w, b, dy = C.conv_inputs(small, np.float32, 73)
    g = pkg.ConvGeom(*small)
    xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
    yd = torch.empty(dy.size, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    ctx.conv_forward(g, xd, wd, bd, yd)                   # eager first: scratch gets its size
    ctx.synchronize()
    assert ctx.last_path == "tcgen05"
    ctx.graph_begin(4096)
    ctx.conv_forward(g, xd, wd, bd, yd)
    graph = ctx.graph_end()
    ctx.graph_launch(graph)
    ctx.synchronize()
    _, x2, w2, b2, dy2 = C.conv_inputs(big, np.float32, 74)
    g2 = pkg.ConvGeom(*big)
    y2 = torch.empty(dy2.size, dtype=torch.float32, device="cuda")
    x2d, w2d, b2d = U.dev(x2), U.dev(w2), U.dev(b2)
    torch.cuda.synchronize()
    ctx.conv_forward(g2, x2d, w2d, b2d, y2)               # four times the filters: the packed-weight scratch grows
    ctx.synchronize()
    with pytest.raises(pkg.Cattl3Error) as e:
        ctx.graph_launch(graph)
    assert e.value.code == pkg.ERR_UNSUPPORTED
    ctx.graph_destroy(graph)
    ctx.close()
