"""-m gpu: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs, and
against the committed golden vectors of the unmodified reference.  Tolerances are the north star's:
1e-4 relative (float), 1e-10 (double), measured norm-relative (cases.relerr)."""
import os

import numpy as np
import pytest

import cases as C
from oracle.binding import Geom, conv_out_dims

pytestmark = pytest.mark.gpu

DTYPES = [np.float32, np.float64]


def _conv_gpu(U, case, x, w, b, dy, transposed, want_dx=True, reps=1, path=None):
    c = U.ctx(path)
    g = U.pkg.ConvGeom(*case)
    og = Geom(*case)
    oh, ow = conv_out_dims(og, transposed)
    dt = x.dtype
    xd, wd, bd, dyd = U.dev(x), U.dev(w), U.dev(b), U.dev(dy)
    yd = U.zeros((og.n, oh, ow, og.f), dt)
    dxd = U.zeros(x.shape, dt) if want_dx else None
    dwd, dbd = U.zeros(w.shape, dt), U.zeros(b.shape, dt)
    c.conv_forward(g, xd, wd, bd, yd, transposed)
    fpath = c.last_path
    for _ in range(reps):
        c.conv_backward(g, xd, wd, dyd, dwd, dbd, dxd, transposed)
    c.synchronize()
    return dict(y=U.host(yd, (og.n, oh, ow, og.f)), dx=U.host(dxd, x.shape) if want_dx else None,
                dw=U.host(dwd, w.shape), db=U.host(dbd, b.shape), path=fpath)


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("tr", [False, True])
def test_conv_simt_vs_oracle(U, orc, dt, tr):
    table = C.TCONV_CASES if tr else C.CONV_CASES
    for name, case in table.items():
        g, x, w, b, dy = C.conv_inputs(case, dt, 21, tr)
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        a = _conv_gpu(U, case, x, w, b, dy, tr, reps=2, path=U.pkg.PATH_SIMT)
        assert a["path"] == "simt"
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(dt)], (name, k, C.relerr(a[k], r[k]))


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_golden(U, golden, dt):
    suf = "f32" if dt == np.float32 else "f64"
    for tr, fam, table in ((False, "conv", C.CONV_CASES), (True, "tconv", C.TCONV_CASES)):
        names = sorted({k.split("/")[1] for k in golden.files if k.startswith(fam + "/")})
        for name in names:
            k = "%s/%s/%s/" % (fam, name, suf)
            x, w, b, dy = (np.asfortranarray(golden[k + s]) for s in ("x", "w", "b", "dy"))
            a = _conv_gpu(U, table[name], x, w, b, dy, tr, reps=2)
            for out in ("y", "dx", "dw", "db"):
                assert C.relerr(a[out], golden[k + out]) < C.TOL[np.dtype(dt)], (name, out)


@pytest.mark.parametrize("dt", DTYPES)
def test_conv_input_layer_and_accumulation(U, orc, dt):
    """dx == NULL skips the input gradient; a second pass_back doubles dW/db (beta = 1)."""
    case = C.CONV_CASES["gt_rank3"]
    g, x, w, b, dy = C.conv_inputs(case, dt, 22)
    r1 = orc.conv(g, x, w, b, dy, back_reps=1)
    a1 = _conv_gpu(U, case, x, w, b, dy, False, want_dx=False, reps=1)
    a3 = _conv_gpu(U, case, x, w, b, dy, False, want_dx=False, reps=3)
    assert a1["dx"] is None
    assert C.relerr(a1["dw"], r1["dw"]) < C.TOL[np.dtype(dt)]
    assert C.relerr(a3["dw"], 3 * r1["dw"]) < C.TOL[np.dtype(dt)]
    assert C.relerr(a3["db"], 3 * r1["db"]) < C.TOL[np.dtype(dt)]


@pytest.mark.parametrize("dt", DTYPES)
def test_dense_vs_oracle(U, orc, dt):
    rng = np.random.default_rng(23)
    c = U.ctx()
    for name, (n, i, o) in C.DENSE_CASES.items():
        x, w, b, dy = C.rand(rng, (n, i), dt), C.rand(rng, (i, o), dt), C.rand(rng, (1, o), dt), C.rand(rng, (n, o), dt)
        r = orc.dense(x, w, b, dy)
        xd, wd, bd, dyd = U.dev(x), U.dev(w), U.dev(b), U.dev(dy)
        yd, dxd, dwd, dbd = U.zeros((n, o), dt), U.zeros((n, i), dt), U.zeros((i, o), dt), U.zeros((1, o), dt)
        c.dense_forward(n, i, o, xd, wd, bd, yd)
        if name.startswith("head"):
            assert c.last_path == "skinny", (name, c.last_path)   # split-K forward of a classifier head
        c.dense_backward(n, i, o, xd, wd, dyd, dwd, dbd, dxd)
        c.synchronize()
        got = dict(y=U.host(yd, (n, o)), dx=U.host(dxd, (n, i)), dw=U.host(dwd, (i, o)), db=U.host(dbd, (1, o)))
        for k in got:
            assert C.relerr(got[k], r[k]) < C.TOL[np.dtype(dt)], (name, k)


@pytest.mark.parametrize("dt", DTYPES)
def test_activation_vs_golden_and_oracle(U, orc, golden, dt):
    suf = "f32" if dt == np.float32 else "f64"
    c = U.ctx()
    rng = np.random.default_rng(24)
    for name, (kind, alpha) in C.ACT_CASES.items():
        k = "act/%s/%s/" % (name, suf)
        sets = [(np.asfortranarray(golden[k + "x"]), np.asfortranarray(golden[k + "dy"]), golden[k + "y"], golden[k + "dx"])]
        xb = C.rand(rng, (33, 7, 5, 3), dt, -3, 3)  # odd element count exercises the scalar tail
        dyb = C.rand(rng, xb.shape, dt)
        rb = orc.activation(kind, alpha, xb, dyb)
        sets.append((xb, dyb, rb["y"], rb["dx"]))
        al = 1e-5 if name == "softmax" else alpha  # softmax: alpha carries epsilon (NumericUtils EPSILON2)
        for x, dy, ry, rdx in sets:
            rows, vol = x.shape[0], x.size // x.shape[0]
            xd, dyd = U.dev(x), U.dev(dy)
            yd, dxd = U.zeros(x.shape, dt), U.zeros(x.shape, dt)
            c.activation_forward(kind, al, rows, vol, xd, yd)
            c.activation_backward(kind, al, rows, vol, xd, yd, dyd, dxd)
            c.synchronize()
            tol = C.TOL[np.dtype(dt)]
            assert C.relerr(U.host(yd, x.shape), ry) < tol, name
            assert C.relerr(U.host(dxd, x.shape), rdx) < tol, name


@pytest.mark.parametrize("dt", DTYPES)
def test_pool_vs_oracle(U, orc, dt):
    import torch
    c = U.ctx()
    rng = np.random.default_rng(25)
    for name, (kind, n, h, w, ch, rh, rw, sh, sw) in C.POOL_CASES.items():
        x = np.asfortranarray(np.round(C.rand(rng, (n, h, w, ch), dt) * 4) / 4)  # ties on purpose
        oh, ow = (h - rh) // sh + 1, (w - rw) // sw + 1
        dy = C.rand(rng, (n, oh, ow, ch), dt)
        r = orc.pool(kind, x, rh, rw, sh, sw, dy)
        g = U.pkg.PoolGeom(n, h, w, ch, rh, rw, sh, sw)
        xd, dyd = U.dev(x), U.dev(dy)
        yd, dxd = U.zeros(dy.shape, dt), U.zeros(x.shape, dt)
        am = torch.zeros(dy.size, dtype=torch.uint8, device="cuda")
        c.pool_forward(kind, g, xd, yd, am)
        c.pool_backward(kind, g, dyd, am, dxd)
        c.synchronize()
        if kind == 0:  # selection / routing only: bit-exact
            assert np.array_equal(U.host(yd, dy.shape), r["y"]), name
        assert C.relerr(U.host(yd, dy.shape), r["y"]) < C.TOL[np.dtype(dt)], name
        assert C.relerr(U.host(dxd, x.shape), r["dx"]) < C.TOL[np.dtype(dt)], name


@pytest.mark.parametrize("dt", DTYPES)
def test_batchnorm_vs_oracle(U, orc, dt):
    c = U.ctx()
    rng = np.random.default_rng(26)
    for name, (pc, n, h, w, ch, steps) in C.BN_CASES.items():
        xs = [C.rand(rng, (n, h, w, ch), dt, -1, 2) for _ in range(steps)]
        G = ch if pc else h * w * ch
        gm, bt, dy = C.rand(rng, (G,), dt, 0.5, 1.5), C.rand(rng, (G,), dt), C.rand(rng, (n, h, w, ch), dt)
        r = orc.batchnorm(pc, xs, gm, bt, dy)
        gmd, btd, dyd = U.dev(gm), U.dev(bt), U.dev(dy)
        rm, rs, sm, ss = (U.zeros((G,), dt) for _ in range(4))
        yd, yi, dxd = U.zeros(dy.shape, dt), U.zeros(dy.shape, dt), U.zeros(dy.shape, dt)
        dg, db = U.zeros((G,), dt), U.zeros((G,), dt)
        for s, x in enumerate(xs):
            xd = U.dev(x)
            c.batchnorm_forward(pc, n, h, w, ch, 1, s > 0, 0.1, 1e-5, xd, gmd, btd, rm, rs, sm, ss, yd)
        c.batchnorm_backward(pc, n, h, w, ch, xd, gmd, sm, ss, dyd, dg, db, dxd)
        c.batchnorm_forward(pc, n, h, w, ch, 0, 1, 0.1, 1e-5, xd, gmd, btd, rm, rs, None, None, yi)
        c.synchronize()
        got = dict(y=U.host(yd, dy.shape), dx=U.host(dxd, dy.shape), dgamma=U.host(dg, (G,)), dbeta=U.host(db, (G,)),
                   run_mean=U.host(rm, (G,)), run_inv_sd=U.host(rs, (G,)), y_infer=U.host(yi, dy.shape))
        tol = C.TOL[np.dtype(dt)]
        for k in got:
            assert C.relerr(got[k], r[k]) < tol, (name, k, C.relerr(got[k], r[k]))


@pytest.mark.parametrize("dt", DTYPES)
def test_optimizer_vs_golden(U, golden, dt):
    suf = "f32" if dt == np.float32 else "f64"
    c = U.ctx()
    for name, (kind, hy) in C.OPT_CASES.items():
        for lam in (0.0, 0.01):
            k = "opt/%s_l2_%g/%s/" % (name, lam, suf)
            p0, grads = np.asfortranarray(golden[k + "p0"]), golden[k + "grads"]
            pd = U.dev(p0)
            s1, s2, s3 = (U.zeros(p0.shape, dt) for _ in range(3))
            for t, gr in enumerate(grads):
                gd = U.dev(np.asfortranarray(gr))
                st = U.pkg.make_opt_step(kind, hy, t, t // 3, lam, True, np.dtype(dt).name)
                c.optimizer_step(st, p0.size, pd, gd, s1, s2, s3)
                assert float(gd.abs().max()) == 0.0  # reset_grad
            c.synchronize()
            tol = 2e-6 if dt == np.float32 else 1e-12
            assert C.relerr(U.host(pd, p0.shape), golden[k + "p"]) < tol, (name, lam)


@pytest.mark.parametrize("dt", DTYPES)
def test_regularize_kernel(U, dt):
    """cattl3_regularize against the reference's formulas (L1 / L2 / ElasticNet ParameterRegularization.hpp:
    d_function = (v >= 0 ? l1 : -l1) + l2 v, function = l1 |v|_1 + l2 / 2 |v|^2), the penalty accumulating across
    calls in a device double; sizes from one block to many, a zero among the values (sign(0) = +1 in the reference)."""
    import torch
    c = U.ctx()
    rng = np.random.default_rng(77)
    penalty = torch.zeros(1, dtype=torch.float64, device="cuda")
    expect = 0.0
    for count, (l1, l2) in ((5, (0.01, 0.0)), (1000, (0.0, 0.02)), (300001, (0.003, 0.01))):
        v = rng.uniform(-1, 1, count).astype(dt)
        v[0] = 0
        g = rng.uniform(-1, 1, count).astype(dt)
        vd, gd = U.dev(v), U.dev(g)
        c.regularize(count, l1, l2, vd, gd, penalty)
        c.synchronize()
        want = g + (np.where(v >= 0, dt(l1), dt(-l1)) + v * dt(l2))
        assert C.relerr(U.host(gd, (count,)), want) < (1e-6 if dt == np.float32 else 1e-15)
        expect += float(dt(l1)) * np.abs(v.astype(np.float64)).sum() + 0.5 * float(dt(l2)) * (v.astype(np.float64) ** 2).sum()
        assert abs(float(penalty.cpu()[0]) - expect) < 1e-12 * max(1.0, expect), (count, float(penalty.cpu()[0]), expect)
        # penalty only: the gradient is left alone
        c.regularize(count, l1, l2, vd, None, penalty)
        c.synchronize()
        expect += float(dt(l1)) * np.abs(v.astype(np.float64)).sum() + 0.5 * float(dt(l2)) * (v.astype(np.float64) ** 2).sum()
        assert C.relerr(U.host(gd, (count,)), want) < (1e-6 if dt == np.float32 else 1e-15)
        assert abs(float(penalty.cpu()[0]) - expect) < 1e-12 * max(1.0, expect)


# "stem": 3 input channels, zero-padded to a 16-channel k-block by TMA out-of-bounds fill (forward on tcgen05)
TC_CASES = ["c2_small", "c2_small_f256", "c2_stride2", "c2_dil1", "c2_1x1", "ragged_c", "stem", "rows64", "rows_f256", "rows_2x3"]


def test_conv_tcgen05_vs_oracle(U, orc):
    """The TMA + tcgen05 3xTF32 implicit GEMM (forward and stride-1 input gradient) against the
    oracle at the float tolerance; also checks that the tensor-core path is the one that ran."""
    for name in TC_CASES:
        case = C.CONV_CASES[name]
        g, x, w, b, dy = C.conv_inputs(case, np.float32, 31)
        r = orc.conv(g, x, w, b, dy)
        a = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_AUTO)
        assert a["path"] == "tcgen05", name
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(np.float32)], (name, k, C.relerr(a[k], r[k]))


def test_conv_tcgen05_matches_simt_tightly(U):
    """3xTF32 must sit near FP32-GEMM accuracy (1xTF32 would be ~3e-4): tcgen05 vs the FFMA kernel.
    The residual (~1e-5 at K=2304) is the tensor core's round-toward-zero accumulation, ~2^-24 per
    MMA step ("accuracy of the 3xTF32 path")."""
    case = C.CONV_CASES["c2_small_f256"]
    g, x, w, b, dy = C.conv_inputs(case, np.float32, 32)
    a = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_AUTO)
    s = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_SIMT)
    assert a["path"] == "tcgen05" and s["path"] == "simt"
    assert C.relerr(a["y"], s["y"]) < 1e-5
    assert C.relerr(a["dx"], s["dx"]) < 3e-5
    assert C.relerr(a["dw"], s["dw"]) < 1e-5


def test_unaligned_tensors_fall_back_instead_of_failing(U, orc):
    """A tensor that starts 4 bytes into a larger buffer cannot be described to TMA (16-byte alignment): the layer must run
    on the FMA / SIMT kernels, not fail (a C-ABI caller passing a slice)."""
    import torch
    case = C.CONV_CASES["c2_small"]
    g, x, w, b, dy = C.conv_inputs(case, np.float32, 34)
    r = orc.conv(g, x, w, b, dy)
    c = U.ctx()
    cg = U.pkg.ConvGeom(*case)
    big = torch.zeros(x.size + 1, device="cuda")
    xd = big[1:]
    xd.copy_(U.dev(x))
    assert xd.data_ptr() % 16 == 4
    wd, bd, dyd = U.dev(w), U.dev(b), U.dev(dy)
    yd, dxd = U.zeros(dy.shape, np.float32), U.zeros(x.shape, np.float32)
    dwd, dbd = U.zeros(w.shape, np.float32), U.zeros(b.shape, np.float32)
    c.conv_forward(cg, xd, wd, bd, yd)
    assert c.last_path != "tcgen05"
    c.conv_backward(cg, xd, wd, dyd, dwd, dbd, dxd)
    c.synchronize()
    tol = C.TOL[np.dtype(np.float32)]
    assert C.relerr(U.host(yd, dy.shape), r["y"]) < tol
    assert C.relerr(U.host(dwd, w.shape), r["dw"]) < tol
    assert C.relerr(U.host(dxd, x.shape), r["dx"]) < tol


def test_tcgen05_path_refuses_unsupported_shapes(U):
    case = C.CONV_CASES["gt_rank3"]  # batch 5: no 32-row TMA boxes
    g, x, w, b, dy = C.conv_inputs(case, np.float32, 33)
    with pytest.raises(U.pkg.Cattl3Error) as e:
        _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_TCGEN05)
    assert e.value.code == U.pkg.ERR_UNSUPPORTED


def test_config2_full_size_properties(U):
    """BASELINE.json configs[1] at full size (N=256, 56x56x64 -> 256, 3x3 pad 1, float): the oracle is
    too slow here, so parity is checked through size-independent properties:
      * the tcgen05 path against the FFMA path (two independent kernels, different summation orders);
      * adjointness: <conv(x) - b, dY> == <x, dX> == <W, dW> (forward, input- and weight-gradient are
        the three faces of one bilinear form);
      * db == column sums of dY."""
    import torch
    pkg = U.pkg
    N = 256
    g = pkg.ConvGeom(N, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    gen = torch.Generator(device="cuda").manual_seed(2001)
    x = torch.rand(N * 56 * 56 * 64, device="cuda", generator=gen) * 2 - 1
    dy = torch.rand(N * 56 * 56 * 256, device="cuda", generator=gen) * 2 - 1
    w = torch.randn(576 * 256, device="cuda", generator=gen) * (2.0 / 576) ** 0.5
    b = torch.rand(256, device="cuda", generator=gen)
    res = {}
    for name, path in (("tc", pkg.PATH_AUTO), ("simt", pkg.PATH_SIMT)):
        c = U.ctx(path)
        y = torch.empty_like(dy)
        dx = torch.empty_like(x)
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        c.conv_forward(g, x, w, b, y)
        assert c.last_path == ("tcgen05" if name == "tc" else "simt")
        c.conv_backward(g, x, w, dy, dw, db, dx)
        c.synchronize()
        res[name] = (y, dx, dw, db)
    rel = lambda a, r: float((a - r).abs().max() / r.abs().max())
    errs = {k: rel(res["tc"][i], res["simt"][i]) for i, k in enumerate(("y", "dx", "dw", "db"))}
    print("config-2 full size, tcgen05 vs FFMA:", errs)
    for k, e in errs.items():
        assert e < 1e-4, (k, e)
    y, dx, dw, db = res["tc"]
    M = N * 56 * 56
    yb = (y.view(256, M) - b.view(256, 1)).double().flatten()
    lhs = float(torch.dot(yb, dy.double()))
    mid = float(torch.dot(x.double(), dx.double()))
    rhs = float(torch.dot(w.double(), dw.double()))
    scale = float(yb.norm() * dy.double().norm())
    print("adjointness: <y-b,dY>=%.6e <x,dX>=%.6e <W,dW>=%.6e (scale %.3e)" % (lhs, mid, rhs, scale))
    assert abs(lhs - mid) / scale < 1e-6 and abs(lhs - rhs) / scale < 1e-6
    colsum = dy.view(256, M).double().sum(dim=1)
    assert float((db.double() - colsum).abs().max() / colsum.abs().max()) < 1e-5


def _relerr_chunked(a, b):
    """cases.relerr without a float64 copy of a 200 M-element tensor: max|a-b| / max|b| over slabs of the flat arrays."""
    a, b = a.ravel(order="K"), b.ravel(order="K")
    num = den = 0.0
    step = 1 << 24
    for i in range(0, a.size, step):
        bb = b[i:i + step].astype(np.float64)
        num = max(num, float(np.max(np.abs(a[i:i + step].astype(np.float64) - bb))))
        den = max(den, float(np.max(np.abs(bb))))
    return num / max(den, 1e-30)


@pytest.mark.parametrize("dt,n", [(np.float32, 256), (np.float64, int(os.environ.get("CATTL3_FULLSIZE_N_F64", "256")))])
def test_config2_full_size_vs_unmodified_reference(U, ref, dt, n):
    """BASELINE.json configs[1] AT FULL SIZE against the unmodified reference on identical inputs: one
    ConvKernelLayer<S,3>({56,56,64}, 256) 3x3 pad 1, batch 256, pass_forward + pass_back
    (/root/reference/C-ATTL3/layer/kernel/ConvKernelLayer.hpp:115-189 through oracle/_ref), float on the tcgen05 3xTF32
    kernels and double on the FP64 tensor-core kernels, at the north-star tolerances (1e-4 / 1e-10 norm-relative).
    x uniform[-1,1) seed 2001, dY uniform[-1,1) seed 2002, He-scaled weights."""
    case = (n, 56, 56, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    og = Geom(*case)
    K = 576
    x = C.rand(np.random.default_rng(2001), (n, 56, 56, 64), dt)
    dy = C.rand(np.random.default_rng(2002), (n, 56, 56, 256), dt)
    rng = np.random.default_rng(2003)
    w = np.asfortranarray((rng.standard_normal((K, 256)) * (2.0 / K) ** 0.5).astype(dt))
    b = C.rand(rng, (1, 256), dt)
    a = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_AUTO)
    assert a["path"] == ("tcgen05" if dt == np.float32 else "dmma"), a["path"]
    r = ref.conv(og, x, w, b, dy)
    errs = {k: _relerr_chunked(a[k], r[k]) for k in ("y", "dx", "dw", "db")}
    print("config 2 full size (N=%d, %s) vs the unmodified reference on %d host threads (fwd %.0f ms, bwd %.0f ms): %s"
          % (n, np.dtype(dt).name, ref.num_threads(), r["times_ms"][0], r["times_ms"][1], errs))
    for k, e in errs.items():
        assert e < C.TOL[np.dtype(dt)], (k, e)


@pytest.mark.parametrize("chunks", ["1", "4", "strips2", "strips4"])
def test_conv_host_entry_points_vs_oracle(U, orc, chunks, monkeypatch):
    """cattl3_conv_forward_host_f32 / cattl3_conv_backward_host_f32 (+ the _async forms and cattl3_host_wait): host
    tensors in and out, the copies pipelined with the kernels over chunks of filters on three streams, against the
    oracle's ConvKernelLayer (ConvKernelLayer.hpp:115-189).  Two steps back to back: the second forward reuses the
    staging buffers while the first step's downloads are still running; gradients accumulate across the two calls."""
    import torch
    case = (32, 9, 8, 16, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    g, x, w, b, dy = C.conv_inputs(case, np.float32, 77, False)
    x2 = np.asfortranarray(x[::-1].copy())
    r1 = orc.conv(g, x, w, b, dy, back_reps=1)
    r2 = orc.conv(g, x2, w, b, dy, back_reps=1)
    c = U.ctx()
    cg = U.pkg.ConvGeom(*case)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory()
    xh, x2h, dyh = pin(x), pin(x2), pin(dy)
    yh = [torch.empty(dy.size).pin_memory() for _ in range(2)]
    dxh = [torch.empty(x.size).pin_memory() for _ in range(2)]
    wd, bd = U.dev(w), U.dev(b)
    dwd, dbd = U.zeros(w.shape, np.float32), U.zeros(b.shape, np.float32)
    xkeep = U.zeros(x.shape, np.float32)
    if chunks.startswith("strips"):   # strips of image columns (stride-1 layers on the tensor-core path), else chunks of filters
        monkeypatch.setenv("CATTL3_HOST_STRIPS", chunks[6:])
    else:
        monkeypatch.setenv("CATTL3_NO_HOST_STRIPS", "1")
        monkeypatch.setenv("CATTL3_HOST_CHUNKS", chunks)
    torch.cuda.synchronize()
    c.conv_forward_host_async(cg, xh, wd, bd, yh[0], xkeep)
    c.conv_backward_host_async(cg, xkeep, wd, dyh, dwd, dbd, dxh[0])
    c.conv_forward_host_async(cg, x2h, wd, bd, yh[1], xkeep)
    c.conv_backward_host_async(cg, xkeep, wd, dyh, dwd, dbd, dxh[1])
    c.host_wait()
    tol = C.TOL[np.dtype(np.float32)]
    for i, r in enumerate((r1, r2)):
        assert C.relerr(yh[i].numpy().reshape(dy.shape, order="F"), r["y"]) < tol, ("y", i)
        assert C.relerr(dxh[i].numpy().reshape(x.shape, order="F"), r["dx"]) < tol, ("dx", i)
    assert C.relerr(U.host(dwd, w.shape), r1["dw"] + r2["dw"]) < tol
    assert C.relerr(U.host(dbd, b.shape), r1["db"] + r2["db"]) < tol
    # the synchronous forms: complete on return
    y3 = torch.empty(dy.size).pin_memory()
    c.conv_forward_host(cg, xh, wd, bd, y3, None)
    assert C.relerr(y3.numpy().reshape(dy.shape, order="F"), r1["y"]) < tol


@pytest.mark.parametrize("dt", DTYPES)
def test_loss_kernels_vs_numpy(U, dt):
    """SquaredLoss / CrossEntropyLoss on the device (loss/SquaredLoss.hpp:25-34, CrossEntropyLoss.hpp:33-42):
    per-sample losses and the gradient divided by the batch size."""
    import torch
    c = U.ctx()
    rng = np.random.default_rng(61)
    for rows, vol in ((5, 10), (64, 784), (1, 1), (33, 7)):
        out = C.rand(rng, (rows, vol), dt, 0.01, 1.0)
        obj = C.rand(rng, (rows, vol), dt, 0.0, 1.0)
        od, td = U.dev(out), U.dev(obj)
        o64, t64 = out.astype(np.float64), obj.astype(np.float64)
        for kind, eps, want_l, want_g in (
                (0, 0.0, ((o64 - t64) ** 2).sum(1), 2 * (o64 - t64) / 32),
                (1, 1e-5, -(np.log(o64 + 1e-5) * t64).sum(1), -t64 / (o64 + 1e-5) / 32)):
            ld, gd = U.zeros((rows,), dt), U.zeros((rows, vol), dt)
            c.loss(kind, rows, vol, eps, 32.0, od, td, ld, gd)
            c.synchronize()
            tol = 1e-5 if dt == np.float32 else 1e-12
            assert C.relerr(U.host(ld, (rows,)), want_l) < tol, (kind, rows, vol)
            assert C.relerr(U.host(gd, (rows, vol)), want_g) < tol, (kind, rows, vol)


@pytest.mark.parametrize("dt", DTYPES)
def test_dropout_kernels(U, dt):
    """DropoutLayer.hpp:74-94 on the device: y = x * mask with mask in {0, 1 / (1 - p + eps)}, the drop rate is p,
    pass_back applies the same mask, a seed reproduces its mask and another seed gives another one."""
    import torch
    c = U.ctx()
    rng = np.random.default_rng(62)
    n, p, eps = 1 << 20, 0.25, 1e-5
    x = C.rand(rng, (n,), dt, 0.5, 1.5)
    dy = C.rand(rng, (n,), dt)
    xd, dyd = U.dev(x), U.dev(dy)
    yd, dxd = U.zeros((n,), dt), U.zeros((n,), dt)
    mask = torch.zeros(n, dtype=torch.uint8, device="cuda")
    c.dropout_forward(n, p, eps, 1234, xd, yd, mask)
    c.dropout_backward(n, p, eps, dyd, mask, dxd)
    c.synchronize()
    m = mask.cpu().numpy().astype(bool)
    scale = dt(1) / (dt(1) - dt(p) + dt(eps))
    y, dx = U.host(yd, (n,)), U.host(dxd, (n,))
    assert np.array_equal(y, np.where(m, x * scale, 0).astype(dt))
    assert np.array_equal(dx, np.where(m, dy * scale, 0).astype(dt))
    assert abs((~m).mean() - p) < 4 * np.sqrt(p * (1 - p) / n)          # drop rate within 4 sigma
    assert abs(np.corrcoef(m[:-1], m[1:])[0, 1]) < 0.01                   # neighbours are uncorrelated
    mask2, mask3 = torch.zeros_like(mask), torch.zeros_like(mask)
    c.dropout_forward(n, p, eps, 1234, xd, yd, mask2)
    c.dropout_forward(n, p, eps, 1235, xd, yd, mask3)
    c.synchronize()
    assert torch.equal(mask, mask2)
    assert 0.3 < (mask != mask3).float().mean().item() < 0.45            # 2 p (1 - p) = 0.375


DFMA_CASES = ["c2_small_f256", "c2_stride2", "c2_1x1", "dfma_odd", "dfma_mid", "dfma_wide"]


@pytest.mark.parametrize("tr", [False, True])
@pytest.mark.parametrize("path,name_of_path", [("AUTO", "dmma"), ("FMA", "dfma")])
def test_conv_double_gemm_paths_vs_oracle(U, orc, tr, path, name_of_path):
    """Double at GEMM-sized shapes against the oracle at 1e-10: the FP64 tensor-core kernels (conv_dmma.cu, what AUTO
    picks) and the big-tile DFMA kernels (conv_dfma.cu, PATH_FMA); checks which one ran."""
    table = C.TCONV_CASES if tr else C.CONV_CASES
    names = ["dfma_t"] if tr else DFMA_CASES
    sel = U.pkg.PATH_AUTO if path == "AUTO" else U.pkg.PATH_FMA
    for name in names:
        g, x, w, b, dy = C.conv_inputs(table[name], np.float64, 71, tr)
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        a = _conv_gpu(U, table[name], x, w, b, dy, tr, reps=2, path=sel)
        assert a["path"] == name_of_path, (name, a["path"])
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(np.float64)], (name, k, C.relerr(a[k], r[k]))


@pytest.mark.parametrize("tr", [False, True])
def test_conv_double_producer_warp_kernels_vs_oracle(U, orc, tr, capfd, monkeypatch):
    """The producer-warp DMMA kernels (conv_dmma.cu, version 2: cp.async ring + mbarriers, channel-outer reduction,
    swizzled weight-gradient tiles) against the oracle at 1e-10 on shapes large enough for them to be chosen; the
    launch trace on stderr says which kernel ran."""
    monkeypatch.setenv("CATTL3_DMMA_TRACE", "1")
    table = C.DMMA2_TCONV_CASES if tr else C.DMMA2_CASES
    seen = ""
    for name, case in table.items():
        g, x, w, b, dy = C.conv_inputs(case, np.float64, 73, tr)
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        capfd.readouterr()
        a = _conv_gpu(U, case, x, w, b, dy, tr, reps=2)
        err = capfd.readouterr().err
        assert a["path"] == "dmma" and "dmma2 gather" in err, (name, a["path"], err)
        assert "dmma2 wgrad" in err or case[0] % 2 == 1, (name, err)
        seen += err
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(np.float64)], (name, k, C.relerr(a[k], r[k]))
    if not tr:   # both gather tiles, 16- and 8-byte copies, all three weight-gradient tiles
        for what in ("gather tile 128 x 128 vec", "gather tile 256 x 64 vec", "gather tile 128 x 128\n", "wgrad tile 256 x 64",
                     "wgrad tile 128 x 128", "wgrad tile 64 x 256"):
            assert what in seen, (what, seen)


def test_conv_dfma_matches_simt_at_size(U):
    """A mid-size double convolution (N=64, 14x14x64 -> 256, M = 12544, K = 576): the DMMA and DFMA kernels against the
    any-shape SIMT kernels, three independent implementations with different summation orders."""
    case = (64, 14, 14, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0)
    g, x, w, b, dy = C.conv_inputs(case, np.float64, 72)
    a = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_AUTO)
    f = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_FMA)
    s = _conv_gpu(U, case, x, w, b, dy, False, path=U.pkg.PATH_SIMT)
    assert a["path"] == "dmma" and f["path"] == "dfma" and s["path"] == "simt"
    for k in ("y", "dx", "dw", "db"):
        assert C.relerr(a[k], s[k]) < 1e-13, (k, C.relerr(a[k], s[k]))
        assert C.relerr(f[k], s[k]) < 1e-13, (k, C.relerr(f[k], s[k]))


def test_conv_ffma_vs_oracle(U, orc):
    """Float layers the tensor-core path refuses (3 input channels: config 4's stem; ragged batches) run on the
    big-tile FFMA kernels (conv_dfma.cu, flattened reduction for few channels), against the oracle at 1e-4."""
    for name in ["stem", "stem_wide", "dfma_odd", "dfma_mid", "dfma_wide"]:
        g, x, w, b, dy = C.conv_inputs(C.CONV_CASES[name], np.float32, 73)
        r = orc.conv(g, x, w, b, dy, back_reps=2)
        a = _conv_gpu(U, C.CONV_CASES[name], x, w, b, dy, False, reps=2, path=U.pkg.PATH_FMA)
        assert a["path"] == "ffma", (name, a["path"])
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(np.float32)], (name, k, C.relerr(a[k], r[k]))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("tr", [False, True])
def test_conv_tiny_channel_kernels_vs_oracle(U, orc, dt, tr):
    """Configs 1 and 3 (1-8 filters / channels): the streaming kernels the AUTO path picks for them, every case of the
    tables whose forward has at most 8 output columns, against the oracle; two backward passes (accumulation)."""
    table = C.TCONV_CASES if tr else C.CONV_CASES
    ran = 0
    for name, case in table.items():
        if case[4] > 8:
            continue
        g, x, w, b, dy = C.conv_inputs(case, dt, 74, tr)
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        a = _conv_gpu(U, case, x, w, b, dy, tr, reps=2, path=U.pkg.PATH_AUTO)
        assert a["path"] == "tiny", (name, a["path"])
        ran += 1
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(dt)], (name, k, C.relerr(a[k], r[k]))
    assert ran >= 3


@pytest.mark.parametrize("dt", DTYPES)
def test_slice_rows_and_fill(U, dt):
    """Mini-batch rows out of a device-resident data set (MemoryDataProvider::get_data, :72-83) and device-side fill."""
    c = U.ctx()
    rng = np.random.default_rng(63)
    for total, vol, first, rows in ((64, 7, 8, 16), (37, 5, 3, 11), (512, 12, 0, 512), (9, 1, 8, 1)):
        data = C.rand(rng, (total, vol), dt)
        src = U.dev(data)
        dst = U.zeros((rows, vol), dt)
        c.slice_rows(total, vol, first, rows, src, dst)
        c.synchronize()
        assert np.array_equal(U.host(dst, (rows, vol)), data[first:first + rows])
    y = U.zeros((1000,), dt)
    c.fill(1000, 2.5, y)
    c.synchronize()
    assert np.array_equal(U.host(y, (1000,)), np.full(1000, 2.5, dt))
    with pytest.raises(U.pkg.Cattl3Error):
        c.slice_rows(10, 3, 8, 5, src, dst)   # rows past the end of the data set


def test_input_feed_ring(U):
    """cattl3_feed: staged uploads on the copy stream.  Every push lands intact (also across the 8 MB staging chunks),
    a slot stays valid for the next slots - 1 pushes while kernels on the compute stream consume it, and is then reused."""
    import torch
    c = U.ctx()
    feed = c.feed_create(3)
    rng = np.random.default_rng(64)
    sizes = [1 << 10, 5 << 20, 3 << 10, (9 << 20) + 13, 1 << 10, 1 << 10, 7 << 20]   # floats; 5M, 9M, 7M floats cross chunks
    addrs, sums = [], []
    for i, n in enumerate(sizes):
        host = rng.uniform(-1, 1, n).astype(np.float32)
        addr = c.feed_push(feed, host)
        addrs.append(addr)
        # consume on the compute stream: copy the slot out
        got = U.zeros((n,), np.float32)
        U.pkg.lib().cattl3_memcpy_d2d(c.h, U.pkg._p(got), ctypes_ptr(addr), ctypes_size(4 * n))
        c.synchronize()
        assert np.array_equal(got.cpu().numpy(), host), i
        host[:] = 0   # the source may be reused as soon as push returns
    # sizes 0 / 3 / 6 share slot 0, which grew twice (a grown slot moves); 1 / 4 share slot 1, which did not grow
    assert addrs[1] == addrs[4] and len(set(addrs[:3])) == 3
    c.feed_destroy(feed)


def ctypes_ptr(addr):
    import ctypes
    return ctypes.c_void_p(addr)


def ctypes_size(n):
    import ctypes
    return ctypes.c_size_t(n)


def test_c_abi_rejects_bad_arguments(U):
    """Errors come back as status codes with a message, never as a crash: null tensors, zero sizes, geometry that
    cannot be ("Errors")."""
    c = U.ctx()
    x = U.zeros((16,), np.float32)
    g = U.pkg.ConvGeom(0, 4, 4, 1, 1, 3, 3, 1, 1, 1, 1, 0, 0)           # batch 0
    with pytest.raises(U.pkg.Cattl3Error) as e:
        c.conv_forward(g, x, x, x, x)
    assert e.value.code == U.pkg.ERR_INVALID
    g = U.pkg.ConvGeom(1, 2, 2, 1, 1, 5, 5, 0, 0, 1, 1, 0, 0)           # receptor larger than the padded input
    with pytest.raises(U.pkg.Cattl3Error):
        c.conv_forward(g, x, x, x, x)
    g = U.pkg.ConvGeom(1, 4, 4, 1, 1, 3, 3, 1, 1, 1, 1, 0, 0)
    with pytest.raises(U.pkg.Cattl3Error):
        c.conv_forward(g, x, None, x, x)                                  # null weights
    with pytest.raises(U.pkg.Cattl3Error):
        c.activation_forward(99, 0.0, 4, 4, x, x)                         # unknown activation kind
    with pytest.raises(U.pkg.Cattl3Error):
        c.dense_forward(0, 4, 4, x, x, x, x)                              # empty batch
    with pytest.raises(U.pkg.Cattl3Error):
        c.loss(7, 4, 4, 0.0, 1.0, x, x, x, x)                             # unknown loss
    with pytest.raises(U.pkg.Cattl3Error):
        c.dropout_forward(16, 1.5, 1e-5, 1, x, x, x)                      # probability out of range
    assert U.pkg.lib().cattl3_last_error()
