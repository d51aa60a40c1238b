"""-m gpu: the fused forward epilogues (cattl3_epilogue: bias + ActivationLayer + BatchNorm statistics) of the
kernel layers against the oracle's layer-by-layer chain conv -> activation / conv -> batch-norm (-> activation),
i.e. what the reference's layer loop computes with separate layers (FeedforwardNeuralNetwork.hpp:112-118).
Tolerances: 1e-4 float / 1e-10 double, norm-relative."""
import numpy as np
import pytest

import cases as C
from oracle.binding import Geom, conv_out_dims

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import gpu_util
    return gpu_util


def _tol(dt):
    return C.TOL[np.dtype(dt)]


# (case, dtype, path it must take)
FUSED_CONV = [("c2_small", np.float32, "tcgen05"), ("c2_small_f256", np.float32, "tcgen05"),
              ("ragged_c", np.float32, "tcgen05"), ("c2_stride2", np.float32, "tcgen05"),
              ("cifar_conv1", np.float32, "tiny"), ("gt_rank3", np.float64, "tiny"), ("c2_small", np.float64, "simt"),
              ("ragged", np.float64, "tiny"),
              # the producer-warp DMMA gather GEMM (128 x 128 and 256 x 64 tiles): activation behind the switch-free epilogue
              ("d2_wide", np.float64, "dmma"), ("d2_ragged", np.float64, "dmma")]


@pytest.mark.parametrize("name,dt,path", FUSED_CONV)
def test_conv_fused_activation_matches_layer_chain(U, orc, name, dt, path):
    case = C.CONV_CASES.get(name) or C.DMMA2_CASES[name]
    g, x, w, b, dy = C.conv_inputs(case, dt, 41)
    oh, ow = conv_out_dims(g)
    shape = (g.n, oh, ow, g.f)
    ref_y = orc.conv(g, x, w, b)["y"]
    c = U.ctx()
    cg = U.pkg.ConvGeom(*case)
    xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
    for aname, (kind, alpha) in C.ACT_CASES.items():
        if aname == "softmax":
            continue
        ref_a = orc.activation(kind, alpha, ref_y)["y"]
        yd, ad = U.zeros(shape, dt), U.zeros(shape, dt)
        c.conv_forward_fused(cg, xd, wd, bd, yd, act_kind=kind, act_param=alpha, act_out=ad)
        assert c.last_path == path, (name, c.last_path)
        c.synchronize()
        assert C.relerr(U.host(yd, shape), ref_y) < _tol(dt), (name, aname, "pre")
        assert C.relerr(U.host(ad, shape), ref_a) < _tol(dt), (name, aname, C.relerr(U.host(ad, shape), ref_a))
        # pre-activation skipped (inference / output-caching activations): same activated tensor, bit for bit
        a2 = U.zeros(shape, dt)
        c.conv_forward_fused(cg, xd, wd, bd, None, act_kind=kind, act_param=alpha, act_out=a2)
        c.synchronize()
        assert np.array_equal(U.host(a2, shape), U.host(ad, shape)), (name, aname)


@pytest.mark.parametrize("name,dt", [("mnist_t0", np.float32), ("wide", np.float32), ("wide_s3_k2", np.float32),
                                     ("gt_rank3", np.float64), ("dfma_t", np.float64)])
def test_transconv_fused_activation_matches_layer_chain(U, orc, name, dt):
    """TransConvKernelLayer -> activation (config 3: TransConv -> Softplus) in the epilogue, bias per output element."""
    case = C.TCONV_CASES[name]
    g, x, w, b, dy = C.conv_inputs(case, dt, 42, True)
    oh, ow = conv_out_dims(g, True)
    shape = (g.n, oh, ow, g.f)
    ref_y = orc.conv(g, x, w, b, transposed=True)["y"]
    c = U.ctx()
    cg = U.pkg.ConvGeom(*case)
    xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
    for aname in ("softplus", "leaky", "tanh"):
        kind, alpha = C.ACT_CASES[aname]
        yd, ad = U.zeros(shape, dt), U.zeros(shape, dt)
        c.conv_forward_fused(cg, xd, wd, bd, yd, act_kind=kind, act_param=alpha, act_out=ad, transposed=True)
        c.synchronize()
        assert C.relerr(U.host(yd, shape), ref_y) < _tol(dt), (name, aname)
        assert C.relerr(U.host(ad, shape), orc.activation(kind, alpha, ref_y)["y"]) < _tol(dt), (name, aname)
    import torch
    with pytest.raises(U.pkg.Cattl3Error):   # no per-filter bias, no column statistics
        c.conv_forward_fused(cg, xd, wd, bd, yd, col_stats=torch.zeros(2 * g.f, dtype=torch.float64, device="cuda"),
                             transposed=True)


@pytest.mark.parametrize("name,dt,path", FUSED_CONV)
def test_conv_fused_batchnorm_statistics(U, orc, name, dt, path):
    """conv (epilogue: column sums) -> batchnorm_forward_stats (+ fused ReLU) equals conv -> BatchNormLayer -> ReLU,
    over two training steps so that the running averages take both branches (assign, then decay)."""
    case = C.CONV_CASES.get(name) or C.DMMA2_CASES[name]
    oh, ow = conv_out_dims(Geom(*case))
    c = U.ctx()
    cg = U.pkg.ConvGeom(*case)
    rng = np.random.default_rng(43)
    F = case[4]
    gm, bt = C.rand(rng, (F,), dt, 0.5, 1.5), C.rand(rng, (F,), dt)
    gmd, btd = U.dev(gm), U.dev(bt)
    rm, rs, sm, ss = (U.zeros((F,), dt) for _ in range(4))
    ys = []
    for step in range(2):
        g, x, w, b, dy = C.conv_inputs(case, dt, 50 + step)
        # a bias far from the column mean and a non-zero-mean input: the shifted sums must not cancel
        x = np.asfortranarray(x + dt(0.75))
        b = np.asfortranarray(b * dt(20))
        shape = (g.n, oh, ow, g.f)
        ys.append(orc.conv(g, x, w, b)["y"])
        xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
        import torch
        yd, nd, ad = U.zeros(shape, dt), U.zeros(shape, dt), U.zeros(shape, dt)
        stats = torch.zeros(2 * F, dtype=torch.float64, device="cuda")
        c.conv_forward_fused(cg, xd, wd, bd, yd, col_stats=stats)
        assert c.last_path == path
        c.batchnorm_forward_stats(1, g.n, oh, ow, F, step > 0, 0.1, 1e-5, yd, stats, bd, gmd, btd, rm, rs, sm, ss, nd,
                                  act_kind=0, act_out=ad)
        c.synchronize()
        # the statistics themselves, against numpy in double
        d = ys[-1].astype(np.float64).reshape(-1, F, order="F") - b.astype(np.float64).reshape(1, F)
        want = np.concatenate([d.sum(0), (d * d).sum(0)])
        got = stats.cpu().numpy()
        assert C.relerr(got[:F], want[:F]) < _tol(dt) and C.relerr(got[F:], want[F:]) < _tol(dt), name
    r = orc.batchnorm(1, ys, gm, bt)
    tol = 10 * _tol(dt) if dt == np.float64 else _tol(dt)
    assert C.relerr(U.host(nd, shape), r["y"]) < tol, (name, C.relerr(U.host(nd, shape), r["y"]))
    assert C.relerr(U.host(rm, (F,)), r["run_mean"]) < tol
    assert C.relerr(U.host(rs, (F,)), r["run_inv_sd"]) < tol
    relu = orc.activation(0, 0.0, r["y"])["y"]
    assert C.relerr(U.host(ad, shape), relu) < tol


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_dense_fused_matches_layer_chain(U, orc, dt):
    """DenseKernelLayer -> activation, and DenseKernelLayer -> per-activation BatchNormLayer (rank 1: one statistic
    per output column over the batch = the column statistics of the GEMM)."""
    c = U.ctx()
    rng = np.random.default_rng(44)
    for name, (n, i, o) in dict(C.DENSE_CASES, tc=(64, 96, 48)).items():
        x, w, b = C.rand(rng, (n, i), dt), C.rand(rng, (i, o), dt, -0.5, 0.5), C.rand(rng, (1, o), dt)
        ref_y = orc.dense(x, w, b)["y"]
        xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
        import torch
        yd, ad = U.zeros((n, o), dt), U.zeros((n, o), dt)
        stats = torch.zeros(2 * o, dtype=torch.float64, device="cuda")
        c.dense_forward_fused(n, i, o, xd, wd, bd, yd, act_kind=3, act_param=1.2, act_out=ad, col_stats=stats)
        if name == "tc" and dt == np.float32:
            assert c.last_path == "tcgen05"
        c.synchronize()
        assert C.relerr(U.host(yd, (n, o)), ref_y) < _tol(dt), name
        assert C.relerr(U.host(ad, (n, o)), orc.activation(3, 1.2, ref_y)["y"]) < _tol(dt), name
        if n < 2:
            continue
        gm, bt = C.rand(rng, (o,), dt, 0.5, 1.5), C.rand(rng, (o,), dt)
        rm, rs, sm, ss = (U.zeros((o,), dt) for _ in range(4))
        nd = U.zeros((n, o), dt)
        c.batchnorm_forward_stats(0, n, 1, 1, o, 0, 0.1, 1e-5, yd, stats, bd, U.dev(gm), U.dev(bt), rm, rs, sm, ss, nd)
        c.synchronize()
        r = orc.batchnorm(0, [np.asfortranarray(ref_y.reshape(n, 1, 1, o, order="F"))], gm, bt)
        tol = 100 * _tol(dt) if dt == np.float64 else 3 * _tol(dt)   # batches of 5-7 rows: 1/sd amplifies rounding
        assert C.relerr(U.host(nd, (n, o)), r["y"].reshape(n, o, order="F")) < tol, (name, "bn")


def test_fused_epilogue_rejects_bad_requests(U):
    import torch
    c = U.ctx()
    case = C.CONV_CASES["c2_small"]
    g, x, w, b, dy = C.conv_inputs(case, np.float32, 45)
    oh, ow = conv_out_dims(g)
    cg = U.pkg.ConvGeom(*case)
    xd, wd, bd = U.dev(x), U.dev(w), U.dev(b)
    yd = U.zeros((g.n, oh, ow, g.f), np.float32)
    with pytest.raises(U.pkg.Cattl3Error):   # softmax is not element-wise
        c.conv_forward_fused(cg, xd, wd, bd, yd, act_kind=7, act_out=yd)
    with pytest.raises(U.pkg.Cattl3Error):   # activation without a destination
        c.conv_forward_fused(cg, xd, wd, bd, yd, act_kind=0)
    with pytest.raises(U.pkg.Cattl3Error):   # statistics without y
        c.conv_forward_fused(cg, xd, wd, bd, None, act_kind=0, act_out=yd,
                             col_stats=torch.zeros(2 * g.f, dtype=torch.float64, device="cuda"))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("pc", [1, 0])
def test_batchnorm_split_entries_reproduce_full_batch(U, orc, dt, pc):
    """The synchronised-statistics protocol on one GPU: two uneven shards of a batch, their shifted sums and counts
    added (what the all-reduce does), forward from the global sums, backward from the global sums with local
    dgamma / dbeta shares -- against the oracle's BatchNormLayer on the whole batch."""
    import torch
    c = U.ctx()
    rng = np.random.default_rng(47)
    n, h, w, ch = 12, 5, 4, 6
    G = ch if pc else h * w * ch
    x = C.rand(rng, (n, h, w, ch), dt, -1, 3)
    gm, bt, dy = C.rand(rng, (G,), dt, 0.5, 1.5), C.rand(rng, (G,), dt), C.rand(rng, (n, h, w, ch), dt)
    r = orc.batchnorm(pc, [x], gm, bt, dy)
    shift = U.dev(C.rand(rng, (G,), dt))
    gmd, btd = U.dev(gm), U.dev(bt)
    shards = [(0, 5), (5, 12)]
    xs = [U.dev(np.asfortranarray(x[lo:hi])) for lo, hi in shards]
    dys = [U.dev(np.asfortranarray(dy[lo:hi])) for lo, hi in shards]
    stats = [torch.zeros(2 * G, dtype=torch.float64, device="cuda") for _ in shards]
    for (lo, hi), xd, st in zip(shards, xs, stats):
        c.batchnorm_stats(pc, hi - lo, h, w, ch, xd, shift, st)
    total = stats[0] + stats[1]
    count = torch.tensor([float(n * (h * w if pc else 1))], dtype=torch.float64, device="cuda")
    dg, db = U.zeros((G,), dt), U.zeros((G,), dt)
    outs, saved, sums = [], [], []
    for (lo, hi), xd, dyd in zip(shards, xs, dys):
        rm, rs, sm, ss = (U.zeros((G,), dt) for _ in range(4))
        yd = U.zeros((hi - lo, h, w, ch), dt)
        c.batchnorm_forward_stats(pc, hi - lo, h, w, ch, 0, 0.1, 1e-5, xd, total, shift, gmd, btd, rm, rs, sm, ss, yd,
                                  global_count=count)
        s = torch.zeros(2 * G, dtype=torch.float64, device="cuda")
        c.batchnorm_backward_sums(pc, hi - lo, h, w, ch, xd, sm, ss, dyd, dg, db, s)
        outs.append(yd); saved.append((rm, rs, sm, ss)); sums.append(s)
    gsum = sums[0] + sums[1]
    dxs = []
    for (lo, hi), xd, dyd, (rm, rs, sm, ss) in zip(shards, xs, dys, saved):
        dxd = U.zeros((hi - lo, h, w, ch), dt)
        c.batchnorm_backward_apply(pc, hi - lo, h, w, ch, count, xd, gmd, sm, ss, dyd, gsum, dxd)
        dxs.append(dxd)
    c.synchronize()
    tol = 10 * _tol(dt) if dt == np.float64 else _tol(dt)
    for (lo, hi), yd, dxd in zip(shards, outs, dxs):
        assert C.relerr(U.host(yd, (hi - lo, h, w, ch)), r["y"][lo:hi]) < tol
        assert C.relerr(U.host(dxd, (hi - lo, h, w, ch)), r["dx"][lo:hi]) < tol
    assert C.relerr(U.host(saved[0][0], (G,)), r["run_mean"]) < tol and C.relerr(U.host(saved[1][1], (G,)), r["run_inv_sd"]) < tol
    assert C.relerr(U.host(dg, (G,)), r["dgamma"]) < tol and C.relerr(U.host(db, (G,)), r["dbeta"]) < tol
