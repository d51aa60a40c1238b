"""CPU tests that PIN the oracle: the C restatement (oracle/cattl3_oracle_impl.h) must reproduce
(a) the committed golden vectors generated from the unmodified reference and (b), where the
reference shim is built, the reference itself on every case of tests/cases.py."""
import numpy as np
import pytest

import cases as C
from oracle.binding import Geom

DTYPES = [("f32", np.float32), ("f64", np.float64)]
# oracle-vs-reference tolerance: both are CPU; float paths differ only in summation order
OTOL = {np.float32: 2e-5, np.float64: 1e-13}


def _keys(golden, prefix):
    names = sorted({k[len(prefix):].split("/")[0] for k in golden.files if k.startswith(prefix)})
    assert names, prefix
    return names


@pytest.mark.parametrize("suf,dt", DTYPES)
@pytest.mark.parametrize("tr", [False, True])
def test_conv_golden(orc, golden, suf, dt, tr):
    fam = "tconv" if tr else "conv"
    table = C.TCONV_CASES if tr else C.CONV_CASES
    for name in _keys(golden, fam + "/"):
        k = "%s/%s/%s/" % (fam, name, suf)
        g = Geom(*table[name])
        x, w, b, dy = (np.asfortranarray(golden[k + s]) for s in ("x", "w", "b", "dy"))
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        for out in ("y", "dx", "dw", "db"):
            assert C.relerr(r[out], golden[k + out]) < OTOL[dt], (name, out)


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_dense_golden(orc, golden, suf, dt):
    for name in _keys(golden, "dense/"):
        k = "dense/%s/%s/" % (name, suf)
        x, w, b, dy = (np.asfortranarray(golden[k + s]) for s in ("x", "w", "b", "dy"))
        r = orc.dense(x, w, b, dy, back_reps=2)
        for out in ("y", "dx", "dw", "db"):
            assert C.relerr(r[out], golden[k + out]) < OTOL[dt], (name, out)


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_activation_golden(orc, golden, suf, dt):
    for name, (kind, alpha) in C.ACT_CASES.items():
        k = "act/%s/%s/" % (name, suf)
        x, dy = np.asfortranarray(golden[k + "x"]), np.asfortranarray(golden[k + "dy"])
        r = orc.activation(kind, alpha, x, dy)
        assert C.relerr(r["y"], golden[k + "y"]) < OTOL[dt], name
        assert C.relerr(r["dx"], golden[k + "dx"]) < OTOL[dt], name


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_pool_golden(orc, golden, suf, dt):
    for name in _keys(golden, "pool/"):
        kind, n, h, w, c, rh, rw, sh, sw = C.POOL_CASES[name]
        k = "pool/%s/%s/" % (name, suf)
        x, dy = np.asfortranarray(golden[k + "x"]), np.asfortranarray(golden[k + "dy"])
        r = orc.pool(kind, x, rh, rw, sh, sw, dy)
        # max pooling is pure selection/routing: bit-exact, ties included
        assert np.array_equal(r["y"], golden[k + "y"]) or C.relerr(r["y"], golden[k + "y"]) < OTOL[dt]
        assert C.relerr(r["dx"], golden[k + "dx"]) < OTOL[dt], name


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_batchnorm_golden(orc, golden, suf, dt):
    for name in _keys(golden, "bn/"):
        pc, n, h, w, c, steps = C.BN_CASES[name]
        k = "bn/%s/%s/" % (name, suf)
        xs = [np.asfortranarray(golden[k + "x%d" % s]) for s in range(steps)]
        r = orc.batchnorm(pc, xs, golden[k + "gamma"], golden[k + "beta"], np.asfortranarray(golden[k + "dy"]))
        for out in ("y", "dx", "dgamma", "dbeta", "run_mean", "run_inv_sd", "y_infer"):
            assert C.relerr(r[out], golden[k + out]) < 4 * OTOL[dt], (name, out)


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_optimizer_golden(orc, golden, suf, dt):
    for name, (kind, hy) in C.OPT_CASES.items():
        for lam in (0.0, 0.01):
            k = "opt/%s_l2_%g/%s/" % (name, lam, suf)
            p0 = np.asfortranarray(golden[k + "p0"])
            grads = [np.asfortranarray(g) for g in golden[k + "grads"]]
            p = orc.optimizer(kind, hy, lam, p0, grads, 3)
            assert C.relerr(p, golden[k + "p"]) < OTOL[dt], (name, lam)
            assert C.relerr(p0, golden[k + "p"]) > 1e-4  # the update moved the parameters


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_oracle_vs_reference_all_cases(orc, ref, suf, dt):
    """Every geometry of tests/cases.py, including the ones too large for the golden file."""
    for tr, table in ((False, C.CONV_CASES), (True, C.TCONV_CASES)):
        for name, case in table.items():
            g, x, w, b, dy = C.conv_inputs(case, dt, 5, tr)
            a = orc.conv(g, x, w, b, dy, transposed=tr)
            r = ref.conv(g, x, w, b, dy, transposed=tr)
            for out in ("y", "dx", "dw", "db"):
                assert C.relerr(a[out], r[out]) < OTOL[dt], (name, out)
    # input layer: no dX requested, gradients unchanged
    g, x, w, b, dy = C.conv_inputs(C.CONV_CASES["gt_rank3"], dt, 6)
    a, r = orc.conv(g, x, w, b, dy, want_dx=False), ref.conv(g, x, w, b, dy, want_dx=False)
    assert a["dx"] is None and C.relerr(a["dw"], r["dw"]) < OTOL[dt]


def test_golden_reproducible_from_reference(ref, golden):
    """The committed vectors are what the reference build in this container computes today."""
    k = "conv/gt_rank3/f64/"
    g = Geom(*C.CONV_CASES["gt_rank3"])
    x, w, b, dy = (np.asfortranarray(golden[k + s]) for s in ("x", "w", "b", "dy"))
    r = ref.conv(g, x, w, b, dy, back_reps=2)
    assert C.relerr(r["y"], golden[k + "y"]) < 1e-14
    assert C.relerr(r["dw"], golden[k + "dw"]) < 1e-14


@pytest.mark.parametrize("suf,dt", DTYPES)
def test_network_fixtures_reproduce_from_the_reference(ref, suf, dt):
    """tests/golden/reference_networks.npz (configs 3, 4, 5 and the regularised config 1) is what the unmodified reference computes from the
    seeded inputs of tests/cases.py: re-run it where oracle/_ref exists."""
    import os
    nets = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_networks.npz"))
    x = C.autoencoder_inputs(dt)
    n = ref.train_autoencoder(x, 4, -1)
    p1, loss, _ = ref.train_autoencoder(x, 4, 2, params_in=C.seeded_params(n, dt, 3002))
    assert C.relerr(p1, nets["autoencoder/%s/p1" % suf]) < 10 * OTOL[dt]
    x, obj = C.resnet_inputs(dt)
    n = ref.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)
    p1, loss, _ = ref.train_resnet(x, obj, 32, 2, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))
    assert C.relerr(p1, nets["resnet/%s/p1" % suf]) < 10 * OTOL[dt]
    assert abs(loss - float(nets["resnet/%s/loss" % suf][0])) < 1e-5
    # config 5: Sequential{Parallel conv lanes, DenseNet, MaxPool} -> convolutional LSTM
    x, obj = C.seqnet_inputs(dt)
    n = ref.train_seqnet(x, obj, 8, -1, **C.SEQNET_SMALL)
    p1, loss, _ = ref.train_seqnet(x, obj, 8, 2, params_in=C.seeded_params(n, dt, 5002), **C.SEQNET_SMALL)
    assert C.relerr(p1, nets["seqnet/%s/p1" % suf]) < 10 * OTOL[dt]
    assert abs(loss - float(nets["seqnet/%s/loss" % suf][0])) < 1e-5
    # config 1 with an ElasticNet penalty on every weight matrix (REF_SHIM_REG)
    vectors = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.npz"))
    x, obj = C.cifar_inputs(dt)
    os.environ["REF_SHIM_REG"] = C.CIFAR_REG
    try:
        p1, loss, _ = ref.train_cifar(x, obj, 16, 2, params_in=np.ascontiguousarray(vectors["cifar/%s/p0" % suf]))
    finally:
        del os.environ["REF_SHIM_REG"]
    assert C.relerr(p1, nets["cifar_reg/%s/p1" % suf]) < 10 * OTOL[dt]
    assert abs(loss - float(nets["cifar_reg/%s/loss" % suf][0])) < 1e-5
