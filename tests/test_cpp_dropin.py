"""-m gpu: the header-only C++ host side (c-attl3_b200/cattle) as a drop-in for the reference.

tests/cpp/Makefile compiles code written against the reference, UNCHANGED, with the B200 headers first
on the include path:

* ``libcattle_b200_shim.so`` -- oracle/ref_shim.cpp, the driver that wraps the reference's layers,
  optimizers and the config-1 training loop for the oracle, now instantiating the B200 classes of the
  same names.  Its results are compared with the C oracle, with the golden vectors of the unmodified
  reference and -- where oracle/_ref travelled -- with the reference itself, at the north-star
  tolerances (1e-4 float, 1e-10 double, norm-relative).
* ``gradient_test_b200`` -- the reference's own test/gradient_test.cpp (finite-difference gradient
  checks, double, N=5), "the repo's own gradient_test" of BASELINE.json's north star.

Both binaries are built where /root/reference exists (``__graft_entry__.build()``) and travel to the GPU box.
"""
import os
import subprocess

import numpy as np
import pytest

import cases as C
from oracle import binding
from oracle.binding import Geom

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "cpp", "_build")
SHIM = os.path.join(BUILD, "libcattle_b200_shim.so")
GRADIENT_TEST = os.path.join(BUILD, "gradient_test_b200")
DTYPES = [np.float32, np.float64]

# the hot-path tests of the reference's gradient_test.cpp and the networks built on them
GTEST_FILTER = ":".join("GradientTest." + t for t in (
    "DenseKernelLayer", "ConvKernelLayer", "TransConvKernelLayer", "ActivationLayer", "PoolLayer",
    "BatchNormLayer", "ResidualNet", "DenseNet", "ParallelNet", "SequentialNet", "RecurrentNet", "LSTMNet", "BidirectionalNet"))


@pytest.fixture(scope="module")
def b200():
    if not os.path.exists(SHIM):
        pytest.fail("%s missing: run __graft_entry__.build() where /root/reference exists" % SHIM)
    lib = binding.Oracle("ref", path=SHIM)
    assert lib.lib.ref_is_b200_build() == 1
    return lib


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("tr", [False, True])
def test_kernel_layers_match_oracle(b200, orc, dt, tr):
    """ConvKernelLayer / TransConvKernelLayer<S,3>::pass_forward / pass_back (host tensors in and out)."""
    table = C.TCONV_CASES if tr else C.CONV_CASES
    for name, case in table.items():
        g, x, w, b, dy = C.conv_inputs(case, dt, 41, tr)
        r = orc.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        a = b200.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(dt)], (name, k, C.relerr(a[k], r[k]))


@pytest.mark.parametrize("dt", DTYPES)
def test_dense_layer_matches_oracle(b200, orc, dt):
    rng = np.random.default_rng(42)
    for name, (n, i, o) in C.DENSE_CASES.items():
        x, w, b, dy = C.rand(rng, (n, i), dt), C.rand(rng, (i, o), dt), C.rand(rng, (1, o), dt), C.rand(rng, (n, o), dt)
        r = orc.dense(x, w, b, dy, back_reps=2)
        a = b200.dense(x, w, b, dy, back_reps=2)
        for k in ("y", "dx", "dw", "db"):
            assert C.relerr(a[k], r[k]) < C.TOL[np.dtype(dt)], (name, k)


@pytest.mark.parametrize("dt", DTYPES)
def test_activation_pool_batchnorm_match_golden(b200, golden, dt):
    suf = "f32" if dt == np.float32 else "f64"
    tol = C.TOL[np.dtype(dt)]
    for name, (kind, alpha) in C.ACT_CASES.items():
        k = "act/%s/%s/" % (name, suf)
        x, dy = np.asfortranarray(golden[k + "x"]), np.asfortranarray(golden[k + "dy"])
        a = b200.activation(kind, alpha, x, dy)
        assert C.relerr(a["y"], golden[k + "y"]) < tol, name
        assert C.relerr(a["dx"], golden[k + "dx"]) < tol, name
    for name, (kind, n, h, w, c, rh, rw, sh, sw) in C.POOL_CASES.items():
        k = "pool/%s/%s/" % (name, suf)
        if k + "x" not in golden.files:
            continue
        x, dy = np.asfortranarray(golden[k + "x"]), np.asfortranarray(golden[k + "dy"])
        a = b200.pool(kind, x, rh, rw, sh, sw, dy)
        if kind == 0:
            assert np.array_equal(a["y"], golden[k + "y"]), name
        assert C.relerr(a["y"], golden[k + "y"]) < tol, name
        assert C.relerr(a["dx"], golden[k + "dx"]) < tol, name
    for name, (pc, n, h, w, c, steps) in C.BN_CASES.items():
        k = "bn/%s/%s/" % (name, suf)
        if k + "x0" not in golden.files:
            continue
        xs = [np.asfortranarray(golden[k + "x%d" % s]) for s in range(steps)]
        a = b200.batchnorm(pc, xs, golden[k + "gamma"], golden[k + "beta"], np.asfortranarray(golden[k + "dy"]))
        btol = 10 * tol if dt == np.float64 else tol
        for out in ("y", "dx", "dgamma", "dbeta", "run_mean", "run_inv_sd", "y_infer"):
            assert C.relerr(a[out], golden[k + out]) < btol, (name, out, C.relerr(a[out], golden[k + out]))


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("hostparams", [False, True])
def test_optimizers_match_golden(b200, golden, dt, hostparams):
    """Every SGDOptimizer subclass: device-resident B200Parameters (fused step in place) and host
    StandardParameters (staged through the same kernel)."""
    suf = "f32" if dt == np.float32 else "f64"
    for name, (kind, hy) in C.OPT_CASES.items():
        for lam in (0.0, 0.01):
            k = "opt/%s_l2_%g/%s/" % (name, lam, suf)
            p0, grads = np.asfortranarray(golden[k + "p0"]), [np.asfortranarray(g) for g in golden[k + "grads"]]
            p = b200.optimizer(kind, hy, lam, p0, grads, 3, hostparams=hostparams)
            tol = 2e-6 if dt == np.float32 else 1e-12
            assert C.relerr(p, golden[k + "p"]) < tol, (name, lam, C.relerr(p, golden[k + "p"]))


@pytest.mark.parametrize("dt", DTYPES)
def test_config1_training_matches_reference(b200, golden, dt):
    """BASELINE.json configs[0]: the cifar ConvNet (Dropout removed) under NadamOptimizer::train through the
    B200 FeedforwardNeuralNetwork / SGDOptimizer batch loop; parameters after the epoch and the epoch
    loss against the unmodified reference (golden fixture; and live where oracle/_ref travelled)."""
    suf = "f32" if dt == np.float32 else "f64"
    rng = np.random.default_rng(1001)
    x = C.rand(rng, (8, 32, 32, 3), dt)
    obj = np.zeros((8, 1, 1, 10), dtype=dt, order="F")
    for i in range(8):
        obj[i, 0, 0, i % 10] = 1
    p0 = np.ascontiguousarray(golden["cifar/%s/p0" % suf])
    p1, loss, _ = b200.train_cifar(x, obj, 4, 1, params_in=p0)
    tol = 1e-4 if dt == np.float32 else 1e-10
    assert C.relerr(p1, golden["cifar/%s/p1" % suf]) < tol, C.relerr(p1, golden["cifar/%s/p1" % suf])
    assert abs(loss - float(golden["cifar/%s/loss" % suf][0])) < tol * max(1.0, abs(loss))
    if binding.have_ref():
        ref = binding.Oracle("ref")
        rng = np.random.default_rng(1002)
        n = 64
        x = C.rand(rng, (n, 32, 32, 3), dt)
        obj = np.zeros((n, 1, 1, 10), dtype=dt, order="F")
        obj[np.arange(n), 0, 0, np.arange(n) % 10] = 1
        pr, lr, _ = ref.train_cifar(x, obj, 16, 2, params_in=p0)
        pb, lb, _ = b200.train_cifar(x, obj, 16, 2, params_in=p0)
        print("config 1, 2 epochs of 64 samples (8 Nadam steps): loss ref %.6f b200 %.6f, param err %.2e"
              % (lr, lb, C.relerr(pb, pr)))
        assert C.relerr(pb, pr) < (1e-4 if dt == np.float32 else 1e-10)
        assert abs(lr - lb) < (1e-4 if dt == np.float32 else 1e-10) * max(1.0, abs(lr))


@pytest.fixture(scope="module")
def golden_nets():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_networks.npz"))


@pytest.mark.parametrize("dt", DTYPES)
def test_config3_autoencoder_training_matches_reference(b200, golden_nets, dt):
    """BASELINE.json configs[2]: examples/mnist_autoencoder.cpp's StackedNeuralNetwork (Conv, Softplus, Dense,
    Reshape, TransConv) + SquaredLoss + Nadam through the device-resident batch loop: parameters and epoch loss
    after 4 steps against the unmodified reference (fixture; live where oracle/_ref travelled)."""
    suf = "f32" if dt == np.float32 else "f64"
    x = C.autoencoder_inputs(dt)
    n = b200.train_autoencoder(x, 4, -1)
    assert n == golden_nets["autoencoder/%s/p1" % suf].size
    p0 = C.seeded_params(n, dt, 3002)
    p1, loss, _ = b200.train_autoencoder(x, 4, 2, params_in=p0)
    tol = 1e-4 if dt == np.float32 else 1e-10
    err = C.relerr(p1, golden_nets["autoencoder/%s/p1" % suf])
    print("config 3 auto-encoder, 4 Nadam steps: loss %.6f (ref %.6f), param err %.2e"
          % (loss, float(golden_nets["autoencoder/%s/loss" % suf][0]), err))
    assert err < tol
    assert abs(loss - float(golden_nets["autoencoder/%s/loss" % suf][0])) < tol * max(1.0, abs(loss))
    if binding.have_ref():
        ref = binding.Oracle("ref")
        x = C.autoencoder_inputs(dt, total=64, seed=3003)
        pr, lr, _ = ref.train_autoencoder(x, 32, 2, params_in=p0)
        pb, lb, _ = b200.train_autoencoder(x, 32, 2, params_in=p0)
        assert C.relerr(pb, pr) < (1e-4 if dt == np.float32 else 1e-10), C.relerr(pb, pr)
        assert abs(lr - lb) < (1e-4 if dt == np.float32 else 1e-10) * max(1.0, abs(lr))


@pytest.mark.parametrize("dt", DTYPES)
def test_config4_resnet_training_matches_reference(b200, golden_nets, dt):
    """BASELINE.json configs[3] at test size: stem + ResidualNeuralNetwork of conv + BatchNorm + ReLU modules +
    head, CrossEntropyLoss, Nadam.  Exercises the fused epilogues (conv -> BatchNorm statistics -> ReLU) inside the
    network loop, the device loss and the residual adds; compared with the unmodified reference."""
    suf = "f32" if dt == np.float32 else "f64"
    x, obj = C.resnet_inputs(dt)
    n = b200.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)
    assert n == golden_nets["resnet/%s/p1" % suf].size
    p0 = C.seeded_params(n, dt, 4002)
    p1, loss, _ = b200.train_resnet(x, obj, 32, 2, C.RESNET_SMALL, params_in=p0)
    tol = 1e-4 if dt == np.float32 else 1e-10
    err = C.relerr(p1, golden_nets["resnet/%s/p1" % suf])
    print("config 4 ResNet (test size), 4 Nadam steps: loss %.6f (ref %.6f), param err %.2e"
          % (loss, float(golden_nets["resnet/%s/loss" % suf][0]), err))
    assert err < tol
    assert abs(loss - float(golden_nets["resnet/%s/loss" % suf][0])) < tol * max(1.0, abs(loss))


@pytest.mark.parametrize("dt", DTYPES)
def test_regularised_training_matches_reference(golden, golden_nets, dt):
    """Config 1 with an ElasticNet penalty on every weight matrix (REF_SHIM_REG): Parameters::regularize and
    get_regularization_penalty run as one device kernel per parameter (cattl3_regularize; the penalty is accumulated on
    the device and read once per epoch), also inside the captured step graph.  Parameters and the epoch loss
    (objective + penalty) against the unmodified reference; graph and eager runs agree."""
    suf = "f32" if dt == np.float32 else "f64"
    code = (
        "import os, sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%r); dt = np.%s\n"
        "x, obj = C.cifar_inputs(dt); os.environ['REF_SHIM_REG'] = C.CIFAR_REG\n"
        "p0 = np.ascontiguousarray(np.load(%r)['cifar/%s/p0'])\n"
        "p, l, _ = lib.train_cifar(x, obj, 16, 2, params_in=p0)\n"
        "np.save(sys.argv[1], np.concatenate([np.asarray(p, dtype=np.float64).ravel(), [l]]))\n"
        % (ROOT, os.path.join(ROOT, "tests"), SHIM, np.dtype(dt).name,
           os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), suf))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({"CATTL3_GRAPH_TRACE": "1"}, {"CATTL3_NO_GRAPH": "1"})):
            path = os.path.join(d, "p%d.npy" % i)
            r = subprocess.run([os.sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            if i == 0:
                assert "step graph captured" in r.stderr, "regularised parameters kept the step out of the graph:\n" + r.stderr[-2000:]
            outs.append(np.load(path))
    ref_p, ref_loss = golden_nets["cifar_reg/%s/p1" % suf], float(golden_nets["cifar_reg/%s/loss" % suf][0])
    err = C.relerr(outs[0][:-1], ref_p)
    print("config 1 + ElasticNet, 8 Nadam steps: loss %.6f (ref %.6f), param err %.2e, graph vs eager %.2e"
          % (outs[0][-1], ref_loss, err, C.relerr(outs[0][:-1], outs[1][:-1])))
    assert err < (1e-4 if dt == np.float32 else 1e-10)
    assert abs(outs[0][-1] - ref_loss) < (1e-4 if dt == np.float32 else 1e-10) * max(1.0, abs(ref_loss))
    assert C.relerr(outs[0][:-1], outs[1][:-1]) < (1e-6 if dt == np.float32 else 1e-13)


def _run_cifar_shim(shim_path, dt, env, epochs, out_path, load_p0=True):
    """The config-1 driver of oracle/ref_shim.cpp in a fresh process (its REF_SHIM_* switches are read from the environment)."""
    suf = "f32" if dt == np.float32 else "f64"
    code = (
        "import os, sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%s); dt = np.%s\n"
        "x, obj = C.cifar_inputs(dt)\n"
        "p0 = np.ascontiguousarray(np.load(%r)['cifar/%s/p0']) if %r else None\n"
        "p, l, _ = lib.train_cifar(x, obj, 16, %d, params_in=p0)\n"
        "np.save(sys.argv[1], np.concatenate([np.asarray(p, dtype=np.float64).ravel(), [l]]))\n"
        % (ROOT, os.path.join(ROOT, "tests"), repr(shim_path) if shim_path else "None", np.dtype(dt).name,
           os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"), suf, load_p0, epochs))
    r = subprocess.run([os.sys.executable, "-c", code, out_path], env=dict(os.environ, **env), capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return np.load(out_path), r.stderr


@pytest.mark.parametrize("dt", DTYPES)
def test_constrained_training_matches_reference(golden_nets, dt):
    """Config 1 with all six constraints of StandardParameters on every weight matrix (REF_SHIM_CONSTRAINTS; value clip,
    the Frobenius-norm "L1" limit, the squared-norm "L2" limit and the same three on the gradient,
    StandardParameters.hpp:150-182): on the B200 build they run as cattl3_constrain behind each backward pass and each
    optimizer step, also inside the captured step graph.  Parameters and loss against the unmodified reference."""
    import tempfile
    suf = "f32" if dt == np.float32 else "f64"
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({"CATTL3_GRAPH_TRACE": "1"}, {"CATTL3_NO_GRAPH": "1"})):
            o, err = _run_cifar_shim(SHIM, dt, dict(env, REF_SHIM_CONSTRAINTS=C.CIFAR_CON), 2, os.path.join(d, "p%d.npy" % i))
            if i == 0:
                assert "step graph captured" in err, "constrained parameters kept the step out of the graph:\n" + err[-2000:]
            outs.append(o)
    ref_p, ref_loss = golden_nets["cifar_con/%s/p1" % suf], float(golden_nets["cifar_con/%s/loss" % suf][0])
    e = C.relerr(outs[0][:-1], ref_p)
    print("config 1 + six constraints, 8 Nadam steps: loss %.6f (ref %.6f), param err %.2e, graph vs eager %.2e"
          % (outs[0][-1], ref_loss, e, C.relerr(outs[0][:-1], outs[1][:-1])))
    assert e < (1e-4 if dt == np.float32 else 1e-10)
    assert abs(outs[0][-1] - ref_loss) < (1e-4 if dt == np.float32 else 1e-10) * max(1.0, abs(ref_loss))
    assert C.relerr(outs[0][:-1], outs[1][:-1]) < (1e-6 if dt == np.float32 else 1e-13)


@pytest.mark.parametrize("dt", DTYPES)
def test_prms_files_round_trip_and_interoperate_with_the_reference(dt, text=False):
    """NeuralNetwork::save_all_unique_params_values / load_all_unique_params_values (NeuralNetwork.hpp:195-222; binary .prms
    format of EigenProxy.hpp:147-224) through B200Parameters: a trained B200 network's files load back bit-identical into a
    fresh B200 network AND into the unmodified reference, the files the reference writes for the same parameters are
    byte-identical, and a reference-written directory loads into the B200 build.  (The reference's TEXT form does not round
    trip in the reference itself: serialize() writes sizeof(Scalar) first, EigenProxy.hpp:117, deserialize() does not read
    it, :176-179 -- that code is the reference's own on both sides and is left alone.)"""
    import filecmp
    import tempfile
    from oracle import binding
    tol_text = 1e-5 if dt == np.float32 else 1e-5   # operator<< prints 6 significant digits
    fmt = {"REF_SHIM_PRMS_TEXT": "1"} if text else {}
    with tempfile.TemporaryDirectory() as d:
        a, b, c = (os.path.join(d, n) for n in ("saved_by_b200", "saved_by_ref", "resaved_by_b200"))
        for p in (a, b, c):
            os.mkdir(p)
        # train one epoch on the B200 and save
        trained, _ = _run_cifar_shim(SHIM, dt, dict(fmt, REF_SHIM_PRMS_SAVE=a), 1, os.path.join(d, "t.npy"))
        files = sorted(os.listdir(a))
        assert len(files) == 8 and all(f.endswith(".prms") for f in files), files
        # a fresh B200 network (random init, no injected parameters) loads them
        loaded, _ = _run_cifar_shim(SHIM, dt, dict(fmt, REF_SHIM_PRMS_LOAD=a, REF_SHIM_PRMS_SAVE=c), 0, os.path.join(d, "l.npy"),
                                    load_p0=False)
        if text:
            assert C.relerr(loaded[:-1], trained[:-1]) < tol_text
        else:
            assert np.array_equal(loaded[:-1], trained[:-1])
            for f in files:
                assert filecmp.cmp(os.path.join(a, f), os.path.join(c, f), shallow=False), f
        if binding.have_ref():
            # the unmodified reference reads the B200's files, writes its own: same parameters, same bytes
            ref_loaded, _ = _run_cifar_shim(None, dt, dict(fmt, REF_SHIM_PRMS_LOAD=a, REF_SHIM_PRMS_SAVE=b), 0,
                                            os.path.join(d, "r.npy"), load_p0=False)
            if text:
                assert C.relerr(ref_loaded[:-1], trained[:-1]) < tol_text
            else:
                assert np.array_equal(ref_loaded[:-1], trained[:-1])
            for f in files:
                assert filecmp.cmp(os.path.join(a if not text else c, f), os.path.join(b, f), shallow=False), f
            # and a directory written by the reference loads into the B200 build
            back, _ = _run_cifar_shim(SHIM, dt, dict(fmt, REF_SHIM_PRMS_LOAD=b), 0, os.path.join(d, "k.npy"), load_p0=False)
            assert np.array_equal(back[:-1], ref_loaded[:-1])


@pytest.mark.parametrize("dt", DTYPES)
def test_config5_sequence_network_training_matches_reference(b200, golden_nets, dt):
    """BASELINE.json configs[4] at test size: SequentialNeuralNetwork{ParallelNeuralNetwork of conv lanes,
    DenseNeuralNetwork, MaxPool} feeding a convolutional LSTMNeuralNetwork (ConvKernelLayer kernels, shared across the
    time steps), sequential SquaredLoss + Nadam: parameters and epoch loss after 4 steps against the unmodified
    reference (fixture; live where oracle/_ref travelled)."""
    suf = "f32" if dt == np.float32 else "f64"
    x, obj = C.seqnet_inputs(dt)
    n = b200.train_seqnet(x, obj, 8, -1, **C.SEQNET_SMALL)
    assert n == golden_nets["seqnet/%s/p1" % suf].size
    p0 = C.seeded_params(n, dt, 5002)
    p1, loss, _ = b200.train_seqnet(x, obj, 8, 2, params_in=p0, **C.SEQNET_SMALL)
    tol = 1e-4 if dt == np.float32 else 1e-10
    err = C.relerr(p1, golden_nets["seqnet/%s/p1" % suf])
    print("config 5 sequence network, 4 Nadam steps: loss %.6f (ref %.6f), param err %.2e"
          % (loss, float(golden_nets["seqnet/%s/loss" % suf][0]), err))
    assert err < tol
    assert abs(loss - float(golden_nets["seqnet/%s/loss" % suf][0])) < tol * max(1.0, abs(loss))
    if binding.have_ref():
        ref = binding.Oracle("ref")
        x, obj = C.seqnet_inputs(dt, total=24, seq=4, seed=5003)
        pr, lr, _ = ref.train_seqnet(x, obj, 8, 2, params_in=p0, **C.SEQNET_SMALL)
        pb, lb, _ = b200.train_seqnet(x, obj, 8, 2, params_in=p0, **C.SEQNET_SMALL)
        assert C.relerr(pb, pr) < (1e-4 if dt == np.float32 else 1e-10), C.relerr(pb, pr)
        assert abs(lr - lb) < (1e-4 if dt == np.float32 else 1e-10) * max(1.0, abs(lr))


@pytest.mark.parametrize("dt", DTYPES)
def test_sequence_network_device_loop_equals_host_loop(dt):
    """Config 5 trains through the device sequence path (b200::DeviceSequenceNetwork: the time-step fold as a view, the
    LSTM unrolled in HBM, device loss); CATTL3_HOST_LOOP=1 keeps the reference's host protocol between network and loss
    (every network then uploads / downloads around itself).  Same run, same parameters."""
    code = (
        "import sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%r); dt = np.%s\n"
        "x, obj = C.seqnet_inputs(dt, total=24, seq=4, seed=5004); n = lib.train_seqnet(x, obj, 8, -1, **C.SEQNET_SMALL)\n"
        "p, l, _ = lib.train_seqnet(x, obj, 8, 2, params_in=C.seeded_params(n, dt, 5002), **C.SEQNET_SMALL)\n"
        "np.save(sys.argv[1], p)\n" % (ROOT, os.path.join(ROOT, "tests"), SHIM, np.dtype(dt).name))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({}, {"CATTL3_HOST_LOOP": "1"})):
            path = os.path.join(d, "p%d.npy" % i)
            r = subprocess.run([os.sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(path))
    err = C.relerr(outs[0], outs[1])
    print("config 5: device loop vs host loop parameters after 6 steps: %.2e" % err)
    assert err < (2e-6 if dt == np.float32 else 1e-13)


@pytest.mark.parametrize("dt", DTYPES)
def test_fused_and_unfused_network_loops_agree(dt):
    """CATTL3_NO_FUSION=1 / CATTL3_HOST_LOOP=1 switch the epilogue fusion and the device batch loop off: the same
    training run must give the same parameters either way (fusion changes where work happens, not what is computed)."""
    code = (
        "import sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%r); dt = np.%s\n"
        "x, obj = C.resnet_inputs(dt); n = lib.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)\n"
        "p, l, _ = lib.train_resnet(x, obj, 32, 1, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))\n"
        "np.save(sys.argv[1], p)\n" % (ROOT, os.path.join(ROOT, "tests"), SHIM, np.dtype(dt).name))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({}, {"CATTL3_NO_FUSION": "1", "CATTL3_HOST_LOOP": "1"})):
            path = os.path.join(d, "p%d.npy" % i)
            r = subprocess.run([os.sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(path))
    err = C.relerr(outs[0], outs[1])
    print("fused vs unfused parameters after 2 steps: %.2e" % err)
    assert err < (2e-5 if dt == np.float32 else 1e-10)


@pytest.mark.parametrize("net", ["cifar", "autoencoder", "resnet", "seqnet"])
@pytest.mark.parametrize("dt", DTYPES)
def test_step_graph_equals_eager_loop(dt, net):
    """The batch loop captures a launch-bound training step as a CUDA graph after two eager steps at a shape
    (cattl3_graph_*, SGDOptimizer.hpp StepGraph) and replays it; CATTL3_NO_GRAPH=1 keeps every step eager.  Same run,
    same parameters and the same epoch loss either way -- including a ragged last batch (run eagerly between replays),
    a second epoch (the graph outlives the epoch), BatchNorm's running statistics (resnet) and the unrolled convolutional
    LSTM of config 5 (seqnet: its time-step copies and gate kernels are nodes of the graph)."""
    body = {
        "cifar": "x = C.rand(np.random.default_rng(1003), (72, 32, 32, 3), dt)\n"
                 "obj = np.zeros((72, 1, 1, 10), dtype=dt, order='F'); obj[np.arange(72), 0, 0, np.arange(72) % 10] = 1\n"
                 "p, l, _ = lib.train_cifar(x, obj, 16, 2, params_in=np.ascontiguousarray(G['cifar/%s/p0']))\n",
        "autoencoder": "x = C.autoencoder_inputs(dt, total=200, seed=3004); n = lib.train_autoencoder(x, 32, -1)\n"
                       "p, l, _ = lib.train_autoencoder(x, 32, 2, params_in=C.seeded_params(n, dt, 3002))\n",
        "resnet": "x, obj = C.resnet_inputs(dt, total=176); n = lib.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)\n"
                  "p, l, _ = lib.train_resnet(x, obj, 32, 2, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))\n",
        "seqnet": "x, obj = C.seqnet_inputs(dt, total=44, seq=4, seed=5005); n = lib.train_seqnet(x, obj, 8, -1, **C.SEQNET_SMALL)\n"
                  "p, l, _ = lib.train_seqnet(x, obj, 8, 2, params_in=C.seeded_params(n, dt, 5002), **C.SEQNET_SMALL)\n",
    }[net]
    if net == "cifar":
        body = body.replace("%s", "f32" if dt == np.float32 else "f64")
    code = (
        "import os, sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%r); dt = np.%s\n"
        "G = np.load(%r)\n" % (ROOT, os.path.join(ROOT, "tests"), SHIM, np.dtype(dt).name,
                                os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
        + body + "np.save(sys.argv[1], np.concatenate([np.asarray(p, dtype=np.float64).ravel(), [l]]))\n")
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({"CATTL3_GRAPH_TRACE": "1"}, {"CATTL3_NO_GRAPH": "1"})):
            path = os.path.join(d, "p%d.npy" % i)
            r = subprocess.run([os.sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            if i == 0:
                assert "step graph captured" in r.stderr, "the graph path did not engage:\n" + r.stderr[-2000:]
                assert "replays" in r.stderr
            outs.append(np.load(path))
    err = C.relerr(outs[0][:-1], outs[1][:-1])
    print("%s: step graph vs eager parameters %.2e, loss %.8f vs %.8f" % (net, err, outs[0][-1], outs[1][-1]))
    assert err < (1e-6 if dt == np.float32 else 1e-13)
    assert abs(outs[0][-1] - outs[1][-1]) < (1e-5 if dt == np.float32 else 1e-12) * max(1.0, abs(outs[1][-1]))


@pytest.mark.parametrize("dt", DTYPES)
def test_dropout_layer_contract(b200, dt):
    """The B200 DropoutLayer class (device RNG) honours the reference layer's contract (DropoutLayer.hpp:74-94):
    inverted dropout with the same mask forward and backward, survivors scaled by 1 / (1 - p + eps), inference = identity;
    where the reference travelled, its own layer is held to the same checks (its masks differ: host RNG)."""
    rng = np.random.default_rng(81)
    x = C.rand(rng, (64, 12, 11, 7), dt, 0.5, 1.5)
    dy = C.rand(rng, x.shape, dt, 0.5, 1.5)
    p = 0.3
    libs = [b200] + ([binding.Oracle("ref")] if binding.have_ref() else [])
    for lib in libs:
        r = lib.dropout(p, x, dy)
        scale = dt(1) / (dt(1) - dt(p) + dt(1e-5))
        keep = r["y"] != 0
        assert np.allclose(r["y"][keep], (x * scale)[keep], rtol=1e-6 if dt == np.float32 else 1e-14)
        assert np.array_equal(r["dx"] != 0, keep), "backward used a different mask"
        assert np.allclose(r["dx"][keep], (dy * scale)[keep], rtol=1e-6 if dt == np.float32 else 1e-14)
        rate = 1.0 - keep.mean()
        assert abs(rate - p) < 5 * np.sqrt(p * (1 - p) / x.size), rate
        assert np.array_equal(r["y_infer"], x)


@pytest.mark.parametrize("dt", [np.float32])
def test_host_provider_path_through_the_input_feed(dt):
    """CATTL3_DEVICE_DATASET_GB=0 keeps the data set on the host: mini-batches then reach the device through the input
    feed (pinned staging + copy stream ring).  Same training run, same parameters as with the HBM-resident data set."""
    code = (
        "import sys, numpy as np; sys.path[:0] = [%r, %r]\n"
        "import cases as C; from oracle import binding\n"
        "lib = binding.Oracle('ref', path=%r); dt = np.%s\n"
        "x, obj = C.resnet_inputs(dt, total=192); n = lib.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)\n"
        "p, l, _ = lib.train_resnet(x, obj, 32, 2, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))\n"
        "np.save(sys.argv[1], p)\n" % (ROOT, os.path.join(ROOT, "tests"), SHIM, np.dtype(dt).name))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(({}, {"CATTL3_DEVICE_DATASET_GB": "0"}, {"CATTL3_NO_ARENA": "1"})):
            path = os.path.join(d, "p%d.npy" % i)
            r = subprocess.run([os.sys.executable, "-c", code, path], env=dict(os.environ, **env), capture_output=True,
                               text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1]), C.relerr(outs[0], outs[1])      # where the data comes from changes nothing
    assert np.array_equal(outs[0], outs[2]), C.relerr(outs[0], outs[2])      # nor does the parameter arena


def test_reference_gradient_test_passes():
    """The reference's own gradient_test.cpp, compiled unchanged against the B200 headers."""
    if not os.path.exists(GRADIENT_TEST):
        pytest.fail("%s missing: run __graft_entry__.build() where /root/reference exists" % GRADIENT_TEST)
    r = subprocess.run([GRADIENT_TEST, "--gtest_filter=" + GTEST_FILTER], capture_output=True, text=True,
                       timeout=1500)
    tail = "\n".join(r.stdout.splitlines()[-25:])
    print(tail)
    assert r.returncode == 0, tail + "\n" + r.stderr[-2000:]
    assert "[  PASSED  ] 13 tests" in r.stdout, tail
