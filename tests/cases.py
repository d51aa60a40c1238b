"""Shared parity-test cases (geometry lists follow the reference's own gradient tests,
/root/reference/test/gradient_test.cpp:143-369, plus the BASELINE.json configs at reduced batch)."""
import numpy as np

# (n, h, w, c, f, rh, rw, ph, pw, sh, sw, dh, dw)
CONV_CASES = {
    # test/gradient_test.cpp:168-180 -- rank-1 {32} viewed as 32x1x1, F=5, R=3, pad 2
    "gt_rank1": (5, 32, 1, 1, 5, 3, 1, 2, 0, 1, 1, 1, 0),
    # rank-2 {10,10} viewed as 10x10x1: F=5, R=3x2, pad 1x2, stride 1x2, dilation 1x0
    "gt_rank2": (5, 10, 10, 1, 5, 3, 2, 1, 2, 1, 2, 1, 0),
    # rank-3 {8,8,2}, same hyper-parameters
    "gt_rank3": (5, 8, 8, 2, 5, 3, 2, 1, 2, 1, 2, 1, 0),
    # second layer of the same tests: F=1, R=2x2 (defaults otherwise: pad 1, stride 1)
    "gt_second": (5, 6, 5, 5, 1, 2, 2, 1, 1, 1, 1, 0, 0),
    # config 1 (examples/cifar_convnet.cpp:25,28) at batch 8
    "cifar_conv0": (8, 32, 32, 3, 8, 3, 3, 1, 1, 1, 1, 0, 0),
    "cifar_conv1": (8, 16, 16, 8, 8, 3, 3, 1, 1, 1, 1, 0, 0),
    # config 3 (examples/mnist_autoencoder.cpp:27,29): 4x4 stride 2 / 4x4, no padding given -> pad 1
    "mnist_conv0": (8, 28, 28, 1, 3, 4, 4, 1, 1, 2, 2, 0, 0),
    # config 2 geometry at reduced size: 3x3 pad 1, TMA-friendly batch (n % 32 == 0, c % 32 == 0)
    "c2_small": (32, 12, 10, 64, 32, 3, 3, 1, 1, 1, 1, 0, 0),
    "c2_small_f256": (64, 7, 6, 64, 256, 3, 3, 1, 1, 1, 1, 0, 0),
    "c2_stride2": (32, 12, 10, 32, 48, 3, 3, 1, 1, 2, 2, 0, 0),
    "c2_dil1": (32, 12, 10, 32, 16, 3, 3, 2, 2, 1, 1, 1, 1),
    "c2_1x1": (64, 9, 9, 64, 64, 1, 1, 0, 0, 1, 1, 0, 0),
    "ragged_c": (32, 9, 7, 20, 24, 3, 3, 1, 1, 1, 1, 0, 0),
    # batch 256, few filters either way: the row-sharing gather GEMM (three output rows per tile) runs the forward pass
    # of "rows64" / "rows_2x3" and the input gradient of all three; 7 and 5 rows leave a ragged last row group
    "rows64": (256, 7, 6, 64, 64, 3, 3, 1, 1, 1, 1, 0, 0),
    "rows_f256": (256, 5, 4, 48, 256, 3, 3, 1, 1, 1, 1, 0, 0),
    "rows_2x3": (256, 6, 5, 64, 32, 2, 3, 0, 1, 1, 1, 0, 0),
    # strided input gradients on the tensor-core path (one stride-1 sub-problem per residue class of the input pixel)
    "s2_d1": (32, 11, 9, 32, 32, 3, 3, 2, 2, 2, 2, 1, 1),    # classes without taps: zero-filled
    "s3": (32, 13, 11, 16, 32, 3, 3, 1, 1, 3, 3, 0, 0),
    "s21": (32, 9, 8, 32, 16, 3, 2, 1, 0, 2, 1, 0, 0),
    "s2_1x1": (32, 8, 8, 32, 32, 1, 1, 0, 0, 2, 2, 0, 0),
    "s2_4x4": (32, 10, 10, 16, 16, 4, 4, 1, 1, 2, 2, 0, 0),
    "s2_ragged_edge": (64, 12, 9, 16, 24, 3, 3, 0, 1, 2, 2, 0, 0),
    # double at GEMM-sized filter counts (conv_dfma.cu): odd batch (scalar stores), 64- and 128-wide filter tiles
    "dfma_odd": (5, 9, 7, 6, 40, 3, 3, 1, 1, 2, 1, 0, 1),
    "dfma_mid": (6, 20, 18, 8, 72, 3, 3, 1, 1, 1, 1, 0, 0),
    "dfma_wide": (4, 10, 9, 20, 136, 3, 2, 1, 0, 1, 1, 0, 0),
    # config 4's stem at test size: 7x7 stride 2 over 3 channels (float: flattened (tap, channel) reduction)
    "stem": (32, 20, 18, 3, 64, 7, 7, 3, 3, 2, 2, 0, 0),
    "stem_wide": (8, 12, 12, 3, 80, 5, 5, 2, 2, 1, 1, 0, 0),
    # ragged everything: odd batch, odd channels
    "ragged": (7, 9, 11, 5, 3, 3, 4, 2, 1, 2, 1, 0, 1),
    "single": (1, 3, 3, 1, 1, 3, 3, 1, 1, 1, 1, 0, 0),
}

TCONV_CASES = {
    # test/gradient_test.cpp:192-204
    "gt_rank1": (5, 16, 1, 1, 5, 4, 1, 1, 0, 1, 1, 0, 0),
    "gt_rank2": (5, 4, 5, 1, 5, 5, 3, 1, 0, 1, 2, 0, 1),
    "gt_rank3": (5, 2, 3, 2, 5, 5, 3, 1, 0, 1, 2, 0, 1),
    # config 3 (examples/mnist_autoencoder.cpp:43,45) at batch 8
    "mnist_t0": (8, 10, 10, 3, 3, 4, 4, 1, 1, 1, 1, 0, 0),
    "mnist_t1": (8, 13, 13, 3, 1, 4, 4, 1, 1, 2, 2, 0, 0),
    "wide": (32, 6, 5, 32, 16, 3, 3, 1, 1, 2, 2, 0, 0),
    "wide_s2_k4": (32, 5, 6, 16, 16, 4, 4, 1, 1, 2, 2, 0, 0),
    "wide_s3_k2": (32, 4, 4, 16, 16, 2, 2, 0, 0, 3, 3, 0, 0),   # output pixels that receive the bias only
    "wide_s1": (32, 6, 5, 16, 32, 3, 3, 1, 1, 1, 1, 0, 0),
    "ragged": (3, 4, 3, 5, 2, 2, 3, 0, 1, 2, 1, 1, 0),
    "dfma_t": (8, 9, 8, 24, 48, 3, 3, 1, 1, 2, 2, 0, 0),
}

# double, sized so that the producer-warp DMMA kernels run (conv_dmma.cu version 2: the grid must fill the device; weight
# gradient: even batch, M >= 16384): ragged rows / channels / filters, 64- and 128-wide column tiles, every weight-gradient
# tile shape (256 x 64, 128 x 128, 64 x 256), an odd batch (8-byte copies), a strided input gradient (den = 2)
DMMA2_CASES = {
    "d2_ragged": (64, 29, 27, 36, 40, 3, 3, 1, 1, 1, 1, 0, 0),
    "d2_wide": (34, 30, 28, 9, 96, 3, 3, 1, 1, 1, 1, 0, 0),
    "d2_odd": (65, 28, 28, 40, 72, 3, 3, 1, 1, 1, 1, 0, 0),
    "d2_s2": (32, 57, 55, 40, 130, 3, 3, 1, 1, 2, 2, 0, 0),
    "d2_dil": (64, 30, 30, 48, 64, 3, 3, 2, 2, 1, 1, 1, 1),
    "d2_1x1": (64, 28, 28, 64, 256, 1, 1, 0, 0, 1, 1, 0, 0),
}
DMMA2_TCONV_CASES = {
    "d2_t_s2": (64, 20, 20, 40, 40, 3, 3, 1, 1, 2, 2, 0, 0),
    "d2_t_s1": (64, 28, 28, 40, 40, 2, 2, 0, 0, 1, 1, 0, 0),
}

# (n, in, out)
DENSE_CASES = {
    "gt_rank1": (5, 32, 16),   # test/gradient_test.cpp:143-156
    "gt_rank3": (5, 32, 16),
    "cifar_fc0": (64, 512, 50),  # examples/cifar_convnet.cpp:32,35
    "cifar_fc1": (64, 50, 10),
    "mnist_fc": (32, 300, 100),
    "ragged": (7, 13, 3),
    "single": (1, 1, 1),
    "head": (64, 3136, 10),     # a classifier head: small batch, long reduction, few outputs (split-K forward)
    "head_ragged": (5, 1030, 3),
}

# (kind, alpha) -- kinds as CATTL3_ACT_*; test/gradient_test.cpp:215-282 uses 0.2 / 0.2 / 1.2
ACT_CASES = {
    "relu": (0, 0.0), "leaky": (1, 0.2), "elu": (2, 0.2), "swish": (3, 1.2),
    "sigmoid": (4, 0.0), "tanh": (5, 0.0), "softplus": (6, 0.0), "softmax": (7, 0.0),
}

# (kind, n, h, w, c, rh, rw, sh, sw); test/gradient_test.cpp:292-315
POOL_CASES = {
    "max_1d_r2s2": (0, 5, 16, 1, 1, 2, 1, 2, 1),
    "max_1d_r3s1": (0, 5, 16, 1, 1, 3, 1, 1, 1),
    "max_2d_overlap": (0, 5, 8, 8, 2, 3, 2, 1, 2),
    "max_cifar": (0, 8, 32, 32, 8, 2, 2, 2, 2),
    "mean_1d_r2s2": (1, 5, 16, 1, 1, 2, 1, 2, 1),
    "mean_2d_overlap": (1, 5, 8, 8, 2, 3, 2, 1, 2),
    "mean_cifar": (1, 8, 16, 16, 8, 2, 2, 2, 2),
    # stride > window: input rows / columns no window covers, ragged trailing edge; vector and scalar batch
    "max_gap": (0, 8, 9, 7, 3, 2, 2, 3, 2),
    "max_gap_scalar": (0, 5, 10, 7, 2, 2, 3, 3, 3),
    "mean_gap": (1, 6, 9, 7, 3, 2, 2, 3, 3),
    "max_vec_overlap": (0, 8, 8, 8, 2, 3, 2, 1, 2),
    "mean_vec_overlap": (1, 12, 7, 8, 2, 3, 3, 2, 1),
}

# (per_channel, n, h, w, c, steps); test/gradient_test.cpp:346-369
BN_CASES = {
    "pc_rank3": (1, 5, 4, 4, 3, 1),
    "pc_running": (1, 8, 5, 3, 4, 3),
    "pa_rank3": (0, 5, 4, 4, 2, 1),
    "pa_running": (0, 6, 3, 2, 2, 3),
    "pc_wide": (1, 32, 6, 6, 16, 2),
}

# kind -> hyper {lr, a, b, eps} (reference defaults)
OPT_CASES = {
    "sgd": (0, (1e-3, 0, 0, 0)),
    "momentum": (1, (1e-3, 1e-3, 0.9, 0)),
    "nesterov": (2, (1e-3, 1e-3, 0.9, 0)),
    "adagrad": (3, (1e-2, 0, 0, 1e-5)),
    "rmsprop": (4, (1e-3, 0, 1e-1, 1e-5)),
    "adadelta": (5, (0, 5e-2, 0, 1e-5)),
    "adam": (6, (1e-3, 1e-1, 1e-3, 1e-5)),
    "adamax": (7, (1e-3, 1e-1, 1e-3, 1e-5)),
    "nadam": (8, (1e-3, 1e-1, 1e-3, 1e-5)),
    "amsgrad": (9, (1e-3, 1e-1, 1e-3, 1e-5)),
}


def rand(rng, shape, dtype, lo=-1.0, hi=1.0):
    return np.asfortranarray(rng.uniform(lo, hi, size=shape).astype(dtype))


def relerr(a, b):
    """Norm-relative error max|a-b| / max(max|b|, tiny): the metric SURVEY.md section 8c prescribes."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


TOL = {np.dtype(np.float32): 1e-4, np.dtype(np.float64): 1e-10}  # BASELINE.json north_star


def conv_inputs(case, dtype, seed, transposed=False):
    from oracle.binding import Geom, conv_out_dims
    g = Geom(*case)
    rng = np.random.default_rng(seed)
    x = rand(rng, (g.n, g.h, g.w, g.c), dtype)
    oh, ow = conv_out_dims(g, transposed)
    if transposed:
        w = rand(rng, (g.c, g.rh * g.rw * g.f), dtype, -0.5, 0.5)
        b = rand(rng, (1, oh * ow * g.f), dtype)
    else:
        w = rand(rng, (g.rh * g.rw * g.c, g.f), dtype, -0.5, 0.5)
        b = rand(rng, (1, g.f), dtype)
    dy = rand(rng, (g.n, oh, ow, g.f), dtype)
    return g, x, w, b, dy


# ---- whole-network training cases (oracle/ref_shim.cpp: ref_train_autoencoder / ref_train_resnet) ----
RESNET_SMALL = (3, 1, 0, 16, 2, 2)   # stem 3x3 stride 1, no stem pool, width 16, 2 residual modules, head mean-pool 2x2


def autoencoder_inputs(dtype, total=8, seed=3001):
    """BASELINE.json configs[2] at test size: synthetic 28x28x1 uniform [0, 1)."""
    rng = np.random.default_rng(seed)
    return rand(rng, (total, 28, 28, 1), dtype, 0.0, 1.0)


def resnet_inputs(dtype, total=64, hw=8, seed=4001):
    """BASELINE.json configs[3] at test size: synthetic hw x hw x 3, one-hot objectives i mod 10."""
    rng = np.random.default_rng(seed)
    x = rand(rng, (total, hw, hw, 3), dtype)
    obj = np.zeros((total, 1, 1, 10), dtype=dtype, order="F")
    obj[np.arange(total), 0, 0, np.arange(total) % 10] = 1
    return x, obj


def seeded_params(n, dtype, seed):
    """Deterministic starting parameters (so that fixtures need not store them): uniform +-0.15."""
    return np.random.default_rng(seed).uniform(-0.15, 0.15, n).astype(dtype)


def seqnet_inputs(dtype, total=16, seq=3, hw=8, state=4, seed=5001):
    """BASELINE.json configs[4] at test size: synthetic sequences of hw x hw x 3 frames, uniform [-1, 1); the objective
    is one hw/2 x hw/2 x state frame per sequence, uniform [-0.5, 0.5) (inside the range of the LSTM's output)."""
    rng = np.random.default_rng(seed)
    x = rand(rng, (total, seq, hw, hw, 3), dtype)
    obj = rand(rng, (total, 1, hw // 2, hw // 2, state), dtype, -0.5, 0.5)
    return x, obj


SEQNET_SMALL = dict(width=4, state=4)


# REF_SHIM_CONSTRAINTS: value clip / Frobenius ("L1") / squared-norm ("L2") limits and the same three on the gradient, on every
# weight matrix of the config-1 network; chosen so that each of the six is active on the seeded parameters and gradients
CIFAR_CON = "0.05,0.8,0.5,0.00003,0.0003,0.0000001"
CIFAR_REG = "en:0.003,0.01"   # REF_SHIM_REG: ElasticNet (l1, l2) on every weight matrix of the config-1 network


def cifar_inputs(dtype, total=64, seed=1004):
    """BASELINE.json configs[0] at test size: synthetic 32x32x3 uniform [-1, 1), one-hot objectives i mod 10."""
    rng = np.random.default_rng(seed)
    x = rand(rng, (total, 32, 32, 3), dtype)
    obj = np.zeros((total, 1, 1, 10), dtype=dtype, order="F")
    obj[np.arange(total), 0, 0, np.arange(total) % 10] = 1
    return x, obj
