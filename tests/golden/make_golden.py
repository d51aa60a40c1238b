"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcattle_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture stores the seeded inputs and the reference's outputs for one operation family on the
small cases of tests/cases.py, for float32 and float64.  The reference pins no golden values of its
own, so these files are the committed record of what the reference computes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.binding import Oracle  # noqa: E402
import cases as C  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_CONV = ["gt_rank1", "gt_rank2", "gt_rank3", "gt_second", "mnist_conv0", "ragged", "single"]
GOLDEN_TCONV = ["gt_rank1", "gt_rank2", "gt_rank3", "mnist_t0", "ragged"]
GOLDEN_DENSE = ["gt_rank1", "cifar_fc1", "ragged", "single"]
DT = {"f32": np.float32, "f64": np.float64}


def main():
    ref = Oracle("ref")
    out = {}
    for suf, dt in DT.items():
        for tr, names, table in ((False, GOLDEN_CONV, C.CONV_CASES), (True, GOLDEN_TCONV, C.TCONV_CASES)):
            for name in names:
                g, x, w, b, dy = C.conv_inputs(table[name], dt, 11, tr)
                r = ref.conv(g, x, w, b, dy, transposed=tr, back_reps=2)
                key = "%s/%s/%s" % ("tconv" if tr else "conv", name, suf)
                for k, v in dict(x=x, w=w, b=b, dy=dy, y=r["y"], dx=r["dx"], dw=r["dw"], db=r["db"]).items():
                    out[key + "/" + k] = v
        rng = np.random.default_rng(12)
        for name in GOLDEN_DENSE:
            n, i, o = C.DENSE_CASES[name]
            x, w, b, dy = C.rand(rng, (n, i), dt), C.rand(rng, (i, o), dt), C.rand(rng, (1, o), dt), C.rand(rng, (n, o), dt)
            r = ref.dense(x, w, b, dy, back_reps=2)
            for k, v in dict(x=x, w=w, b=b, dy=dy, y=r["y"], dx=r["dx"], dw=r["dw"], db=r["db"]).items():
                out["dense/%s/%s/%s" % (name, suf, k)] = v
        for name, (kind, alpha) in C.ACT_CASES.items():
            x = C.rand(rng, (5, 3, 2, 2), dt, -2, 2)
            x.ravel(order="K")[:3] = 0  # exercise the x == 0 branch (ReLU'(0) = 1 in the reference)
            dy = C.rand(rng, x.shape, dt)
            r = ref.activation(kind, alpha, x, dy)
            for k, v in dict(x=x, dy=dy, y=r["y"], dx=r["dx"]).items():
                out["act/%s/%s/%s" % (name, suf, k)] = v
        for name, (kind, n, h, w, c, rh, rw, sh, sw) in C.POOL_CASES.items():
            if n * h * w * c > 4096:
                continue
            x = np.asfortranarray(np.round(C.rand(rng, (n, h, w, c), dt) * 4) / 4)  # deliberate ties
            oh, ow = (h - rh) // sh + 1, (w - rw) // sw + 1
            dy = C.rand(rng, (n, oh, ow, c), dt)
            r = ref.pool(kind, x, rh, rw, sh, sw, dy)
            for k, v in dict(x=x, dy=dy, y=r["y"], dx=r["dx"]).items():
                out["pool/%s/%s/%s" % (name, suf, k)] = v
        for name, (pc, n, h, w, c, steps) in C.BN_CASES.items():
            if n * h * w * c > 4096:
                continue
            xs = [C.rand(rng, (n, h, w, c), dt, -1, 2) for _ in range(steps)]
            G = c if pc else h * w * c
            gm, bt, dy = C.rand(rng, (G,), dt, 0.5, 1.5), C.rand(rng, (G,), dt), C.rand(rng, (n, h, w, c), dt)
            r = ref.batchnorm(pc, xs, gm, bt, dy)
            d = dict(gamma=gm, beta=bt, dy=dy, **{k: r[k] for k in
                     ("y", "dx", "dgamma", "dbeta", "run_mean", "run_inv_sd", "y_infer")})
            for s, xv in enumerate(xs):
                d["x%d" % s] = xv
            for k, v in d.items():
                out["bn/%s/%s/%s" % (name, suf, k)] = v
        for name, (kind, hy) in C.OPT_CASES.items():
            p0 = C.rand(rng, (7, 5), dt)
            grads = [C.rand(rng, (7, 5), dt) for _ in range(6)]
            for lam in (0.0, 0.01):
                r = ref.optimizer(kind, hy, lam, p0, grads, 3)
                key = "opt/%s_l2_%g/%s" % (name, lam, suf)
                out[key + "/p0"] = p0
                out[key + "/grads"] = np.stack(grads)
                out[key + "/p"] = r
    # config 1 (BASELINE.json configs[0]): 2 Nadam steps of the cifar ConvNet at batch 4
    for suf, dt in DT.items():
        rng = np.random.default_rng(1001)
        x = C.rand(rng, (8, 32, 32, 3), dt)
        obj = np.zeros((8, 1, 1, 10), dtype=dt, order="F")
        for i in range(8):
            obj[i, 0, 0, i % 10] = 1
        p0, _, _ = ref.train_cifar(x, obj, 4, 0)
        p1, loss, _ = ref.train_cifar(x, obj, 4, 1, params_in=p0)
        out["cifar/%s/p0" % suf] = p0
        out["cifar/%s/p1" % suf] = p1
        out["cifar/%s/loss" % suf] = np.array([loss])
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote %d arrays, %.1f KB" % (len(out), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")) / 1024))


if __name__ == "__main__":
    main()
