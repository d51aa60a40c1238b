"""ctypes binding for the oracle libraries -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two libraries share one call signature per operation:

* ``orc``  = oracle/libcattl3_oracle.so, the plain-C restatement (oracle/cattl3_oracle_impl.h);
* ``ref``  = oracle/_ref/libcattle_ref.so, the unmodified reference behind oracle/ref_shim.cpp
  (only present where it was built from /root/reference; it travels to the GPU box as a binary).

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs
import this module.  All arrays are numpy arrays in the reference layout: Fortran (column-major)
order with the batch dimension first, so ``x[n, h, w, c]`` has N fastest in memory.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORC_PATH = os.path.join(HERE, "libcattl3_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libcattle_ref.so")


class Geom(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int) for k in
                ("n", "h", "w", "c", "f", "rh", "rw", "ph", "pw", "sh", "sw", "dh", "dw")]


def build(verbose=False):
    """Compile the C restatement and, where /root/reference exists, the reference shim."""
    subprocess.run(["make", "-f", os.path.join(HERE, "Makefile"), "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_SUF = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}
_CT = {np.dtype(np.float32): ctypes.c_float, np.dtype(np.float64): ctypes.c_double}


def _ptr(a):
    if a is None:
        return None
    assert a.flags["F_CONTIGUOUS"] or a.ndim <= 1, "oracle arrays must be column-major"
    return a.ctypes.data_as(ctypes.c_void_p)


def conv_out_dims(g, transposed=False):
    if not transposed:
        oh = (g.h - g.rh - (g.rh - 1) * g.dh + 2 * g.ph) // g.sh + 1
        ow = (g.w - g.rw - (g.rw - 1) * g.dw + 2 * g.pw) // g.sw + 1
    else:
        oh = (g.h - 1) * g.sh + g.rh + (g.rh - 1) * g.dh - 2 * g.ph
        ow = (g.w - 1) * g.sw + g.rw + (g.rw - 1) * g.dw - 2 * g.pw
    return oh, ow


def F(shape, dtype):
    return np.zeros(shape, dtype=dtype, order="F")


class Oracle:
    """One of the two libraries (prefix 'orc' or 'ref') behind a numpy API."""

    def __init__(self, prefix, path=None):
        """``path`` overrides the library: tests/cpp builds oracle/ref_shim.cpp a second time against the
        B200 headers (same ``ref_*`` entry points, product arithmetic) for the drop-in tests."""
        self.prefix = prefix
        if path is None:
            path = ORC_PATH if prefix == "orc" else REF_PATH
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = ctypes.CDLL(path)

    def _fn(self, name, dtype):
        fn = getattr(self.lib, "%s_%s_%s" % (self.prefix, name, _SUF[np.dtype(dtype)]))
        fn.restype = ctypes.c_int
        return fn

    def num_threads(self):
        if self.prefix == "ref":
            return int(self.lib.ref_num_threads())
        return os.cpu_count()

    def conv(self, g, x, w, b, dy=None, transposed=False, want_dx=True, back_reps=1):
        """Returns dict(y, dx, dw, db, times_ms)."""
        dt = x.dtype
        oh, ow = conv_out_dims(g, transposed)
        y = F((g.n, oh, ow, g.f), dt)
        dx = F(x.shape, dt) if (dy is not None and want_dx) else None
        dw = F(w.shape, dt) if dy is not None else None
        db = F(b.shape, dt) if dy is not None else None
        times = (ctypes.c_double * 2)()
        rc = self._fn("conv", dt)(ctypes.byref(g), ctypes.c_int(int(transposed)), _ptr(x), _ptr(w),
                                  _ptr(b), _ptr(dy), _ptr(y), _ptr(dx), _ptr(dw), _ptr(db),
                                  ctypes.c_int(back_reps), times)
        assert rc == 0, rc
        return dict(y=y, dx=dx, dw=dw, db=db, times_ms=(times[0], times[1]))

    def dense(self, x, w, b, dy=None, want_dx=True, back_reps=1):
        dt = x.dtype
        n, i = x.shape
        o = w.shape[1]
        y = F((n, o), dt)
        dx = F(x.shape, dt) if (dy is not None and want_dx) else None
        dw = F(w.shape, dt) if dy is not None else None
        db = F(b.shape, dt) if dy is not None else None
        times = (ctypes.c_double * 2)()
        rc = self._fn("dense", dt)(n, i, o, _ptr(x), _ptr(w), _ptr(b), _ptr(dy), _ptr(y), _ptr(dx),
                                   _ptr(dw), _ptr(db), back_reps, times)
        assert rc == 0, rc
        return dict(y=y, dx=dx, dw=dw, db=db, times_ms=(times[0], times[1]))

    def activation(self, kind, alpha, x, dy=None):
        dt = x.dtype
        n = x.shape[0]
        vol = x.size // n
        y = F(x.shape, dt)
        dx = F(x.shape, dt) if dy is not None else None
        rc = self._fn("activation", dt)(kind, _CT[np.dtype(dt)](alpha), n, vol, _ptr(x), _ptr(dy),
                                        _ptr(y), _ptr(dx))
        assert rc == 0, rc
        return dict(y=y, dx=dx)

    def pool(self, kind, x, rh, rw, sh, sw, dy=None):
        dt = x.dtype
        n, h, w, c = x.shape
        oh, ow = (h - rh) // sh + 1, (w - rw) // sw + 1
        y = F((n, oh, ow, c), dt)
        dx = F(x.shape, dt) if dy is not None else None
        times = (ctypes.c_double * 2)()
        rc = self._fn("pool", dt)(kind, n, h, w, c, rh, rw, sh, sw, _ptr(x), _ptr(dy), _ptr(y),
                                  _ptr(dx), times)
        assert rc == 0, rc
        return dict(y=y, dx=dx, times_ms=(times[0], times[1]))

    def batchnorm(self, per_channel, xs, gamma, beta, dy=None, decay=0.1, eps=1e-5, want_dx=True):
        """xs: (steps, n, h, w, c) stacked so that each step is a column-major block."""
        dt = xs[0].dtype
        steps = len(xs)
        n, h, w, c = xs[0].shape
        xcat = np.concatenate([a.ravel(order="F") for a in xs])
        groups = c if per_channel else h * w * c
        y, yi = F(xs[0].shape, dt), F(xs[0].shape, dt)
        dx = F(xs[0].shape, dt) if (dy is not None and want_dx) else None
        dgamma = np.zeros(groups, dt) if dy is not None else None
        dbeta = np.zeros(groups, dt) if dy is not None else None
        rm, rs = np.zeros(groups, dt), np.zeros(groups, dt)
        ct = _CT[np.dtype(dt)]
        rc = self._fn("batchnorm", dt)(int(per_channel), n, h, w, c, ct(decay), ct(eps), steps,
                                       _ptr(xcat), _ptr(gamma), _ptr(beta), _ptr(dy), _ptr(y), _ptr(dx),
                                       _ptr(dgamma), _ptr(dbeta), _ptr(rm), _ptr(rs), _ptr(yi))
        assert rc == 0, rc
        return dict(y=y, dx=dx, dgamma=dgamma, dbeta=dbeta, run_mean=rm, run_inv_sd=rs, y_infer=yi)

    def optimizer(self, kind, hyper, l2_lambda, p0, grads, steps_per_epoch, hostparams=False):
        """p0: rows x cols; grads: list of rows x cols arrays (one per step)."""
        dt = p0.dtype
        rows, cols = p0.shape
        steps = len(grads)
        gcat = np.concatenate([g.ravel(order="F") for g in grads])
        hy = np.asarray(hyper, dtype=dt)
        out = F(p0.shape, dt)
        rc = self._fn("optimizer_hostparams" if hostparams else "optimizer", dt)(kind, _ptr(hy), _CT[np.dtype(dt)](l2_lambda), rows, cols, steps,
                                       steps_per_epoch, _ptr(p0), _ptr(gcat), _ptr(out))
        assert rc == 0, rc
        return out

    def train_cifar(self, x, obj, batch, epochs, params_in=None, n_params=26968):
        """Reference only: config-1 network, Nadam; returns (params, loss, train_ms)."""
        dt = x.dtype
        total = x.shape[0]
        out = np.zeros(n_params, dt)
        loss, ms = ctypes.c_double(), ctypes.c_double()
        rc = self._fn("train_cifar", dt)(total, batch, epochs, _ptr(x), _ptr(obj), _ptr(params_in),
                                         _ptr(out), ctypes.byref(loss), ctypes.byref(ms))
        assert rc == 0, rc
        return out, loss.value, ms.value


    def dropout(self, prob, x, dy):
        """DropoutLayer: training forward + backward, inference forward.  Returns dict(y, dx, y_infer)."""
        dt = x.dtype
        n, h, w, c = x.shape
        y, dx, yi = F(x.shape, dt), F(x.shape, dt), F(x.shape, dt)
        rc = self._fn("dropout", dt)(n, h, w, c, _CT[np.dtype(dt)](prob), _ptr(x), _ptr(dy), _ptr(y), _ptr(dx), _ptr(yi))
        assert rc == 0, rc
        return dict(y=y, dx=dx, y_infer=yi)

    def train_autoencoder(self, x, batch, epochs, params_in=None):
        """Config-3 auto-encoder (mnist_autoencoder.cpp), SquaredLoss against the input, Nadam.
        epochs < 0: only returns the parameter count.  Returns (params, loss, train_ms)."""
        dt = x.dtype
        n = ctypes.c_int()
        fn = self._fn("train_autoencoder", dt)
        rc = fn(0, 1, -1, None, None, None, ctypes.byref(n), None, None)
        assert rc == 0, rc
        if epochs < 0:
            return n.value
        out = np.zeros(n.value, dt)
        loss, ms = ctypes.c_double(), ctypes.c_double()
        rc = fn(x.shape[0], batch, epochs, _ptr(x), _ptr(params_in), _ptr(out), ctypes.byref(n), ctypes.byref(loss),
                ctypes.byref(ms))
        assert rc == 0, rc
        return out, loss.value, ms.value

    def train_resnet(self, x, obj, batch, epochs, arch, params_in=None):
        """Config-4 ResNet-style network; arch = (stem_r, stem_s, stem_pool, width, blocks, head_pool).
        x: total x h x w x c, obj: total x 1 x 1 x classes.  epochs < 0: parameter count only."""
        dt = x.dtype
        total, h, w, c = x.shape
        classes = obj.shape[-1]
        n = ctypes.c_int()
        fn = self._fn("train_resnet", dt)
        rc = fn(0, 1, -1, h, w, c, *arch, classes, None, None, None, None, ctypes.byref(n), None, None)
        assert rc == 0, rc
        if epochs < 0:
            return n.value
        out = np.zeros(n.value, dt)
        loss, ms = ctypes.c_double(), ctypes.c_double()
        rc = fn(total, batch, epochs, h, w, c, *arch, classes, _ptr(x), _ptr(obj), _ptr(params_in), _ptr(out),
                ctypes.byref(n), ctypes.byref(loss), ctypes.byref(ms))
        assert rc == 0, rc
        return out, loss.value, ms.value

    def train_seqnet(self, x, obj, batch, epochs, width, state, params_in=None):
        """Config-5 composite network over frame sequences (SequentialNeuralNetwork{Parallel conv lanes, DenseNet
        modules, MaxPool} -> convolutional LSTMNeuralNetwork), SquaredLoss, Nadam.
        x: total x seq x hw x hw x 3, obj: total x 1 x hw/2 x hw/2 x state.  epochs < 0: parameter count only."""
        dt = x.dtype
        total, seq, hw = x.shape[:3]
        n = ctypes.c_int()
        fn = self._fn("train_seqnet", dt)
        rc = fn(0, 1, 1, -1, hw, width, state, None, None, None, None, ctypes.byref(n), None, None)
        assert rc == 0, rc
        if epochs < 0:
            return n.value
        out = np.zeros(n.value, dt)
        loss, ms = ctypes.c_double(), ctypes.c_double()
        rc = fn(total, seq, batch, epochs, hw, width, state, _ptr(x), _ptr(obj), _ptr(params_in), _ptr(out),
                ctypes.byref(n), ctypes.byref(loss), ctypes.byref(ms))
        assert rc == 0, rc
        return out, loss.value, ms.value


def have_ref():
    return os.path.exists(REF_PATH)
