"""Python-side loader for libcattl3_b200.so (the C ABI of include/cattl3_b200.h).

The product is the shared library plus the header-only C++ classes in ``cattle/``; this module only
exists so that the Python test-suite and ``bench.py`` can drive the same C ABI through ctypes, with
PyTorch used purely for device memory, streams and ``torch.distributed`` plumbing.

The directory name contains a hyphen, so import it through ``load_package()`` in
``__graft_entry__.py`` (``importlib`` by path, module name ``cattl3_b200``).

There is no CPU fallback and no alternative implementation: if the library is missing or a call
fails, an exception is raised.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcattl3_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(HERE), "include")
CATTLE_INCLUDE_DIR = os.path.join(HERE, "cattle")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = 0, 1, 2, 3, 4
ACT = dict(relu=0, leaky_relu=1, elu=2, swish=3, sigmoid=4, tanh=5, softplus=6, softmax=7)
ACT_NONE = -1
POOL = dict(max=0, mean=1)
LOSS = dict(squared=0, cross_entropy=1)
OPT = dict(sgd=0, momentum=1, nesterov=2, adagrad=3, rmsprop=4, adadelta=5, adam=6, adamax=7, nadam=8,
           amsgrad=9)
PATH_AUTO, PATH_SIMT, PATH_TCGEN05, PATH_FMA = 0, 1, 2, 3


class Cattl3Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("cattl3 error %d: %s" % (code, message))
        self.code = code


class ConvGeom(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in
                ("n", "h", "w", "c", "f", "rh", "rw", "ph", "pw", "sh", "sw", "dh", "dw")]


class PoolGeom(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("n", "h", "w", "c", "rh", "rw", "sh", "sw")]


class Epilogue(ctypes.Structure):
    """cattl3_epilogue: what a kernel layer's forward pass fuses besides the bias."""
    _fields_ = [("act_kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("act_param", ctypes.c_double),
                ("act_out", ctypes.c_void_p), ("col_stats", ctypes.c_void_p)]


class OptStep(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reset_grad", ctypes.c_int32)] + \
               [(k, ctypes.c_double) for k in ("lr", "a", "b", "eps", "lr_epoch", "c1", "c1n", "c2", "l2_lambda")]


def build(verbose=False):
    """Compile every CUDA translation unit for sm_100a into libcattl3_b200.so (in-tree)."""
    subprocess.run(["make", "-C", HERE, "-j8", "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_lib = None

# every symbol include/cattl3_b200.h declares (tests/test_abi.py checks the header against this)
_VOID_P = ctypes.c_void_p
_SYMBOLS = [
    "cattl3_abi_version", "cattl3_last_error", "cattl3_device_count", "cattl3_ctx_create",
    "cattl3_ctx_destroy", "cattl3_ctx_synchronize", "cattl3_ctx_throttle", "cattl3_ctx_set_conv_path", "cattl3_weights_stable_begin", "cattl3_weights_stable_end", "cattl3_ctx_launch_count",
    "cattl3_ctx_last_path", "cattl3_ctx_stream", "cattl3_malloc", "cattl3_free", "cattl3_memset",
    "cattl3_memcpy_h2d", "cattl3_memcpy_d2h", "cattl3_memcpy_d2d", "cattl3_memcpy_2d", "cattl3_host_alloc", "cattl3_host_free",
    "cattl3_conv_output_dims", "cattl3_pool_output_dims", "cattl3_feed_create", "cattl3_feed_destroy", "cattl3_feed_push",
    "cattl3_ctx_allocated_bytes", "cattl3_graph_begin", "cattl3_graph_end", "cattl3_graph_launch", "cattl3_graph_destroy",
    "cattl3_conv_forward_host_f32", "cattl3_conv_backward_host_f32",
    "cattl3_conv_forward_host_async_f32", "cattl3_conv_backward_host_async_f32", "cattl3_host_wait",
    "cattl3_comm_unique_id", "cattl3_comm_create", "cattl3_comm_create_from_env", "cattl3_comm_destroy",
    "cattl3_comm_world_size", "cattl3_comm_rank", "cattl3_comm_group_start", "cattl3_comm_group_end",
    "cattl3_comm_allreduce_sum_f32", "cattl3_comm_allreduce_sum_f64",
    "cattl3_comm_allreduce_sum_async_f32", "cattl3_comm_allreduce_sum_async_f64", "cattl3_comm_wait",
] + [n + s for s in ("_f32", "_f64") for n in (
    "cattl3_conv_forward", "cattl3_conv_backward", "cattl3_transconv_forward", "cattl3_transconv_backward",
    "cattl3_dense_forward", "cattl3_dense_backward", "cattl3_activation_forward", "cattl3_activation_backward",
    "cattl3_pool_forward", "cattl3_pool_backward", "cattl3_batchnorm_forward", "cattl3_batchnorm_backward",
    "cattl3_optimizer_step", "cattl3_optimizer_step_indirect", "cattl3_add_inplace", "cattl3_mul_inplace", "cattl3_muladd", "cattl3_regularize", "cattl3_constrain", "cattl3_scale", "cattl3_axpy",
    "cattl3_conv_forward_fused", "cattl3_dense_forward_fused", "cattl3_transconv_forward_fused", "cattl3_batchnorm_forward_stats",
    "cattl3_dropout_forward", "cattl3_dropout_backward", "cattl3_loss",
    "cattl3_batchnorm_stats", "cattl3_batchnorm_backward_sums", "cattl3_batchnorm_backward_apply",
    "cattl3_slice_rows", "cattl3_fill")]


def lib():
    """Loads the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError("%s is missing: run __graft_entry__.build() (nvcc, sm_100a); "
                                    "there is no CPU or PyTorch fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name in _SYMBOLS:
            getattr(L, name).restype = ctypes.c_int
        L.cattl3_last_error.restype = ctypes.c_char_p
        L.cattl3_ctx_last_path.restype = ctypes.c_char_p
        L.cattl3_ctx_last_path.argtypes = [_VOID_P]
        L.cattl3_ctx_launch_count.restype = ctypes.c_int64
        L.cattl3_ctx_launch_count.argtypes = [_VOID_P]
        L.cattl3_ctx_stream.restype = _VOID_P
        L.cattl3_ctx_stream.argtypes = [_VOID_P]
        _lib = L
    return _lib


def _p(t):
    """Device (or host) address of a torch tensor / numpy array / int / None."""
    if t is None:
        return None
    if isinstance(t, int):
        return ctypes.c_void_p(t)
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    if hasattr(t, "ctypes"):
        return ctypes.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _suffix(dtype):
    s = str(dtype)
    if s.endswith("float32"):
        return "f32", ctypes.c_float
    if s.endswith("float64"):
        return "f64", ctypes.c_double
    raise TypeError("cattl3 supports float32 and float64 only (the reference's Scalar types), got %s" % s)


class Context:
    """One cattl3_ctx: a device, a stream and its workspaces."""

    def __init__(self, device=0, stream=None):
        self.L = lib()
        self.h = ctypes.c_void_p()
        # stream None -> the context creates its own stream; 0 -> the legacy default stream, passed
        # as the explicit handle cudaStreamLegacy (0x1) because NULL means "create one" in the C ABI
        sp = None if stream is None else ctypes.c_void_p(stream if stream else 1)
        self._chk(self.L.cattl3_ctx_create(ctypes.byref(self.h), int(device), sp))

    def _chk(self, rc):
        if rc != OK:
            raise Cattl3Error(rc, (self.L.cattl3_last_error() or b"").decode())

    def close(self):
        if self.h:
            self.L.cattl3_ctx_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._chk(self.L.cattl3_ctx_synchronize(self.h))

    def set_conv_path(self, path):
        self._chk(self.L.cattl3_ctx_set_conv_path(self.h, int(path)))

    @property
    def launches(self):
        return int(self.L.cattl3_ctx_launch_count(self.h))

    @property
    def last_path(self):
        return self.L.cattl3_ctx_last_path(self.h).decode()

    def _call(self, name, dtype, *args):
        suf, _ = _suffix(dtype)
        self._chk(getattr(self.L, "%s_%s" % (name, suf))(self.h, *args))

    # ---- kernel layers (device pointers; tensors are torch CUDA tensors in the reference layout) ----
    def conv_forward(self, g, x, w, b, y, transposed=False):
        self._call("cattl3_transconv_forward" if transposed else "cattl3_conv_forward", x.dtype,
                   ctypes.byref(g), _p(x), _p(w), _p(b), _p(y))

    def conv_backward(self, g, x, w, dy, dw, db, dx, transposed=False):
        self._call("cattl3_transconv_backward" if transposed else "cattl3_conv_backward", x.dtype,
                   ctypes.byref(g), _p(x), _p(w), _p(dy), _p(dw), _p(db), _p(dx))

    def dense_forward(self, n, i, o, x, w, b, y):
        self._call("cattl3_dense_forward", x.dtype, n, i, o, _p(x), _p(w), _p(b), _p(y))

    def dense_backward(self, n, i, o, x, w, dy, dw, db, dx):
        self._call("cattl3_dense_backward", x.dtype, n, i, o, _p(x), _p(w), _p(dy), _p(dw), _p(db), _p(dx))

    @staticmethod
    def _epilogue(act_kind, act_param, act_out, col_stats):
        ep = Epilogue()
        ep.act_kind = ACT_NONE if act_kind is None else int(act_kind)
        ep.act_param = float(act_param)
        ep.act_out = act_out.data_ptr() if act_out is not None else None
        ep.col_stats = col_stats.data_ptr() if col_stats is not None else None
        return ep

    def conv_forward_fused(self, g, x, w, b, y, act_kind=None, act_param=0.0, act_out=None, col_stats=None,
                           transposed=False):
        """Forward with a fused epilogue: act_out = f(y) and / or col_stats (2*F float64) for a BatchNormLayer."""
        ep = self._epilogue(act_kind, act_param, act_out, col_stats)
        self._call("cattl3_transconv_forward_fused" if transposed else "cattl3_conv_forward_fused", x.dtype, ctypes.byref(g), _p(x), _p(w), _p(b), _p(y), ctypes.byref(ep))

    def dense_forward_fused(self, n, i, o, x, w, b, y, act_kind=None, act_param=0.0, act_out=None, col_stats=None):
        ep = self._epilogue(act_kind, act_param, act_out, col_stats)
        self._call("cattl3_dense_forward_fused", x.dtype, n, i, o, _p(x), _p(w), _p(b), _p(y), ctypes.byref(ep))

    def batchnorm_forward_stats(self, per_channel, n, h, w, c, running_init, decay, eps, x, col_stats, shift, gamma,
                                beta, running_mean, running_inv_sd, saved_mean, saved_inv_sd, y, act_kind=None,
                                act_param=0.0, act_out=None, global_count=None):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_batchnorm_forward_stats", x.dtype, int(per_channel), n, h, w, c, int(running_init),
                   ct(decay), ct(eps), _p(x), _p(col_stats), _p(global_count), _p(shift), _p(gamma), _p(beta), _p(running_mean),
                   _p(running_inv_sd), _p(saved_mean), _p(saved_inv_sd), _p(y),
                   ACT_NONE if act_kind is None else int(act_kind), ct(act_param), _p(act_out))

    def constrain(self, x, clip=0.0, max_l1_norm=0.0, max_l2_norm=0.0):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_constrain", x.dtype, ctypes.c_int64(x.numel()), ct(clip), ct(max_l1_norm), ct(max_l2_norm), _p(x))

    def fill(self, count, value, y):
        _, ct = _suffix(y.dtype)
        self._call("cattl3_fill", y.dtype, ctypes.c_int64(count), ct(value), _p(y))

    def throttle(self, max_in_flight=2):
        self._chk(self.L.cattl3_ctx_throttle(self.h, int(max_in_flight)))

    def feed_create(self, slots=3):
        f = ctypes.c_void_p()
        self._chk(self.L.cattl3_feed_create(ctypes.byref(f), self.h, int(slots)))
        return f

    def feed_push(self, feed, host_array):
        """Stages a host numpy array through the feed; returns the device address of its slot."""
        dev = ctypes.c_void_p()
        self._chk(self.L.cattl3_feed_push(feed, _p(host_array), ctypes.c_size_t(host_array.nbytes), ctypes.byref(dev)))
        return dev.value

    def feed_destroy(self, feed):
        self._chk(self.L.cattl3_feed_destroy(feed))

    def slice_rows(self, total, vol, first, rows, src, dst):
        self._call("cattl3_slice_rows", src.dtype, ctypes.c_int64(total), ctypes.c_int64(vol), ctypes.c_int64(first),
                   ctypes.c_int64(rows), _p(src), _p(dst))

    def dropout_forward(self, count, prob, eps, seed, x, y, mask):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_dropout_forward", x.dtype, ctypes.c_int64(count), ct(prob), ct(eps), ctypes.c_uint64(seed),
                   _p(x), _p(y), _p(mask))

    def dropout_backward(self, count, prob, eps, dy, mask, dx):
        _, ct = _suffix(dy.dtype)
        self._call("cattl3_dropout_backward", dy.dtype, ctypes.c_int64(count), ct(prob), ct(eps), _p(dy), _p(mask),
                   _p(dx))

    def loss(self, kind, rows, vol, eps, grad_div, out, obj, loss, grad):
        _, ct = _suffix(out.dtype)
        self._call("cattl3_loss", out.dtype, int(kind), ctypes.c_int64(rows), ctypes.c_int64(vol), ct(eps),
                   ct(grad_div), _p(out), _p(obj), _p(loss), _p(grad))

    def conv_forward_host(self, g, x_host, w, b, y_host, x_dev_keep=None):
        self._chk(self.L.cattl3_conv_forward_host_f32(self.h, ctypes.byref(g), _p(x_host), _p(w), _p(b),
                                                      _p(y_host), _p(x_dev_keep)))

    def conv_backward_host(self, g, x_dev, w, dy_host, dw, db, dx_host):
        self._chk(self.L.cattl3_conv_backward_host_f32(self.h, ctypes.byref(g), _p(x_dev), _p(w), _p(dy_host),
                                                       _p(dw), _p(db), _p(dx_host)))

    def conv_forward_host_async(self, g, x_host, w, b, y_host, x_dev_keep=None):
        self._chk(self.L.cattl3_conv_forward_host_async_f32(self.h, ctypes.byref(g), _p(x_host), _p(w), _p(b),
                                                            _p(y_host), _p(x_dev_keep)))

    def conv_backward_host_async(self, g, x_dev, w, dy_host, dw, db, dx_host):
        self._chk(self.L.cattl3_conv_backward_host_async_f32(self.h, ctypes.byref(g), _p(x_dev), _p(w), _p(dy_host),
                                                             _p(dw), _p(db), _p(dx_host)))

    def host_wait(self):
        self._chk(self.L.cattl3_host_wait(self.h))

    def activation_forward(self, kind, alpha, rows, vol, x, y):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_activation_forward", x.dtype, int(kind), ct(alpha), ctypes.c_int64(rows),
                   ctypes.c_int64(vol), _p(x), _p(y))

    def activation_backward(self, kind, alpha, rows, vol, x, y, dy, dx):
        _, ct = _suffix(dy.dtype)
        self._call("cattl3_activation_backward", dy.dtype, int(kind), ct(alpha), ctypes.c_int64(rows),
                   ctypes.c_int64(vol), _p(x), _p(y), _p(dy), _p(dx))

    def pool_forward(self, kind, g, x, y, argmax):
        self._call("cattl3_pool_forward", x.dtype, int(kind), ctypes.byref(g), _p(x), _p(y), _p(argmax))

    def pool_backward(self, kind, g, dy, argmax, dx):
        self._call("cattl3_pool_backward", dy.dtype, int(kind), ctypes.byref(g), _p(dy), _p(argmax), _p(dx))

    def batchnorm_forward(self, per_channel, n, h, w, c, training, running_init, decay, eps, x, gamma, beta,
                          running_mean, running_inv_sd, saved_mean, saved_inv_sd, y):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_batchnorm_forward", x.dtype, int(per_channel), n, h, w, c, int(training),
                   int(running_init), ct(decay), ct(eps), _p(x), _p(gamma), _p(beta), _p(running_mean),
                   _p(running_inv_sd), _p(saved_mean), _p(saved_inv_sd), _p(y))

    def batchnorm_backward(self, per_channel, n, h, w, c, x, gamma, saved_mean, saved_inv_sd, dy, dgamma, dbeta,
                           dx):
        self._call("cattl3_batchnorm_backward", x.dtype, int(per_channel), n, h, w, c, _p(x), _p(gamma),
                   _p(saved_mean), _p(saved_inv_sd), _p(dy), _p(dgamma), _p(dbeta), _p(dx))

    def batchnorm_stats(self, per_channel, n, h, w, c, x, shift, col_stats):
        self._call("cattl3_batchnorm_stats", x.dtype, int(per_channel), n, h, w, c, _p(x), _p(shift), _p(col_stats))

    def batchnorm_backward_sums(self, per_channel, n, h, w, c, x, saved_mean, saved_inv_sd, dy, dgamma, dbeta, sums):
        self._call("cattl3_batchnorm_backward_sums", x.dtype, int(per_channel), n, h, w, c, _p(x), _p(saved_mean),
                   _p(saved_inv_sd), _p(dy), _p(dgamma), _p(dbeta), _p(sums))

    def batchnorm_backward_apply(self, per_channel, n, h, w, c, global_count, x, gamma, saved_mean, saved_inv_sd, dy,
                                 sums, dx):
        self._call("cattl3_batchnorm_backward_apply", x.dtype, int(per_channel), n, h, w, c,
                   _p(global_count), _p(x), _p(gamma), _p(saved_mean), _p(saved_inv_sd), _p(dy), _p(sums),
                   _p(dx))

    def optimizer_step(self, step, count, p, g, s1=None, s2=None, s3=None):
        self._call("cattl3_optimizer_step", p.dtype, ctypes.byref(step), ctypes.c_int64(count), _p(p), _p(g),
                   _p(s1), _p(s2), _p(s3))

    def add_inplace(self, count, y, x):
        self._call("cattl3_add_inplace", y.dtype, ctypes.c_int64(count), _p(y), _p(x))

    def regularize(self, count, l1, l2, values, grad, penalty):
        """grad += sign(values) * l1 + l2 * values; penalty (float64 device scalar) += l1 |values|_1 + l2 / 2 |values|^2."""
        _, ct = _suffix(values.dtype)
        self._call("cattl3_regularize", values.dtype, ctypes.c_int64(count), ct(l1), ct(l2), _p(values), _p(grad),
                   _p(penalty))

    def muladd(self, count, accumulate, a, b, c, d, out):
        """out = (accumulate ? out : 0) + a * b + (c ? c * d : 0)."""
        self._call("cattl3_muladd", out.dtype, ctypes.c_int64(count), int(bool(accumulate)), _p(a), _p(b), _p(c), _p(d),
                   _p(out))

    def optimizer_step_indirect(self, kind, dev_step, count, p, g, s1=None, s2=None, s3=None):
        """The fused update with the step scalars read from device memory (``dev_step``: a device copy of cattl3_opt_step)."""
        self._call("cattl3_optimizer_step_indirect", p.dtype, int(kind), _p(dev_step), ctypes.c_int64(count), _p(p), _p(g),
                   _p(s1), _p(s2), _p(s3))

    # ---- step graphs and the stream-ordered allocator behind them -------------------------------------------------
    def malloc(self, nbytes):
        ptr = ctypes.c_void_p()
        self._chk(self.L.cattl3_malloc(self.h, ctypes.byref(ptr), ctypes.c_size_t(nbytes)))
        return ptr.value

    def free(self, ptr):
        self._chk(self.L.cattl3_free(self.h, ctypes.c_void_p(ptr)))

    def allocated_bytes(self):
        fn = self.L.cattl3_ctx_allocated_bytes
        fn.restype = ctypes.c_int64
        return int(fn(self.h))

    def graph_begin(self, arena_bytes):
        self._chk(self.L.cattl3_graph_begin(self.h, ctypes.c_size_t(arena_bytes)))

    def graph_end(self):
        g = ctypes.c_void_p()
        self._chk(self.L.cattl3_graph_end(self.h, ctypes.byref(g)))
        return g

    def graph_launch(self, graph):
        self._chk(self.L.cattl3_graph_launch(self.h, graph))

    def graph_destroy(self, graph):
        self._chk(self.L.cattl3_graph_destroy(graph))

    def scale(self, count, alpha, x, y):
        _, ct = _suffix(x.dtype)
        self._call("cattl3_scale", x.dtype, ctypes.c_int64(count), ct(alpha), _p(x), _p(y))


class Comm:
    """One cattl3_comm: this process's seat in the data-parallel exchange (one process per GPU; WORLD_SIZE / RANK from
    the environment, NCCL underneath).  All-reduces are in place, sums, on the context's stream or -- the _async form --
    on the communicator's side stream, joined by wait()."""

    def __init__(self, ctx):
        self.ctx, self.L = ctx, ctx.L
        self.h = ctypes.c_void_p()
        ctx._chk(self.L.cattl3_comm_create_from_env(ctypes.byref(self.h), ctx.h))

    @property
    def world_size(self):
        return int(self.L.cattl3_comm_world_size(self.h))

    @property
    def rank(self):
        return int(self.L.cattl3_comm_rank(self.h))

    def allreduce(self, t, asynchronous=False):
        suf, _ = _suffix(t.dtype)
        fn = getattr(self.L, "cattl3_comm_allreduce_sum_%s%s" % ("async_" if asynchronous else "", suf))
        self.ctx._chk(fn(self.h, _p(t), ctypes.c_int64(t.numel())))

    def wait(self):
        self.ctx._chk(self.L.cattl3_comm_wait(self.h))

    def destroy(self):
        if self.h:
            self.L.cattl3_comm_destroy(self.h)
            self.h = None


def make_opt_step(kind, hyper, timestep, epoch, l2_lambda=0.0, reset_grad=True, dtype="float32"):
    """Evaluates the step-dependent scalars the way the reference's optimizers do (in double, then
    rounded to the Scalar type): AdamOptimizer.hpp:67-68, NadamOptimizer.hpp:45-47,
    MomentumSGDOptimizer.hpp:70-72."""
    import numpy as np
    S = np.float32 if str(dtype).endswith("32") else np.float64
    lr, a, b, eps = (S(v) for v in hyper)
    st = OptStep()
    st.kind, st.reset_grad = int(kind), int(bool(reset_grad))
    st.lr, st.a, st.b, st.eps = float(lr), float(a), float(b), float(eps)
    st.lr_epoch = float(S(lr / S(S(1) + a * S(epoch))))
    one = S(1)
    st.c1 = float(S(one / (1.0 - float(one - a) ** (timestep + 1) + float(eps))))
    st.c1n = float(S(one / (1.0 - float(one - a) ** (timestep + 2) + float(eps))))
    st.c2 = float(S(one / (1.0 - float(one - b) ** (timestep + 1) + float(eps))))
    st.l2_lambda = float(S(l2_lambda))
    return st


def shard_rows(n, rank, world):
    """Rows [lo, hi) of an n-row mini-batch that rank `rank` of `world` keeps -- the same split as
    cattle::SGDOptimizer::shard (cattle/optimizer/SGDOptimizer.hpp): contiguous, covers every row once,
    sizes differ by at most one.  A batch with fewer rows than ranks (a ragged last one) is not split at all: every rank
    takes all of it and the loop divides the loss gradient by the world size as well (replicas(n, world)), so that no rank
    sits out the collectives of the step."""
    if n < world:
        return 0, n
    return n * rank // world, n * (rank + 1) // world


def replicas(n, world):
    """How many ranks process the SAME rows of an n-row mini-batch (1, or `world` for a batch smaller than the world)."""
    return world if world > 1 and n < world else 1
