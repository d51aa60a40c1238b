"""Helpers for the -m gpu parity tests: numpy (column-major, reference layout) <-> flat CUDA tensors."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
_ctx = {}


def torch_mod():
    import torch
    return torch


def ctx(path=None):
    """A context on cuda:0 bound to torch's current stream."""
    torch = torch_mod()
    key = "c"
    if key not in _ctx:
        _ctx[key] = pkg.Context(0, torch.cuda.current_stream().cuda_stream)
    c = _ctx[key]
    c.set_conv_path(pkg.PATH_AUTO if path is None else path)
    return c


def dev(a):
    torch = torch_mod()
    return torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).cuda()


def zeros(shape, dtype):
    torch = torch_mod()
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    return torch.zeros(int(np.prod(shape)), dtype=tdt, device="cuda")


def host(t, shape):
    return t.cpu().numpy().reshape(shape, order="F")
