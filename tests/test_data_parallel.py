"""The N > 1 path.

CPU (gloo, world size 2, runs anywhere): the data-parallel contract of the batch loop -- every rank keeps
rows [r*B/G, (r+1)*B/G) of the mini-batch, back-propagates the loss gradient divided by the GLOBAL batch
size, and the sum-all-reduce of the shard gradients equals the single-process gradient (SURVEY.md
section 8e) -- checked with the CPU oracle standing in for the device kernels.

GPU (-m gpu, needs 2 devices; skipped on a 1-GPU box): two processes drive the C++ batch loop
(cattle::SGDOptimizer::_train + b200::Communicator, NCCL) on config 1 and must end with the parameters
of the single-process run.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases as C

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases as C
from oracle import binding
from __graft_entry__ import load_package
pkg = load_package()
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
orc = binding.Oracle("orc")
case = (6, 7, 5, 3, 4, 3, 2, 1, 0, 1, 2, 0, 1)          # n = 6 rows over 2 ranks, ragged geometry
g, x, w, b, dy = C.conv_inputs(case, np.float64, 77)
dy = dy / g.n                                            # loss gradient / GLOBAL batch size
full = orc.conv(g, x, w, b, dy)
lo, hi = pkg.shard_rows(g.n, rank, world)
gs = binding.Geom(hi - lo, *case[1:])
part = orc.conv(gs, np.asfortranarray(x[lo:hi]), w, b, np.asfortranarray(dy[lo:hi]))
arena = torch.from_numpy(np.concatenate([part["dw"].ravel(order="F"), part["db"].ravel(order="F")]))
dist.all_reduce(arena)                                   # what ncclAllReduce does on the device arena
want = np.concatenate([full["dw"].ravel(order="F"), full["db"].ravel(order="F")])
err = float(np.max(np.abs(arena.numpy() - want)) / np.max(np.abs(want)))
ok_y = np.allclose(part["y"], full["y"][lo:hi], rtol=0, atol=1e-13)
ok_dx = np.allclose(part["dx"], full["dx"][lo:hi], rtol=0, atol=1e-13)
sizes = [pkg.shard_rows(7, r, 3) for r in range(3)]
assert sizes == [(0, 2), (2, 4), (4, 7)], sizes
# a ragged last batch with fewer rows than ranks: every rank takes all of it, the loss gradient is divided by the world size as
# well, and the all-reduced gradient is still the single-process one (no rank skips the step and its collectives)
assert [pkg.shard_rows(1, r, 2) for r in range(2)] == [(0, 1), (0, 1)] and pkg.replicas(1, 2) == 2 and pkg.replicas(6, 2) == 1
g1 = binding.Geom(1, *case[1:])
one = orc.conv(g1, np.asfortranarray(x[:1]), w, b, np.asfortranarray(dy[:1]))
lo1, hi1 = pkg.shard_rows(1, rank, world)
rep = orc.conv(g1, np.asfortranarray(x[lo1:hi1]), w, b, np.asfortranarray(dy[lo1:hi1] / pkg.replicas(1, world)))
small = torch.from_numpy(np.concatenate([rep["dw"].ravel(order="F"), rep["db"].ravel(order="F")]))
dist.all_reduce(small)
want1 = np.concatenate([one["dw"].ravel(order="F"), one["db"].ravel(order="F")])
assert float(np.max(np.abs(small.numpy() - want1)) / np.max(np.abs(want1))) < 1e-13
print("RESULT", rank, err, ok_y, ok_dx)
assert err < 1e-13 and ok_y and ok_dx
dist.destroy_process_group()
"""


def test_sharded_gradients_sum_to_full_batch_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
        assert "RESULT" in o


_DP_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases as C
from oracle import binding
lib = binding.Oracle("ref", path={shim!r})
dt = np.float32 if sys.argv[1] == "f32" else np.float64
rng = np.random.default_rng(1003)
n = 96
x = C.rand(rng, (n, 32, 32, 3), dt)
obj = np.zeros((n, 1, 1, 10), dtype=dt, order="F")
obj[np.arange(n), 0, 0, np.arange(n) % 10] = 1
p0 = np.load({p0!r})
p, loss, ms = lib.train_cifar(x, obj, 32, 2, params_in=p0.astype(dt))
np.save(sys.argv[2], p)
print("LOSS %.9f" % loss)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_two_process_training_equals_one_process(tmp_path, golden, suf):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    shim = os.path.join(ROOT, "tests", "cpp", "_build", "libcattle_b200_shim.so")
    p0 = tmp_path / "p0.npy"
    np.save(p0, golden["cifar/%s/p0" % suf])
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER.format(root=ROOT, shim=shim, p0=str(p0)))

    def run(world):
        procs = []
        for r in range(world):
            env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), LOCAL_RANK=str(r), MASTER_PORT="29542",
                       CATTL3_COMM_ID_FILE=str(tmp_path / ("id_%d" % world)))
            procs.append(subprocess.Popen([sys.executable, str(script), suf, str(tmp_path / ("p_%d_%d.npy" % (world, r)))],
                                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
        losses = [float(o.split("LOSS")[1].split()[0]) for o in outs]
        return [np.load(tmp_path / ("p_%d_%d.npy" % (world, r))) for r in range(world)], losses

    one, loss1 = run(1)
    two, loss2 = run(2)
    tol = 2e-5 if suf == "f32" else 1e-12
    assert np.array_equal(two[0], two[1]), "ranks diverged"
    err = C.relerr(two[0], one[0])
    print("1 vs 2 processes (%s): param err %.2e, loss %.7f vs %.7f" % (suf, err, loss1[0], loss2[0]))
    assert err < tol, err
    assert abs(loss1[0] - loss2[0]) < 1e-5 * max(1.0, abs(loss1[0]))


_DP_RESNET_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases as C
from oracle import binding
lib = binding.Oracle("ref", path={shim!r})
dt = np.float32 if sys.argv[1] == "f32" else np.float64
x, obj = C.resnet_inputs(dt, total=int(sys.argv[3]), seed=4003)
n = lib.train_resnet(x, obj, 64, -1, C.RESNET_SMALL)
p, loss, ms = lib.train_resnet(x, obj, 64, 2, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))
np.save(sys.argv[2], p)
print("LOSS %.9f" % loss)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("total", [128, 129])
@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_two_process_resnet_with_synchronised_batchnorm_equals_one_process(tmp_path, suf, total):
    """Config 4 at test size on 2 GPUs: with the BatchNorm statistics all-reduced (forward sums + count, backward
    sums) two processes on half-batches must reproduce the single-process run on the full batch -- parameters
    INCLUDING the running mean / inverse standard deviation -- and both ranks must hold identical parameters.
    total = 129 leaves a ragged last batch of ONE row for two ranks: no rank may sit the step out (it would skip the
    BatchNorm collectives and hang the other), so both take the row and the loop rescales (SGDOptimizer::_train)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    shim = os.path.join(ROOT, "tests", "cpp", "_build", "libcattle_b200_shim.so")
    script = tmp_path / "dp_resnet_worker.py"
    script.write_text(_DP_RESNET_WORKER.format(root=ROOT, shim=shim))

    def run(world):
        procs = []
        for r in range(world):
            env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), LOCAL_RANK=str(r), MASTER_PORT="29543",
                       CATTL3_COMM_ID_FILE=str(tmp_path / ("rid_%d" % world)))
            procs.append(subprocess.Popen([sys.executable, str(script), suf, str(tmp_path / ("rp_%d_%d.npy" % (world, r))),
                                           str(total)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
        losses = [float(o.split("LOSS")[1].split()[0]) for o in outs]
        return [np.load(tmp_path / ("rp_%d_%d.npy" % (world, r))) for r in range(world)], losses

    one, loss1 = run(1)
    two, loss2 = run(2)
    assert np.array_equal(two[0], two[1]), "ranks diverged"
    err = C.relerr(two[0], one[0])
    print("ResNet, 1 vs 2 processes (%s): param err %.2e, loss %.7f vs %.7f" % (suf, err, loss1[0], loss2[0]))
    assert err < (1e-4 if suf == "f32" else 1e-10), err
    assert abs(loss1[0] - loss2[0]) < 1e-5 * max(1.0, abs(loss1[0]))


_DP_SEQNET_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases as C
from oracle import binding
lib = binding.Oracle("ref", path={shim!r})
dt = np.float32 if sys.argv[1] == "f32" else np.float64
x, obj = C.seqnet_inputs(dt, total=48, seq=4, seed=5006)
n = lib.train_seqnet(x, obj, 16, -1, **C.SEQNET_SMALL)
p, loss, ms = lib.train_seqnet(x, obj, 16, 2, params_in=C.seeded_params(n, dt, 5002), **C.SEQNET_SMALL)
np.save(sys.argv[2], p)
print("LOSS %.9f" % loss)
"""


@pytest.mark.gpu
@pytest.mark.parametrize("suf", ["f32", "f64"])
def test_two_process_sequence_network_equals_one_process(tmp_path, suf):
    """Config 5 at test size on 2 GPUs: the sequence batch is sharded by samples (rows of the (samples * steps) x volume
    layout stay together per sample), the LSTM's shared-parameter gradients are all-reduced with the rest of the
    arena; two processes on half-batches must reproduce the single-process run and hold identical parameters."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    shim = os.path.join(ROOT, "tests", "cpp", "_build", "libcattle_b200_shim.so")
    script = tmp_path / "dp_seqnet_worker.py"
    script.write_text(_DP_SEQNET_WORKER.format(root=ROOT, shim=shim))

    def run(world):
        procs = []
        for r in range(world):
            env = dict(os.environ, WORLD_SIZE=str(world), RANK=str(r), LOCAL_RANK=str(r), MASTER_PORT="29547",
                       CATTL3_COMM_ID_FILE=str(tmp_path / ("sid_%d" % world)))
            procs.append(subprocess.Popen([sys.executable, str(script), suf, str(tmp_path / ("sp_%d_%d.npy" % (world, r)))],
                                          env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
        losses = [float(o.split("LOSS")[1].split()[0]) for o in outs]
        return [np.load(tmp_path / ("sp_%d_%d.npy" % (world, r))) for r in range(world)], losses

    one, loss1 = run(1)
    two, loss2 = run(2)
    assert np.array_equal(two[0], two[1]), "ranks diverged"
    err = C.relerr(two[0], one[0])
    print("sequence network, 1 vs 2 processes (%s): param err %.2e, loss %.7f vs %.7f" % (suf, err, loss1[0], loss2[0]))
    assert err < (1e-4 if suf == "f32" else 1e-10), err
    assert abs(loss1[0] - loss2[0]) < 1e-5 * max(1.0, abs(loss1[0]))


_GLOO_BN_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import cases as C
from oracle import binding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
orc = binding.Oracle("orc")
rng = np.random.default_rng(91)
n, h, w, c = 10, 4, 3, 5
x = C.rand(rng, (n, h, w, c), np.float64, -1, 3)
gm, bt, dy = C.rand(rng, (c,), np.float64, 0.5, 1.5), C.rand(rng, (c,), np.float64), C.rand(rng, (n, h, w, c), np.float64)
full = orc.batchnorm(1, [x], gm, bt, dy)
lo, hi = (0, 4) if rank == 0 else (4, 10)                 # uneven shards
xs, dys = x[lo:hi], dy[lo:hi]
shift = np.linspace(-1, 1, c)                              # any rank-invariant shift (the layer uses its running mean)
# forward: [sum (x - shift) | sum (x - shift)^2 | count] all-reduced in one message (cattle/layer/BatchNormLayer.hpp)
d = xs.reshape(-1, c, order="F") - shift
msg = torch.from_numpy(np.concatenate([d.sum(0), (d * d).sum(0), [d.shape[0]]]))
dist.all_reduce(msg)
s1, s2, L = msg[:c].numpy(), msg[c:2 * c].numpy(), float(msg[2 * c])
mean = shift + s1 / L
inv_sd = 1.0 / np.sqrt(s2 / L - (s1 / L) ** 2 + 1e-5)
xhat = (xs - mean) * inv_sd
y = xhat * gm + bt
# backward: local sums -> dgamma / dbeta shares; all-reduced sums -> dx with the GLOBAL count
g = (dys * gm)
sums = torch.from_numpy(np.concatenate([dys.reshape(-1, c, order="F").sum(0), (dys * xhat).reshape(-1, c, order="F").sum(0)]))
grads = sums.clone()
dist.all_reduce(sums)
dist.all_reduce(grads)                                      # what the optimizer's gradient all-reduce does
sg, sxg = gm * sums[:c].numpy(), gm * sums[c:].numpy()
dx = (L * g - sg - xhat * sxg) * inv_sd / L
err = max(C.relerr(y, full["y"][lo:hi]), C.relerr(dx, full["dx"][lo:hi]), C.relerr(mean, full["run_mean"]),
          C.relerr(inv_sd, full["run_inv_sd"]), C.relerr(grads[:c].numpy(), full["dbeta"]),
          C.relerr(grads[c:].numpy(), full["dgamma"]))
print("RESULT", rank, err)
assert err < 1e-11, err
dist.destroy_process_group()
"""


def test_synchronised_batchnorm_contract_gloo(tmp_path):
    """CPU, world size 2, uneven shards: the message layout and arithmetic of synchronised BatchNorm (shifted sums
    + element count forward, two sums backward, local dgamma / dbeta shares) reproduce the single-process layer
    (the oracle on the full batch)."""
    script = tmp_path / "bn_worker.py"
    script.write_text(_GLOO_BN_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29544", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
        assert "RESULT" in o
