#!/usr/bin/env python
"""Network-level training throughput through the header-only C++ host side (c-attl3_b200/cattle) -- the layer loop,
the fused epilogues, the device loss, the fused optimizer step and, for N > 1, the NCCL gradient all-reduce and the
synchronised BatchNorm statistics -- on BASELINE.json's network configs:

  --config 1 : examples/cifar_convnet.cpp ConvNet (Dropout removed), 32x32x3, batch 64, CrossEntropy + Nadam
  --config 3 : examples/mnist_autoencoder.cpp auto-encoder (Stacked: Conv/Softplus/Dense/Reshape/TransConv), 28x28x1,
               batch 512, SquaredLoss + Nadam
  --config 4 : ResNet-style ResidualNeuralNetwork (stem conv 7x7/2 + BN + ReLU + MaxPool, `--blocks` modules of
               Conv3x3-BN-ReLU-Conv3x3-BN at `--width` channels, MeanPool + Dense + Softmax head), 224x224x3,
               batch 64 per GPU, data parallel (weak scaling)
  --config 5 : sequences of `--seq` 32x32x3 frames through SequentialNeuralNetwork{ParallelNeuralNetwork of conv lanes,
               DenseNeuralNetwork, MaxPool} and a convolutional LSTMNeuralNetwork head (`--width` channels, default 16),
               batch 64 per GPU, sequential SquaredLoss + Nadam

The driver is oracle/ref_shim.cpp compiled UNCHANGED against the B200 headers (tests/cpp/_build/libcattle_b200_shim.so);
`--impl reference` runs the same driver compiled against the unmodified reference on the host cores (bounded sample).
Launch N ranks with torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_PORT): one process per GPU.  One JSON line
(rank 0).  This is a secondary measurement; bench.py carries the headline metric.

  python scripts/bench_networks.py --config 4
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 \
      scripts/bench_networks.py --config 4
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def run_network(config, impl="b200", dtype="f32", batch=0, steps=8, epochs=5, width=64, blocks=4, image=224, seq=8):
    """One network config through the C++ batch loop on this rank (every rank of a torchrun job calls it); returns the
    result line (meaningful on rank 0)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dt = np.float32 if dtype == "f32" else np.float64
    from oracle import binding
    saved_world = os.environ.get("WORLD_SIZE")
    if impl == "reference":
        if rank != 0:
            return None
        lib = binding.Oracle("ref")
        os.environ["WORLD_SIZE"] = "1"
        world_eff = 1
    else:
        shim = os.path.join(ROOT, "tests", "cpp", "_build", "libcattle_b200_shim.so")
        lib = binding.Oracle("ref", path=shim)
        assert lib.lib.ref_is_b200_build() == 1
        world_eff = world
    per_gpu = batch or {1: 64, 3: 512, 4: 64, 5: 64}[config]
    batch = per_gpu * world_eff                     # the nominal (global) batch the optimizer is built with
    total = batch * steps
    rng = np.random.default_rng({1: 1001, 3: 3001, 4: 4001, 5: 5001}[config])   # identical data on every rank
    if config == 1:
        x = np.asfortranarray(rng.uniform(-1, 1, (total, 32, 32, 3)).astype(dt))
        obj = np.zeros((total, 1, 1, 10), dtype=dt, order="F")
        obj[np.arange(total), 0, 0, np.arange(total) % 10] = 1
        run = lambda n_epochs: lib.train_cifar(x, obj, batch, n_epochs)
        name = "cifar ConvNet 32x32x3"
    elif config == 3:
        x = np.asfortranarray(rng.uniform(0, 1, (total, 28, 28, 1)).astype(dt))
        run = lambda n_epochs: lib.train_autoencoder(x, batch, n_epochs)
        name = "mnist auto-encoder 28x28x1"
    elif config == 5:
        # BASELINE.json configs[4]: sequences of 32x32x3 frames through Sequential{Parallel conv lanes, DenseNet modules,
        # MaxPool} and a convolutional LSTM head (oracle/ref_shim.cpp seqnet_impl); --width lanes / modules / state channels
        width = width if width != 64 else 16
        x = np.asfortranarray(rng.random((total, seq, 32, 32, 3), dtype=np.float32).astype(dt) * 2 - 1)
        obj = np.asfortranarray(rng.random((total, 1, 16, 16, width), dtype=np.float32).astype(dt) - 0.5)
        run = lambda n_epochs: lib.train_seqnet(x, obj, batch, n_epochs, width, width)
        name = ("sequence net: %d frames 32x32x3, Parallel{conv3x3, conv1x1} -> DenseNet x2 -> MaxPool -> conv LSTM, %d ch"
                % (seq, width))
    else:
        s = image
        x = np.asfortranarray(rng.random((total, s, s, 3), dtype=np.float32).astype(dt) * 2 - 1)
        obj = np.zeros((total, 1, 1, 10), dtype=dt, order="F")
        obj[np.arange(total), 0, 0, np.arange(total) % 10] = 1
        arch = (7, 2, 1, width, blocks, 4)   # stem 7x7 stride 2 + max-pool 2x2, head mean-pool 4x4
        run = lambda n_epochs: lib.train_resnet(x, obj, batch, n_epochs, arch)
        name = "ResNet-style residual net %dx%dx3, stem 7x7/2 + pool, %d modules x (conv3x3-BN-ReLU-conv3x3-BN) @ %d ch" % (
            s, s, blocks, width)
    t0 = time.perf_counter()
    _, loss0, ms_warm = run(1)                      # warm-up call: allocations, NCCL communicator, clocks
    # every timed call = one untimed epoch (data set placement, this call's allocations) + one timed epoch; the value
    # is the MEDIAN epoch: single epochs on a shared box occasionally stall for hundreds of ms (min / max reported)
    os.environ["REF_SHIM_WARMUP_EPOCHS"] = "1"
    times = []
    for _ in range(max(3, epochs)):
        _, loss, ms_epoch = run(1)
        times.append(ms_epoch)
    os.environ.pop("REF_SHIM_WARMUP_EPOCHS", None)
    if saved_world is not None:
        os.environ["WORLD_SIZE"] = saved_world
    wall = time.perf_counter() - t0
    ms = sorted(times)[len(times) // 2]
    value = total / (ms / 1000.0)
    return {"metric": "train samples/s (fwd+bwd+step), network level", "impl": impl, "value": round(value, 1),
            "unit": "samples/s", "n_gpus": world_eff, "ms_per_step": round(ms / steps, 3),
            "epoch_ms": {"median": round(ms, 2), "min": round(min(times), 2), "max": round(max(times), 2), "epochs": len(times)},
            "scaling": "weak", "dtype": dtype, "data": "synthetic",
            "config": {"workload": name, "config": config, "batch_per_gpu": per_gpu, "global_batch": batch,
                       "steps_per_epoch": steps, "epochs_timed": epochs,
                       "includes": "data-set slicing + layer loop + loss + gradient exchange (NCCL, cattl3_comm_*) + "
                                   "optimizer step (cattle::NadamOptimizer::train)"},
            "epoch_loss": round(float(loss), 6), "warmup_epoch_ms": round(ms_warm, 1), "wall_s": round(wall, 1),
            "threads": lib.num_threads() if impl == "reference" else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=4, choices=[1, 3, 4, 5])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--steps", type=int, default=8, help="mini-batches per epoch")
    ap.add_argument("--epochs", type=int, default=5, help="timed epochs (median reported; each after an untimed one)")
    ap.add_argument("--width", type=int, default=64)
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--image", type=int, default=224)
    ap.add_argument("--seq", type=int, default=8, help="config 5: frames per sequence")
    args = ap.parse_args()
    line = run_network(args.config, args.impl, args.dtype, args.batch, args.steps, args.epochs, args.width, args.blocks,
                       args.image, args.seq)
    if line is not None and int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
