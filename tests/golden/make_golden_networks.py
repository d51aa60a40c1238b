#!/usr/bin/env python
"""Regenerates tests/golden/reference_networks.npz from the UNMODIFIED reference (oracle/_ref/libcattle_ref.so,
compiled from /root/reference by oracle/Makefile): parameters and epoch loss of BASELINE.json configs[2] (the
mnist auto-encoder, StackedNeuralNetwork + SquaredLoss) and configs[3] (a ResNet-style ResidualNeuralNetwork of
conv + BatchNorm + ReLU modules, CrossEntropyLoss) and configs[4] (SequentialNeuralNetwork{Parallel conv lanes, DenseNet
modules, MaxPool} -> convolutional LSTMNeuralNetwork, SquaredLoss) after a few Nadam steps from seeded inputs and seeded starting
parameters (tests/cases.py), float and double.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_networks.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases as C  # noqa: E402
from oracle.binding import Oracle  # noqa: E402


def main():
    ref = Oracle("ref")
    vectors = np.load(os.path.join(HERE, "reference_vectors.npz"))
    out = {}
    for suf, dt in (("f32", np.float32), ("f64", np.float64)):
        x = C.autoencoder_inputs(dt)
        n = ref.train_autoencoder(x, 4, -1)
        p1, loss, _ = ref.train_autoencoder(x, 4, 2, params_in=C.seeded_params(n, dt, 3002))
        out["autoencoder/%s/p1" % suf] = p1
        out["autoencoder/%s/loss" % suf] = np.array([loss])
        x, obj = C.resnet_inputs(dt)
        n = ref.train_resnet(x, obj, 32, -1, C.RESNET_SMALL)
        p1, loss, _ = ref.train_resnet(x, obj, 32, 2, C.RESNET_SMALL, params_in=C.seeded_params(n, dt, 4002))
        out["resnet/%s/p1" % suf] = p1
        out["resnet/%s/loss" % suf] = np.array([loss])
        x, obj = C.seqnet_inputs(dt)
        n5 = ref.train_seqnet(x, obj, 8, -1, **C.SEQNET_SMALL)
        p1, loss5, _ = ref.train_seqnet(x, obj, 8, 2, params_in=C.seeded_params(n5, dt, 5002), **C.SEQNET_SMALL)
        out["seqnet/%s/p1" % suf] = p1
        out["seqnet/%s/loss" % suf] = np.array([loss5])
        print(suf, "seqnet params", n5, "loss", loss5)
        # config 1 with an ElasticNet penalty on every weight matrix (REF_SHIM_REG, tests only): the device regularisation
        x, obj = C.cifar_inputs(dt)
        os.environ["REF_SHIM_REG"] = C.CIFAR_REG
        p1, loss_reg, _ = ref.train_cifar(x, obj, 16, 2, params_in=np.ascontiguousarray(vectors["cifar/%s/p0" % suf]))
        del os.environ["REF_SHIM_REG"]
        out["cifar_reg/%s/p1" % suf] = p1
        out["cifar_reg/%s/loss" % suf] = np.array([loss_reg])
        print(suf, "cifar + ElasticNet loss", loss_reg)
        # config 1 with all six constraints of StandardParameters on every weight matrix (REF_SHIM_CONSTRAINTS, tests only)
        os.environ["REF_SHIM_CONSTRAINTS"] = C.CIFAR_CON
        p1, loss_con, _ = ref.train_cifar(x, obj, 16, 2, params_in=np.ascontiguousarray(vectors["cifar/%s/p0" % suf]))
        del os.environ["REF_SHIM_CONSTRAINTS"]
        out["cifar_con/%s/p1" % suf] = p1
        out["cifar_con/%s/loss" % suf] = np.array([loss_con])
        print(suf, "cifar + constraints loss", loss_con)
        print(suf, "autoencoder params", out["autoencoder/%s/p1" % suf].size, "resnet params", n, "loss", loss)
    path = os.path.join(HERE, "reference_networks.npz")
    np.savez_compressed(path, **out)
    print("wrote %d arrays, %.1f KB" % (len(out), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
