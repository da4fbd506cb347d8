"""CPU: the group-Fourier tables and weight packing of `yoho_b200.fourier`, emulated in numpy exactly as the device runs
them (orthogonal transform -> per-irrep gather-GEMM with d taps and remapped output rows -> inverse transform), reproduce
the oracle's group convolution."""
import numpy as np
import torch

import yoho_oracle as O
from yoho_b200 import fourier


def test_fourier_dataflow_equals_group_convolution(tables):
    _, _, N = tables
    T = fourier.build()
    F = T["F"]
    rs = np.random.RandomState(1)
    B, C, Oc = 3, 64, 96
    x = rs.standard_normal((B, C, 60))
    W = rs.standard_normal((Oc, C, 1, 13))
    bias = rs.standard_normal(Oc)
    want = O.gconv(torch.from_numpy(x), {"w.weight": W, "w.bias": bias}, "w", N, torch.float64).numpy()   # [B,O,60]
    a = np.transpose(x, (0, 2, 1))                    # device layout [b][g][c]
    X = np.einsum("mg,bgc->bmc", F, a)                # forward transform: rows m, channels contiguous
    Y = np.zeros((B, 60, Oc))
    for p in fourier.pack_layer(W, T):
        d = p["d"]
        for b in range(B):
            for j in range(d):                        # GEMM row (b, j)
                acc = np.zeros(d * Oc)
                for l in range(d):                    # taps
                    acc += X[b, p["idx"][j][l]] @ p["w"][l].astype(np.float64)
                for i in range(d):                    # column group i -> output row omap[j][i]
                    Y[b, p["omap"][j][i]] = acc[i * Oc:(i + 1) * Oc]
    y = np.einsum("mg,bmo->bgo", F, Y) + bias         # inverse transform (F orthogonal) + bias in the group domain
    assert np.abs(np.transpose(y, (0, 2, 1)) - want).max() < 1e-5


def test_tables_are_deterministic_and_orthogonal():
    a, b = fourier.build(), fourier.build.__wrapped__()
    assert np.array_equal(a["F"], b["F"])
    assert np.allclose(a["F"] @ a["F"].T, np.eye(60), atol=1e-10)
    assert [ir["d"] for ir in a["irreps"]] == [1, 3, 3, 4, 5] and [ir["off"] for ir in a["irreps"]] == [0, 1, 10, 19, 35]
