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


def _fconv(X, packed, Oc):
    """Per-irrep gather-GEMMs of one layer on Fourier coefficients X [B,60,C] -> [B,60,Oc] (no bias), as the device runs them."""
    B = X.shape[0]
    Y = np.zeros((B, 60, Oc))
    for p in packed:
        d = p["d"]
        w = p["w"].astype(np.float64)
        for j in range(d):
            acc = sum(X[:, p["idx"][j][l]] @ w[l] for l in range(d))          # [B, d*Oc]
            for i in range(d):
                Y[:, p["omap"][j][i]] = acc[:, i * Oc:(i + 1) * Oc]
    return Y


def test_all_fourier_part1_equals_oracle(tables):
    """The whole PartI stack kept in the Fourier domain between the BatchNorm/ReLU points (csrc/part1.cu, all-Fourier branch):
    layer 1 and layer 4 as per-irrep GEMMs, the identity shortcut added as Fourier coefficients with its bias joining the
    residual block's bias in the group domain — equals the oracle's PartI (utils/network.py:86-105)."""
    from yoho_b200 import synth
    _, _, N = tables
    T = fourier.build()
    F = T["F"]
    sd = {k: np.asarray(v, np.float64) for k, v in synth.synth_state_dict("PartI", 3).items()}
    x, _ = synth.make_fragment(5, 17)
    want = O.part1_forward(x, synth.synth_state_dict("PartI", 3), N, dtype=torch.float64)
    blk = "PartI_net.SO3_Conv_layers.0."

    def bn(prefix):
        s = sd[prefix + ".weight"] / np.sqrt(sd[prefix + ".running_var"] + 1e-5)
        return s, sd[prefix + ".bias"] - sd[prefix + ".running_mean"] * s

    fwd = lambda a: np.einsum("mg,bgc->bmc", F, a)
    inv = lambda A: np.einsum("mg,bmc->bgc", F, A)
    p_in = fourier.pack_layer(sd["PartI_net.Conv_in.0.weight"], T)
    p_a = fourier.pack_layer(sd[blk + "comb_layer_in.2.weight"], T)
    p_b = fourier.pack_layer(sd[blk + "comb_layer_out.2.weight"], T)
    p_out = fourier.pack_layer(sd["PartI_net.Conv_out.comb_layer.2.weight"], T)
    b1, b2 = sd["PartI_net.Conv_in.0.bias"], sd[blk + "comb_layer_in.2.bias"]
    b3, b4 = sd[blk + "comb_layer_out.2.bias"], sd["PartI_net.Conv_out.comb_layer.2.bias"]
    sa, ta = bn(blk + "comb_layer_in.0")
    sb, tb = bn(blk + "comb_layer_out.0")
    so, to = bn("PartI_net.Conv_out.comb_layer.0")
    x0 = np.transpose(x.astype(np.float64), (0, 2, 1))                         # [b][g][c]
    Y1 = _fconv(fwd(x0), p_in, 256)
    X1 = fwd(np.maximum((inv(Y1) + b1) * sa + ta, 0))
    Y2 = _fconv(X1, p_a, 512)
    X2 = fwd(np.maximum((inv(Y2) + b2) * sb + tb, 0))
    Y3 = _fconv(X2, p_b, 256)
    X3 = fwd(np.maximum((inv(Y3 + Y1) + (b3 + b1)) * so + to, 0))
    Y4 = _fconv(X3, p_out, 32)
    e = inv(Y4) + b4 + x0                                                       # [b][g][c]
    eqv = e / np.maximum(np.linalg.norm(e, axis=2, keepdims=True), 1e-4)
    got = np.transpose(eqv, (0, 2, 1))
    assert np.abs(got - want["eqv"].numpy()).max() < 1e-6      # the packed weights are float32
