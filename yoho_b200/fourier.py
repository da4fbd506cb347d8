"""Group-Fourier tables for the icosahedral group (host side, numpy) — used by the `tcgen05_fourier` implementation of
the two wide PartI layers.

The 13-tap group convolution  y(o,g) = sum_{c,k} W[o,c,k] x(c, h_k g)  (utils/network.py:46-52 + Conv2d(1,13)) becomes, in
the basis of the real irreducible representations rho (dims 1,3,3,4,5),
    Y^[(rho,i,j)](o) = sum_c sum_l  w^_rho(o,c)[i][l] * X^[(rho,l,j)](c),     w^_rho(o,c) = sum_k W[o,c,k] rho(h_k)^T
with X^ = F x, F[(rho,l,j)][g] = sqrt(d/60) rho(g)[l][j] an ORTHOGONAL 60x60 matrix — 244 instead of 780 multiply-adds per
channel pair.  Each irrep is one gather-GEMM with d taps: rows (b,j), tap l reads coefficient row (rho,l,j), columns (i,o).
The irreps are computed from the reference's multiplication table (`60_60.npy`): a random element of the commutant of the
regular representation has one d-fold eigenvalue per copy of each d-dimensional irrep (deterministic seed).
"""
import functools
import numpy as np

from . import group as _group

G = 60


@functools.lru_cache(maxsize=4)
def build(so3_dir=None):
    t = _group.load(so3_dir)
    P, N = t.P, t.N
    mul = lambda b, a: int(P[a][b])                      # idx(R_b R_a)
    L = np.zeros((G, G, G))
    for g in range(G):
        for a in range(G):
            L[g, mul(g, a), a] = 1.0
    rs = np.random.RandomState(20240925)
    H = rs.standard_normal((G, G))
    H = H + H.T
    Havg = sum(L[g] @ H @ L[g].T for g in range(G)) / G
    w, U = np.linalg.eigh(Havg)
    groups, start = [], 0
    for i in range(1, G + 1):
        if i == G or abs(w[i] - w[start]) > 1e-8:
            groups.append((start, i))
            start = i
    g72 = int(N[0][1])
    found = {}
    for (s, e) in groups:
        B = U[:, s:e]
        rho = np.stack([B.T @ L[g] @ B for g in range(G)])
        key = (e - s, round(float(np.trace(rho[g72])), 3))
        found.setdefault(key, rho)
    keys = sorted(found)
    assert [k[0] for k in keys] == [1, 3, 3, 4, 5], keys
    irreps, off = [], 0
    F = np.zeros((G, G))
    for k in keys:
        rho = found[k]
        d = k[0]
        for l in range(d):
            for j in range(d):
                F[off + l * d + j, :] = np.sqrt(d / G) * rho[:, l, j]
        irreps.append(dict(d=d, off=off, rho=rho))
        off += d * d
    assert np.allclose(F @ F.T, np.eye(G), atol=1e-10)
    taps = [int(v) for v in N[0]]                        # h_k, with N[g][k] = idx(R_{h_k} R_g)
    return dict(F=F, irreps=irreps, taps=taps)


def pack_layer(W, tables):
    """W [O,C,1,13] (reference layout) -> per irrep: weights [d taps][C][d*O] float32 with column n = i*O + o, the input-row
    table idx[j][l] = off + l*d + j and the output-row table omap[j][i] = off + i*d + j."""
    W = np.asarray(W, np.float64)[:, :, 0, :]
    O_, C = W.shape[0], W.shape[1]
    out = []
    for ir in tables["irreps"]:
        d, off, rho = ir["d"], ir["off"], ir["rho"]
        what = np.einsum("ock,kli->ocil", W, rho[tables["taps"]])          # sum_k W rho(h_k)^T  -> [o,c,i,l]
        wt = np.ascontiguousarray(np.transpose(what, (3, 1, 2, 0))).reshape(d, C, d * O_)   # [l][c][(i,o)]
        idx = np.array([[off + l * d + j for l in range(d)] for j in range(d)], np.int32)
        omap = np.array([[off + i * d + j for i in range(d)] for j in range(d)], np.int32)
        out.append(dict(d=d, off=off, w=wt.astype(np.float32), idx=idx, omap=omap))
    return out
