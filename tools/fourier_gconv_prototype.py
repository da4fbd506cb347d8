"""CPU prototype (numpy, FP64) of the group-Fourier form of YOHO's 13-tap group convolution — the planned way past the
1/3 tensor-core ceiling of the 3-product gather-GEMM.  Not used by the product.

  y(o,g) = sum_c sum_k W[o,c,k] x(c, h_k g)            (utils/network.py:46-52 gather + Conv2d(1,13); h_k g = N[g][k])
  x^_rho(c) = sum_g x(c,g) rho(g)                      (d x d matrix per real irrep rho of the icosahedral group, d in 1,3,3,4,5)
  y^_rho(o) = sum_c [ sum_k W[o,c,k] rho(h_k)^T ] x^_rho(c)
  y(o,g)   = (1/60) sum_rho d_rho tr( y^_rho(o) rho(g)^T )

MACs per (in,out) channel pair: 60*13 = 780 in the group domain, sum d^3 = 244 in the Fourier domain, plus two 60x60
transforms per channel at every layer boundary (BN+ReLU live in the group domain).
The irreps are obtained numerically from the reference's multiplication table: a random element of the commutant of the
regular representation has one d-fold eigenvalue per copy of a d-dimensional irrep.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import yoho_oracle as O

R, P, N = O.load_tables()
G = 60
# group law on indices: P[a][b] = idx(R_b R_a)  ->  mul(b, a) = idx(R_b R_a)
mul = lambda b, a: int(P[a][b])
inv = np.array([[b for b in range(G) if mul(b, a) == 0][0] for a in range(G)])

# left-regular representation L(g) e_a = e_{g a}
L = np.zeros((G, G, G))
for g in range(G):
    for a in range(G):
        L[g, mul(g, a), a] = 1.0
rs = np.random.RandomState(0)
H = rs.standard_normal((G, G)); H = H + H.T
Havg = sum(L[g] @ H @ L[g].T for g in range(G)) / G            # commutes with every L(g)
w, U = np.linalg.eigh(Havg)
groups, start = [], 0
for i in range(1, G + 1):
    if i == G or abs(w[i] - w[start]) > 1e-8:
        groups.append((start, i)); start = i
irreps = {}
g72 = int(N[0][1])                                              # a 72-degree rotation (first non-identity tap)
for (s, e) in groups:
    B = U[:, s:e]
    rho = np.stack([B.T @ L[g] @ B for g in range(G)])          # d x d real orthogonal
    d = e - s
    key = (d, round(float(np.trace(rho[g72])), 3))             # the two 3-dim irreps differ in chi(72 deg) = (1 +- sqrt5)/2
    irreps.setdefault(key, rho)
dims = sorted(k[0] for k in irreps)
assert dims == [1, 3, 3, 4, 5], dims
for rho in irreps.values():                                     # homomorphism + orthogonality
    a, b = 7, 23
    assert np.allclose(rho[mul(a, b)], rho[a] @ rho[b]) and np.allclose(rho[a] @ rho[a].T, np.eye(rho.shape[1]))

# ---- convolution theorem against the oracle's group convolution ------------------------------------------------------
C, Oc, Bn = 6, 5, 3
x = rs.standard_normal((Bn, C, G))
W = rs.standard_normal((Oc, C, 1, 13))
bias = rs.standard_normal(Oc)
sd = {"w.weight": W.astype(np.float64), "w.bias": bias.astype(np.float64)}
want = O.gconv(torch.from_numpy(x), sd, "w", N, torch.float64).numpy()           # [B,O,60]
h = [int(v) for v in N[0]]                                                       # taps: N[g][k] = idx(R_{h_k} R_g) = mul(h_k, g)
assert all(int(N[g][k]) == mul(h[k], g) for g in range(G) for k in range(13))
y = np.zeros((Bn, Oc, G))
macs_f = 0
for (d, _), rho in irreps.items():
    xh = np.einsum("bcg,gij->bcij", x, rho)                                      # forward transform
    wh = np.einsum("ock,kij->ocji", W[:, :, 0, :], rho[h])                        # sum_k W rho(h_k)^T
    yh = np.einsum("ocil,bclj->boij", wh, xh)                                     # d x d matrix products: the GEMM part
    macs_f += d ** 3
    y += (d / G) * np.einsum("boij,gij->bog", yh, rho)                           # inverse transform
y += bias[None, :, None]
err = np.abs(y - want).max()
print(f"irreps {dims}; max |Fourier-domain conv - oracle gconv| = {err:.2e} (FP64)")
print(f"MACs per channel pair: group domain {G * 13}, Fourier domain {macs_f} ({G * 13 / macs_f:.2f}x fewer); "
      f"transforms 2 x {G * G} MACs per channel per layer boundary")
assert err < 1e-10
