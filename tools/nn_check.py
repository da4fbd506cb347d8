"""Mutual 1-NN search in isolation (target for ncu): tensor-core search vs SIMT twin, timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from yoho_b200.engine import get_engine
eng = get_engine()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rs = np.random.RandomState(0)
base = rs.standard_normal((1, 32)).astype(np.float32) * 0.5
dA = (base + rs.standard_normal((K, 32)) * 0.05).astype(np.float32)
dB = (base + rs.standard_normal((K, 32)) * 0.05).astype(np.float32)
dB[: K // 2] = dA[rs.permutation(K)[: K // 2]] + (rs.standard_normal((K // 2, 32)) * 0.005).astype(np.float32)
a, b = torch.from_numpy(dA).cuda(), torch.from_numpy(dB).cuda()
for flags, name in ((eng.DEFAULT_TUNING, "tensor-core"), (eng.DEFAULT_TUNING | 16384, "simt")):
    eng.set_tuning(0, flags)
    p, n = eng.mutual_nn(a, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        p, n = eng.mutual_nn(a, b)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per mutual_nn, M = {int(n.item())}")
eng.set_tuning(0, eng.DEFAULT_TUNING)
