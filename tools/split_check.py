"""PartII with / without split accumulators in its two 13-tap GEMMs (tuning key 1): error against the oracle on the shipped
checkpoint and device time."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import yoho_oracle as O
from yoho_b200 import synth
from yoho_b200.engine import get_engine
eng = get_engine()
R, P, N = O.load_tables()
ck = os.path.join(ROOT, "oracle", "_ref", "ckpt", "PartII.npz")
for name, sd in (("real", dict(np.load(ck)) if os.path.exists(ck) else None), ("synth", synth.synth_state_dict("PartII", 2))):
    if sd is None:
        continue
    eng.load_part2(sd)
    K, M = 5000, 2800
    rs = np.random.RandomState(5)
    fA, kA = synth.make_fragment(K, 41); fB, kB = synth.make_fragment(K, 42)
    yA, _ = synth.make_fragment(K, 43); yB, _ = synth.make_fragment(K, 44)
    pairs = np.stack([np.sort(rs.permutation(K)[:M]), rs.permutation(K)[:M]], 1).astype(np.int64)
    pre = rs.randint(0, 60, M).astype(np.int64)
    rows = np.arange(0, M, 9)
    want = O.part2_forward(fA[pairs[rows, 0]], fB[pairs[rows, 1]], yA[pairs[rows, 0]], yB[pairs[rows, 1]], pre[rows], sd, P, N).numpy()
    w64 = O.part2_forward(fA[pairs[rows, 0]], fB[pairs[rows, 1]], yA[pairs[rows, 0]], yB[pairs[rows, 1]], pre[rows], sd, P, N, torch.float64).numpy()
    dev = eng.device
    args = [torch.from_numpy(v).to(dev) for v in (fA, fB, yA, yB)]
    pd, prd = torch.from_numpy(pre).to(dev), torch.from_numpy(pairs).to(dev)
    for mk in (1024, 2048, 4096):
        eng.set_tuning(1, mk)
        q, _ = eng.part2(*args, pd, pairs=prd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.part2(*args, pd, pairs=prd)
        e1.record(); torch.cuda.synchronize()
        qq = q.cpu().numpy()[rows]
        print(f"{name} split_min_k={mk}: err vs oracle f32 {np.abs(qq - want).max():.2e}  vs f64 {np.abs(qq - w64).max():.2e}  PartII {e0.elapsed_time(e1) / 10:.3f} ms")
    eng.set_tuning(1, 1024)
