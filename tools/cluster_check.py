"""Bring-up check of the 2-CTA-cluster weight multicast of the grouped tensor-core launch (tuning flag 32768, csrc/gconv_tc.cu):
PartI outputs must equal the default launch BIT FOR BIT (same products, same accumulation order), at ragged and full sizes;
prints whole-PartI time for both.

    python tools/cluster_check.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                      # noqa: E402
from yoho_b200 import synth                        # noqa: E402
from yoho_b200.engine import get_engine            # noqa: E402

eng = get_engine()
eng.set_gconv_impl("tcgen05_fourier")
eng.load_part1(synth.synth_state_dict("PartI", 2))


def run(K, flag, reps=20):
    x, _ = synth.make_fragment(K, 7 + K)
    xd = torch.from_numpy(x).to(eng.device)
    eng.set_tuning(0, eng.DEFAULT_TUNING | flag)
    o = eng.part1(xd, want_inv=True, want_desc=True)
    out = {k: o[k].clone() for k in ("eqv", "inv", "desc")}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.part1(xd, want_inv=False, want_desc=True)
    e1.record()
    torch.cuda.synchronize()
    eng.set_tuning(0, eng.DEFAULT_TUNING)
    return out, e0.elapsed_time(e1) / reps


bad = 0
for K in (1, 3, 129, 600, 2101, 5000, 10000):
    a, ta = run(K, 0)
    b, tb = run(K, 32768)
    same = all(torch.equal(a[k].view(torch.int32), b[k].view(torch.int32)) for k in a)
    bad += 0 if same else 1
    print(f"K={K}: identical={same}  default {ta:.3f} ms  cluster {tb:.3f} ms", flush=True)
for rnd in range(3):                              # interleaved timing at the benchmark sizes
    for K in (5000, 10000):
        _, ta = run(K, 0, 40)
        _, tb = run(K, 32768, 40)
        print(f"round {rnd} K={K}: default {ta:.3f} ms  cluster {tb:.3f} ms", flush=True)
print("CLUSTER CHECK", "OK" if bad == 0 else "FAILED")
sys.exit(1 if bad else 0)
