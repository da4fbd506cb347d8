"""Timing sweep of the tensor-core group convolution (one layer in isolation) over implementation / tuning flags.
Run on the GPU box:  python tools/tc_sweep.py > gpurun_out/tc_sweep.log"""
import itertools
import os
import sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yoho_b200 import synth
from yoho_b200.engine import get_engine

eng = get_engine()
eng.load_part1(synth.synth_state_dict("PartI", 0))
B = 2048
rs = np.random.RandomState(0)
layers = {1: (256, 512), 2: (512, 256), 0: (32, 256)}
acts = {l: torch.from_numpy(np.maximum(rs.standard_normal((B, 60, cin)), 0).astype(np.float32)).cuda() for l, (cin, _) in layers.items()}
ref = {l: eng.debug_layer(l, "simt", acts[l], layers[l][1]) for l in layers}


def time_it(fn, reps=6):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


print("layer impl flags ms TFLOPs(alg) maxerr_vs_simt")
for l, (cin, cout) in (layers.items() if "--layers" in sys.argv else []):
    flops = 2.0 * B * 60 * 13 * cin * cout
    for impl, flags in itertools.product(["tcgen05", "tcgen05_split"], [0, 1, 2, 3]):
        eng.set_tuning(0, flags)
        out = eng.debug_layer(l, impl, acts[l], cout)
        err = float((out - ref[l]).abs().max())
        # per-launch device time of the convolution kernel alone (events around the launch inside the library)
        eng.profile(True)
        for _ in range(6):
            eng.debug_layer(l, impl, acts[l], cout)
        pr = eng.profile_read()[l]
        eng.profile(False)
        ms = pr["ms"] / max(pr["launches"], 1)
        print(f"{l} {impl} flags={flags} {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s(alg)  err={err:.3e}", flush=True)
eng.set_tuning(0, 2)

# ---- whole PartI forward per implementation (5000 keypoints) ------------------------------------------------------
x = torch.from_numpy(synth.make_fragment(5000, 3)[0]).cuda()
for impl, flags in [("tcgen05_split", 3), ("tcgen05_fourier", 3 | 256), ("tcgen05_fourier", 3), ("tcgen05_fourier", 3 | 64), ("tcgen05_fourier", 3 | 128), ("tcgen05_fourier", 3 | 32)]:
    eng.set_gconv_impl(impl)
    eng.set_tuning(0, flags)
    ms = time_it(lambda: eng.part1(x, want_inv=False))
    eng.profile(True)
    eng.part1(x, want_inv=False)
    pr = eng.profile_read()
    eng.profile(False)
    print(f"part1 5000 kpts {impl} flags={flags}: {ms:.3f} ms; " + ", ".join(f"{q['name']}={q['ms']:.3f}" for q in pr if q['launches']), flush=True)
eng.set_tuning(0, eng.DEFAULT_TUNING)
eng.set_gconv_impl("tcgen05_fourier")
