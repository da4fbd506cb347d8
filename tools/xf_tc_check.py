"""Bring-up check of the tcgen05 transform kernel (tuning flag 256; 512 swaps the LBO/SBO fields of its MN-major descriptor):
each variant runs in its own process (a wrong descriptor may poison the CUDA context), prints the max |difference| of the
PartI output against the default path and the per-layer timing on 5000 keypoints.

    python tools/xf_tc_check.py            # parent: spawns the variants
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(flags):
    import numpy as np
    import torch
    from yoho_b200 import synth
    from yoho_b200.engine import get_engine
    eng = get_engine()
    eng.set_gconv_impl("tcgen05_fourier")
    eng.load_part1(synth.synth_state_dict("PartI", 2))
    x, _ = synth.make_fragment(300, 41)
    eng.set_tuning(0, 3)
    a = eng.part1(x)["eqv"].clone()
    eng.set_tuning(0, flags)
    b = eng.part1(x)["eqv"].clone()
    torch.cuda.synchronize()
    d = (a - b).abs().max().item()
    print(f"flags={flags}: max|eqv - default| = {d:.3e}  (ref max {a.abs().max().item():.3f})", flush=True)
    x5, _ = synth.make_fragment(5000, 7)
    xd = torch.from_numpy(x5).to(eng.device)
    for f in (3, 3 | 256 | 512, flags):
        eng.set_tuning(0, f)
        for _ in range(2):
            eng.part1(xd)
        eng.profile(True)
        for _ in range(5):
            eng.part1(xd)
        torch.cuda.synchronize()
        pr = eng.profile_read()
        eng.profile(False)
        print(f"  5000 kpts flags={f}: " + ", ".join(f"{q['name']}={q['ms'] / 5:.3f}" for q in pr if q["launches"]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
    else:
        for flags in (3 | 256,):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(flags)], capture_output=True, text=True, timeout=240)
                print(r.stdout.strip())
                if r.returncode != 0:
                    print(f"flags={flags}: rc={r.returncode}\n" + r.stderr.strip()[-1500:])
            except subprocess.TimeoutExpired:
                print(f"flags={flags}: TIMEOUT")
