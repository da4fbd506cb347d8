"""PartI on one 5000-keypoint fragment, a few times (a target for ncu: `-k regex:"gconv_tc|group_transform" --launch-skip 14
--launch-count 7` captures the seven tensor-core launches of the third pass).  Prints the per-layer device times.

    python tools/part1_once.py [kpts] [passes]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                      # noqa: E402
from yoho_b200 import synth                        # noqa: E402
from yoho_b200.engine import get_engine            # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = get_engine()
eng.set_gconv_impl(os.environ.get("YOHO_B200_GCONV", "tcgen05_fourier"))
eng.load_part1(synth.synth_state_dict("PartI", 2))
if os.environ.get("YOHO_B200_TUNING"):
    eng.set_tuning(0, int(os.environ["YOHO_B200_TUNING"]))
x, _ = synth.make_fragment(K, 7)
xd = torch.from_numpy(x).to(eng.device)
eng.profile(True)
for _ in range(passes):
    eng.part1(xd, want_inv=False, want_desc=True)
torch.cuda.synchronize()
pr = eng.profile_read()
eng.profile(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    eng.part1(xd, want_inv=False, want_desc=True)
e1.record()
torch.cuda.synchronize()
print(f"part1 {K} kpts x{passes}: " + ", ".join(f"{q['name']}={q['ms'] / passes:.3f}" for q in pr if q["launches"])
      + f" | whole PartI {e0.elapsed_time(e1) / reps:.3f} ms (mean of {reps}, back to back)")
