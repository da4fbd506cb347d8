"""Where does the end-to-end time go?  Times, on one GPU, 20 cold 5000-keypoint pairs (a) device resident, (b) through
`register_pinned` (one pair at a time: H2D, pipeline, D2H), (c) through `register_stream` (prefetching), plus the raw H2D rate
of one pinned fragment.

    python tools/e2e_check.py [steps]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                      # noqa: E402
from yoho_b200 import synth                        # noqa: E402
from yoho_b200.engine import get_engine            # noqa: E402
from yoho_b200.pipeline import PairPipeline        # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = get_engine()
eng.load_part1(synth.synth_state_dict("PartI", 0))
eng.load_part2(synth.synth_state_dict("PartII", 0))
dev = eng.device
N = 4
pairs = [synth.make_fragment_pair(5000, seed=s, overlap=0.5, sigma=0.05) for s in range(N)]
sets_d = [tuple(torch.from_numpy(p[k]).to(dev) for k in ("feat_A", "feat_B", "kps_A", "kps_B")) for p in pairs]
sets_p = [PairPipeline.pin(p["feat_A"], p["feat_B"], p["kps_A"], p["kps_B"]) for p in pairs]
pipe = PairPipeline(eng, seed=0)


def timed(fn, label):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1) / steps:.3f} ms/pair (events), {(time.perf_counter() - t0) * 1e3 / steps:.3f} ms/pair (wall)", flush=True)


for rep in range(2):
    timed(lambda: [pipe.register(*sets_d[i % N]) for i in range(steps)], "device-resident register")
    timed(lambda: [pipe.register_pinned(*sets_p[i % N]) for i in range(steps)], "register_pinned          ")
    timed(lambda: [0 for _ in pipe.register_stream(sets_p[i % N] for i in range(steps))], "register_stream          ")
x = sets_p[0][0]
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    y = x.to(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"H2D {x.numel() * 4 / 1e6:.1f} MB pinned: {dt * 1e3:.3f} ms = {x.numel() * 4 / dt / 1e9:.1f} GB/s")
