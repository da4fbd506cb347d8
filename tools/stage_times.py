"""Per-stage wall time (synchronised) of the per-pair tail on a few pairs of a synthetic scene, next to the match count M.

    python tools/stage_times.py [c3|c4]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                       # noqa: E402
from yoho_b200 import synth                        # noqa: E402
from yoho_b200.engine import get_engine            # noqa: E402

key = sys.argv[1] if len(sys.argv) > 1 else "c3"
lo, hi = (0.55, 0.95) if key == "c3" else (0.32, 0.55)
e = get_engine()
e.load_part1(synth.synth_state_dict("PartI", 0))
e.load_part2(synth.synth_state_dict("PartII", 0))
dev = e.device
frags, pair_ids, _ = synth.make_scene(8, 5000, seed=7, overlap_lo=lo, overlap_hi=hi)
d = {k: (torch.from_numpy(f).to(dev), torch.from_numpy(p).to(dev)) for k, (f, p) in frags.items()}
p1 = {k: e.part1(d[k][0], want_inv=False, want_desc=True) for k in d}


def tm(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


for rep in range(2):
    for (a, b) in pair_ids[:10]:
        (pb, nd), t_nn = tm(lambda: e.mutual_nn(p1[a]["desc"], p1[b]["desc"]))
        M = int(nd.item())
        pairs = pb[:M]
        dr, t_rot = tm(lambda: e.rot_argmax(p1[b]["eqv"], p1[a]["eqv"], pairs=pairs))
        (k0, k1), t_g = tm(lambda: e.gather_kps(d[a][1], d[b][1], pairs))
        (hyp, st), t_d = tm(lambda: e.c_draw(dr, 1000, 5))
        rc, t_c = tm(lambda: e.c_ransac(k0, k1, hyp, 0.07))
        (q, tr), t_p2 = tm(lambda: e.part2(d[a][0], d[b][0], p1[a]["eqv"], p1[b]["eqv"], dr, pairs=pairs, kps0=d[a][1], kps1=d[b][1]))
        order, t_o = tm(lambda: e.o_order(M, 5))
        ro, t_s = tm(lambda: e.o_score(k0, k1, tr, 0.09, order=order, max_hyp=1000))
        if rep:
            print(f"M={M:5d} nn {t_nn:.3f} rot {t_rot:.3f} gather {t_g:.3f} draw {t_d:.3f} yohoc {t_c:.3f} part2 {t_p2:.3f} order {t_o:.3f} yohoo {t_s:.3f}  "
                  f"sum {t_nn + t_rot + t_g + t_d + t_c + t_p2 + t_o + t_s:.3f} ms", flush=True)
