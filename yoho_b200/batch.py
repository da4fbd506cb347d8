"""Scene-level driver (BASELINE.json configs 3-4: a 3DMatch-shaped set of fragments and pairs, sharded over the ranks).

The reference runs PartI once per FRAGMENT (tests/extractor.py:46-47) and everything else once per PAIR
(tests/matcher.py:30, tests/extractor.py:91,162, tests/estimator.py:91,305).  Here:
  phase 1  every rank runs PartI on the fragments its pairs touch (no collective; a fragment shared by pairs on different
           ranks is recomputed rather than exchanged — 38 MB over NVLink would also do, PartI is 5 ms),
  phase 2  every rank registers its round-robin share of the pairs from the cached eqv / matcher descriptors,
  phase 3  one tiny gather of the [n_pairs, 2, 3, 4] float64 transforms.
"""
import numpy as np
import torch

from . import dist as ydist
from .pipeline import PairPipeline


def register_scene(pipe: PairPipeline, fragments, pair_ids, timing=None):
    """fragments: dict id -> (feat [K,32,60] f32, kps [K,3] f64) as numpy or CUDA tensors; pair_ids: list of (id0, id1).
    Returns a CUDA tensor [n_pairs, 2, 3, 4] (YOHO-C, YOHO-O transform per pair, in `pair_ids` order) on every rank.
    `timing` (optional dict) receives the device milliseconds of phase 1 (PartI) and phase 2 (pairs) of this rank."""
    eng = pipe.eng
    dev = eng.device
    mine = ydist.shard(list(range(len(pair_ids))))
    cache = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timing is not None else None
    if ev:
        ev[0].record()
    # phase 1: PartI once per fragment this rank's pairs touch (tests/extractor.py:46-47)
    for pi in mine:
        for fid in pair_ids[pi]:
            if fid not in cache:
                feat, kps = fragments[fid]
                feat = eng._f32(feat)
                kps = eng._f64(kps)
                o = eng.part1(feat, want_inv=False, want_desc=True)
                cache[fid] = (feat, kps, o["eqv"], o["desc"])
    if ev:
        ev[1].record()
    # phase 2: everything else once per pair
    out = torch.zeros((len(mine), 2, 3, 4), dtype=torch.float64, device=dev)
    for n, pi in enumerate(mine):
        a, b = pair_ids[pi]
        fa, ka, ea, da = cache[a]
        fb, kb, eb, db = cache[b]
        # the hypothesis draws are seeded by the pair's position, so the result does not depend on the sharding
        r = pipe.register(fa, fb, ka, kb, eqvA=ea, eqvB=eb, descA=da, descB=db, seed=pipe.seed + 1 + pi)
        out[n, 0], out[n, 1] = r["T_c"], r["T_o"]
    if ev:
        ev[2].record()
        torch.cuda.synchronize()
        timing["part1_ms"] = ev[0].elapsed_time(ev[1])
        timing["pairs_ms"] = ev[1].elapsed_time(ev[2])
        timing["fragments"] = len(cache)
        timing["pairs"] = len(mine)
    if ydist.world() == 1:
        return out
    # gather per transform kind so that `gather_transforms` can restore the pair order
    tc = ydist.gather_transforms(out[:, 0].contiguous())
    to = ydist.gather_transforms(out[:, 1].contiguous())
    return torch.stack([tc, to], dim=1)
