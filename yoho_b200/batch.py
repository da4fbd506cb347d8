"""Scene-level driver (BASELINE.json configs 3-4: a 3DMatch-shaped set of fragments and pairs, sharded over the ranks).

The reference runs PartI once per FRAGMENT (tests/extractor.py:46-47) and everything else once per PAIR
(tests/matcher.py:30, tests/extractor.py:91,162, tests/estimator.py:91,305); pairs never cross scenes
(utils/dataset.py:167: one dataset object per scene).  Here, with W ranks:

  plan     `plan_scene`: fragments are ordered scene by scene and cut into W contiguous blocks of equal estimated cost
           (a fragment's PartI + the pairs anchored at it); a pair belongs to the rank that owns its earlier fragment.  Only
           the pairs that straddle a cut need a fragment another rank owns.
  phase 1  PartI once per fragment on its owner.
  exchange the PartI outputs (eqv [K,32,60] + matcher descriptor [K,32]) of exactly those straddling fragments travel
           owner -> user with one batched NCCL send/recv group.  No all-gather: a scene-aware cut moves a few fragments per
           cut instead of landing every fragment on every rank (433 x 38.4 MB = 16.6 GB per rank for config 3).  The owner
           computes those fragments FIRST and posts the group at once; it progresses on NCCL's stream behind the rest of
           phase 1 and the all-local pairs of phase 2, and only the straddling pairs wait for it.
  phase 2  every rank registers its pairs from the cached PartI outputs with the split-phase pair call (pair i+1's matching is
           queued before the host waits for pair i's match count).
  gather   one tiny all-gather of the [n_pairs, 2, 3, 4] float64 transforms.

Inputs (FCGF group features, keypoints) are "on disk" for every rank: `fragments` is a dict or a loader callable, and a rank
loads the inputs of every fragment its own pairs touch.
"""
import numpy as np
import torch
import torch.distributed as tdist

from . import dist as ydist
from .pipeline import PairPipeline

# estimated device milliseconds (5000-keypoint fragments, measured on B200: profiles/README.md); only their RATIO matters
FRAG_COST_MS = 2.15
PAIR_COST_MS = 1.15


class ScenePlan:
    """owner_f {fid: rank}, owner_p [rank per pair], order [fid in cut order], transfers [(src, dst, fid)], cost [ms per rank]."""

    def __init__(self, owner_f, owner_p, order, transfers, cost, scene_of):
        self.owner_f, self.owner_p, self.order, self.transfers, self.cost, self.scene_of = owner_f, owner_p, order, transfers, cost, scene_of

    @property
    def balance(self):
        """mean / max of the estimated per-rank cost = the scaling efficiency the partition allows."""
        return float(np.mean(self.cost) / max(np.max(self.cost), 1e-12))

    def frags_of(self, r):
        return [f for f in self.order if self.owner_f[f] == r]

    def pairs_of(self, r):
        return [i for i, o in enumerate(self.owner_p) if o == r]


def _components(frag_ids, pair_ids):
    parent = {f: f for f in frag_ids}

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for a, b in pair_ids:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[rb] = ra
    return {f: find(f) for f in frag_ids}


def plan_scene(frag_ids, pair_ids, world, scene_of=None, frag_cost=FRAG_COST_MS, pair_cost=PAIR_COST_MS):
    """Deterministic (every rank computes the same plan).  frag_ids: list in dataset order; pair_ids: [(id0, id1)];
    scene_of: optional {fid: scene key} (default: connected components of the pair graph — pairs never cross scenes)."""
    frag_ids = list(frag_ids)
    pos0 = {f: i for i, f in enumerate(frag_ids)}
    comp = scene_of or _components(frag_ids, pair_ids)
    first = {}
    for f in frag_ids:
        first.setdefault(comp[f], pos0[f])
    order = sorted(frag_ids, key=lambda f: (first[comp[f]], pos0[f]))           # scene by scene, dataset order inside
    pos = {f: i for i, f in enumerate(order)}
    anchor = [a if pos[a] <= pos[b] else b for a, b in pair_ids]
    w_f = {f: float(frag_cost) for f in order}
    for an in anchor:
        w_f[an] += float(pair_cost)
    total = sum(w_f.values())
    owner_f, cost = {}, [0.0] * world
    r, acc = 0, 0.0
    for i, f in enumerate(order):
        # move on when this rank has reached its share of the cumulative cost (midpoint rule), keeping at least one fragment for
        # every remaining rank
        left = len(order) - i
        if r < world - 1 and (acc + 0.5 * w_f[f] > total * (r + 1) / world or left <= world - 1 - r) and cost[r] > 0:
            r += 1
        owner_f[f] = r
        cost[r] += w_f[f]
        acc += w_f[f]
    owner_p = [owner_f[an] for an in anchor]
    need = set()
    for (a, b), op in zip(pair_ids, owner_p):
        for f in (a, b):
            if owner_f[f] != op:
                need.add((owner_f[f], op, pos[f]))
    transfers = [(s, d, order[p]) for s, d, p in sorted(need)]
    return ScenePlan(owner_f, owner_p, order, transfers, cost, comp)


def exchange_part1_start(plan, local, template, device, rank=None):
    """Post the sends / receives of the PartI outputs of the fragments of `plan.transfers` that involve this rank.
    local: {fid: (eqv, desc)} holding at least the fragments this rank SENDS; template(fid) -> K (rows of that fragment).
    Returns ({fid: (eqv, desc)} receive buffers, requests, bytes to receive).  One batched isend/irecv group (NCCL over NVLink on
    the box): it progresses on NCCL's own stream while the caller keeps computing; `req.wait()` before touching the buffers."""
    rank = ydist.rank() if rank is None else rank
    ops, got = [], {}
    nbytes = 0
    for s, d, f in plan.transfers:
        if s == rank:
            e, ds = local[f]
            ops.append(tdist.P2POp(tdist.isend, e, d))
            ops.append(tdist.P2POp(tdist.isend, ds, d))
        elif d == rank:
            K = template(f)
            e = torch.empty((K, 32, 60), dtype=torch.float32, device=device)
            ds = torch.empty((K, 32), dtype=torch.float32, device=device)
            ops.append(tdist.P2POp(tdist.irecv, e, s))
            ops.append(tdist.P2POp(tdist.irecv, ds, s))
            got[f] = (e, ds)
            nbytes += e.numel() * 4 + ds.numel() * 4
    reqs = tdist.batch_isend_irecv(ops) if ops else []
    return got, reqs, nbytes


def warmup_exchange(plan, device, rank=None):
    """One 4-byte send/recv per (source, destination) pair of the plan: NCCL sets up its point-to-point channels lazily at the
    first transfer between two ranks (~1 s measured), which a benchmark keeps out of its timed region with this call."""
    rank = ydist.rank() if rank is None else rank
    ops, keep = [], []
    for s, d in sorted({(s, d) for s, d, _ in plan.transfers}):
        if s == rank or d == rank:
            t = torch.zeros((1,), dtype=torch.float32, device=device)
            keep.append(t)
            ops.append(tdist.P2POp(tdist.isend if s == rank else tdist.irecv, t, d if s == rank else s))
    if ops:
        for req in tdist.batch_isend_irecv(ops):
            req.wait()
    if torch.cuda.is_available() and device.type == "cuda":
        torch.cuda.synchronize()


def exchange_part1(plan, local, template, device, rank=None):
    """Blocking form of `exchange_part1_start`: returns ({fid: (eqv, desc)} received, bytes received)."""
    got, reqs, nbytes = exchange_part1_start(plan, local, template, device, rank)
    for req in reqs:
        req.wait()
    return got, nbytes


def register_scene(pipe: PairPipeline, fragments, pair_ids, timing=None, frag_ids=None, scene_of=None, plan=None):
    """fragments: dict id -> (feat [K,32,60] f32, kps [K,3] f64) as numpy or CUDA tensors, or a callable id -> that tuple (a
    loader: called only for the fragments this rank needs); pair_ids: list of (id0, id1); frag_ids: dataset order of the
    fragment ids (default: order of first appearance in pair_ids).
    Returns a CUDA tensor [n_pairs, 2, 3, 4] (YOHO-C, YOHO-O transform per pair, in `pair_ids` order) on every rank.
    `timing` (optional dict) receives this rank's device milliseconds per phase, the exchange volume and the plan's balance."""
    eng = pipe.eng
    dev = eng.device
    w, rk = ydist.world(), ydist.rank()
    if frag_ids is None:
        frag_ids = list(dict.fromkeys(f for p in pair_ids for f in p))
    plan = plan or plan_scene(frag_ids, pair_ids, w, scene_of=scene_of)
    load = fragments if callable(fragments) else (lambda f: fragments[f])
    my_pairs = plan.pairs_of(rk)
    my_frags = plan.frags_of(rk)
    touched = list(dict.fromkeys([f for f in my_frags] + [f for i in my_pairs for f in pair_ids[i]]))
    inputs = {}
    for f in touched:                                   # inputs resident before the clock starts (like bench.py's `value`)
        feat, kps = load(f)
        inputs[f] = (eng._f32(feat), eng._f64(kps))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timing is not None else None
    if ev:
        torch.cuda.synchronize()
        if w > 1:
            tdist.barrier()
        ev[0].record()
    # phase 1: PartI once per fragment, on its owner (tests/extractor.py:46-47) — the fragments other ranks are waiting for first,
    # so that their transfer is posted early and travels (NCCL's own stream) behind the rest of this rank's work
    cache = {}
    outgoing = list(dict.fromkeys(f for s_, d_, f in plan.transfers if s_ == rk))
    incoming = {f for s_, d_, f in plan.transfers if d_ == rk}

    def part1(f):
        o = eng.part1(inputs[f][0], want_inv=False, want_desc=True)
        cache[f] = (o["eqv"], o["desc"])
    for f in outgoing:
        part1(f)
    reqs, nbytes, got = [], 0, {}
    if w > 1 and plan.transfers:
        got, reqs, nbytes = exchange_part1_start(plan, cache, lambda f: inputs[f][0].shape[0], dev, rk)
    for f in my_frags:
        if f not in cache:
            part1(f)
    if ev:
        ev[1].record()
    # phase 2: everything else once per pair; the hypothesis draws are seeded by the pair's position in `pair_ids`, so the
    # result does not depend on the sharding.  Pairs whose fragments are all local run first; the others after the receives.
    out = torch.zeros((len(my_pairs), 2, 3, 4), dtype=torch.float64, device=dev)
    order = [n for n, pi in enumerate(my_pairs) if not (set(pair_ids[pi]) & incoming)]
    late = [n for n, pi in enumerate(my_pairs) if set(pair_ids[pi]) & incoming]
    pend = []

    def finish(tok_n):
        tok, n = tok_n
        t = eng.register_pair_end(tok)
        out[n] = t["T_co"]

    def run(ns):
        for n in ns:
            pi = my_pairs[n]
            a, b = pair_ids[pi]
            if pipe.fused:
                tok = eng.register_pair_begin(inputs[a][0], inputs[b][0], inputs[a][1], inputs[b][1], pipe.c_iters, pipe.o_iters,
                                              pipe.c_dist, pipe.o_dist, pipe.seed + 1 + pi, eqvA=cache[a][0], eqvB=cache[b][0],
                                              descA=cache[a][1], descB=cache[b][1])
                pend.append((tok, n))
                if len(pend) > 1:                       # pair i+1's matching is queued before the host waits for pair i's count
                    finish(pend.pop(0))
            else:
                r = pipe.register(inputs[a][0], inputs[b][0], inputs[a][1], inputs[b][1], eqvA=cache[a][0], eqvB=cache[b][0],
                                  descA=cache[a][1], descB=cache[b][1], seed=pipe.seed + 1 + pi)
                out[n, 0], out[n, 1] = r["T_c"], r["T_o"]
    run(order)
    if ev:
        ev[2].record()
    for req in reqs:                                    # the compute stream now waits for whatever has not landed yet
        req.wait()
    cache.update(got)
    if ev:
        ev[3].record()
    run(late)
    while pend:
        finish(pend.pop(0))
    if ev:
        ev[4].record()
        torch.cuda.synchronize()
        timing.update(part1_ms=ev[0].elapsed_time(ev[1]), pairs_ms=ev[1].elapsed_time(ev[2]) + ev[3].elapsed_time(ev[4]),
                      exchange_stall_ms=ev[2].elapsed_time(ev[3]), total_ms=ev[0].elapsed_time(ev[4]), fragments_owned=len(my_frags),
                      fragments_touched=len(touched), pairs=len(my_pairs), pairs_after_exchange=len(late),
                      exchange_bytes_received=int(nbytes), transfers_total=len(plan.transfers), plan_balance=plan.balance)
    if w == 1:
        full = torch.zeros((len(pair_ids), 2, 3, 4), dtype=torch.float64, device=dev)
        if my_pairs:
            full[torch.as_tensor(my_pairs, device=dev)] = out
        return full
    return ydist.gather_rows(out.reshape(len(my_pairs), 24), my_pairs, len(pair_ids)).reshape(len(pair_ids), 2, 3, 4)
