"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on the box, gloo in the CPU tests).

The hot path shards over INDEPENDENT units: PartI per fragment (tests/extractor.py:46), everything
else per pair (tests/matcher.py:30, tests/extractor.py:91,162, tests/estimator.py:91,305) — the reference itself
parallelises pairs with a process pool (tests/estimator.py:269-273).  So there is no data-path collective:
  * `shard(items)`            round-robin assignment of fragments / pairs to ranks,
  * `gather_transforms(T)`    the one tiny exchange: every rank's [n_i,3,4] float64 results -> rank order,
  * `allgather_sharded(...)`  scene driver only: PartI outputs computed once per fragment across the job, one all-gather.
Config 5 (a 10 000-keypoint pair whose fragment-1 descriptors are sharded over the ranks) does have an exchange:
  * `sharded_mutual_nn(...)`  each rank searches all of A against its shard of B with the single-GPU kernel, then one
                              all-gather of the packed (distance, index) keys; the lexicographic min over ranks keeps
                              torch.min's lowest-index tie-break (utils/knn_search.py:41).
"""
import os
import numpy as np
import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  Returns (rank, local_rank, world)."""
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, local_rank, world


def world():
    return dist.get_world_size() if dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_initialized() else 0


def shard(items, r=None, w=None):
    """Round-robin: rank r owns items[r::w].  Deterministic, balanced to within one item."""
    r = rank() if r is None else r
    w = world() if w is None else w
    return list(items)[r::w]


def unshard(per_rank_lists):
    """Inverse of `shard` over the gathered per-rank lists: restores the original item order."""
    w = len(per_rank_lists)
    n = sum(len(x) for x in per_rank_lists)
    out = [None] * n
    for r, lst in enumerate(per_rank_lists):
        for i, v in enumerate(lst):
            out[r + i * w] = v
    return out


def gather_transforms(T_local):
    """T_local [n_local,3,4] float64 on this rank -> [n_total,3,4] in the original (un-sharded) pair order."""
    T_local = T_local.reshape(-1, 3, 4)
    if world() == 1:
        return T_local
    w = world()
    n = torch.tensor([T_local.shape[0]], device=T_local.device, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    cap = max(ns)
    buf = torch.zeros((cap, 3, 4), device=T_local.device, dtype=T_local.dtype)
    buf[: T_local.shape[0]] = T_local
    bufs = [torch.zeros_like(buf) for _ in range(w)]
    dist.all_gather(bufs, buf)
    per_rank = [list(bufs[r][: ns[r]]) for r in range(w)]
    rows = unshard(per_rank)
    return torch.stack(rows) if rows else T_local


def gather_rows(rows_local, index_local, n_total):
    """rows_local [n_local, C] on this rank, index_local = the global row number of each -> [n_total, C] on every rank
    (the scene driver's result gather: pairs are owned by arbitrary ranks).  Two small all-gathers (padded rows + indices)."""
    w = world()
    dev, dt = rows_local.device, rows_local.dtype
    C = rows_local.shape[1]
    idx = torch.as_tensor(list(index_local), dtype=torch.int64, device=dev)
    out = torch.zeros((n_total, C), dtype=dt, device=dev)
    if w == 1:
        if idx.numel():
            out[idx] = rows_local
        return out
    n = torch.tensor([idx.numel()], device=dev, dtype=torch.int64)
    ns = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    cap = max(max(ns), 1)
    buf = torch.zeros((cap, C), device=dev, dtype=dt)
    ibuf = torch.full((cap,), -1, device=dev, dtype=torch.int64)
    buf[: idx.numel()] = rows_local
    ibuf[: idx.numel()] = idx
    bufs = [torch.zeros_like(buf) for _ in range(w)]
    ibufs = [torch.zeros_like(ibuf) for _ in range(w)]
    dist.all_gather(bufs, buf)
    dist.all_gather(ibufs, ibuf)
    for r in range(w):
        if ns[r]:
            out[ibufs[r][: ns[r]]] = bufs[r][: ns[r]]
    return out


def allgather_sharded(items_local, n_total):
    """Every rank holds the tensors of items r, r+w, r+2w, ... (the `shard` order) of a list of n_total equally shaped
    tensors; returns the full list, in item order, on every rank.  ONE all-gather of the stacked (zero-padded) shards — used
    after phase 1 of the scene driver so that PartI runs once per fragment across the whole job (38 MB per 5000-keypoint
    fragment over NVLink) instead of once per rank."""
    w = world()
    if w == 1:
        return list(items_local)
    cap = (n_total + w - 1) // w
    assert 0 < len(items_local) <= cap, "every rank must own at least one item (n_total >= world size)"
    ref = items_local[0]
    shape, dt, dev = ref.shape, ref.dtype, ref.device
    buf = torch.zeros((cap,) + tuple(shape), dtype=dt, device=dev)
    for i, t in enumerate(items_local):
        buf[i].copy_(t)
    out = torch.empty((w * cap,) + tuple(shape), dtype=dt, device=dev)
    dist.all_gather_into_tensor(out, buf)
    return [out[(i % w) * cap + i // w] for i in range(n_total)]


def pack_key(dist_f32, idx):
    """(float32 distance >= 0, index) -> int64 key whose order is the lexicographic (distance, index) order."""
    bits = dist_f32.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (bits << 32) | idx.to(torch.int64)


def unpack_key(key):
    d = (key >> 32).to(torch.int32).view(torch.float32)
    return d, key & 0xFFFFFFFF


def sharded_mutual_nn(dA, dB_local, b_offset, Kb_total, nn1_fn):
    """Cross-rank mutual 1-NN for one pair: dA [Ka,32] replicated, dB_local = this rank's rows
    [b_offset, b_offset+len) of fragment 1.  nn1_fn(source, target) -> (dist f32 [m], idx i64 [m]) is the
    single-device search (Engine.nn1 on the box).  Returns int64 [M,2] matches on every rank."""
    Ka = dA.shape[0]
    w = world()
    # A -> B: best over my shard, then lexicographic min over ranks
    d_loc, i_loc = nn1_fn(dA, dB_local)
    key = pack_key(d_loc, i_loc + b_offset)
    if w > 1:
        keys = [torch.zeros_like(key) for _ in range(w)]
        dist.all_gather(keys, key)
        key = torch.stack(keys).min(dim=0).values
    _, nnA = unpack_key(key)                                   # [Ka] index into all of B
    # B -> A: every B row lives on exactly one rank
    _, j_loc = nn1_fn(dB_local, dA)
    nnB = torch.full((Kb_total,), -1, dtype=torch.int64, device=dA.device)
    nnB[b_offset: b_offset + dB_local.shape[0]] = j_loc
    if w > 1:
        dist.all_reduce(nnB, op=dist.ReduceOp.MAX)
    a = torch.arange(Ka, device=dA.device)
    keep = nnB[nnA] == a
    return torch.stack([a[keep], nnA[keep]], 1)
