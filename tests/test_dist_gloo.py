"""CPU, world_size 2, gloo: the N>1 host logic — sharding, the result gather and the cross-rank mutual-NN merge
(config 5) — with the oracle standing in for the single-device search."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from yoho_b200 import dist as yd
    import yoho_oracle as O
    r, lr, w = yd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    try:
        # 1. sharding + gather restores the original pair order
        pairs = list(range(7))
        mine = yd.shard(pairs)
        T = torch.zeros((len(mine), 3, 4), dtype=torch.float64)
        for i, p in enumerate(mine):
            T[i] = float(p)
        allT = yd.gather_transforms(T)
        assert allT.shape == (7, 3, 4) and [int(allT[i, 0, 0]) for i in range(7)] == pairs
        # 2. config-5 style sharded mutual NN == single-process result
        rs = np.random.RandomState(0)
        Ka, Kb = 301, 257
        dA = (rs.standard_normal((Ka, 32)) * 0.1).astype(np.float32)
        dB = (rs.standard_normal((Kb, 32)) * 0.1).astype(np.float32)
        dB[:120] = dA[rs.permutation(Ka)[:120]] + (rs.standard_normal((120, 32)) * 0.01).astype(np.float32)
        dB[200:210] = dB[100:110]                       # exact duplicates across the shard boundary: lowest index wins
        bounds = np.linspace(0, Kb, world + 1).astype(int)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        nn1 = lambda s, t: O.nn1(s, t)
        got = yd.sharded_mutual_nn(torch.from_numpy(dA), torch.from_numpy(dB[lo:hi]), lo, Kb, nn1).numpy()
        want, _, _ = O.mutual_matches(dA, dB)
        assert np.array_equal(got, want), (got.shape, want.shape)
        # 3. scene-driver exchange: items sharded round-robin come back complete and in order on every rank
        for n_items in (2, 5):
            mine_i = yd.shard(list(range(n_items)))
            got_all = yd.allgather_sharded([torch.full((3, 4), float(i)) for i in mine_i], n_items)
            assert [float(t[0, 0]) for t in got_all] == [float(i) for i in range(n_items)] and tuple(got_all[0].shape) == (3, 4)
        # 4. scene plan + PartI-output exchange + result gather (yoho_b200.batch, BASELINE.json configs 3-4) with CPU tensors
        from yoho_b200 import batch as yb
        frag_ids = list(range(12))
        pair_ids = [(i, i + 1) for i in range(5)] + [(0, 5), (2, 4)] + [(i, i + 1) for i in range(6, 11)] + [(6, 11), (5, 4)]
        plan = yb.plan_scene(frag_ids, pair_ids, world)
        assert sorted(plan.owner_f) == frag_ids and set(plan.owner_f.values()) == set(range(world))
        mine_f = plan.frags_of(rank)
        local = {f: (torch.full((4, 32, 60), float(f)), torch.full((4, 32), float(f) + 0.5)) for f in mine_f}
        got, nbytes = yb.exchange_part1(plan, local, lambda f: 4, torch.device("cpu"), rank)
        want_f = sorted({f for (s_, d_, f) in plan.transfers if d_ == rank})
        assert sorted(got) == want_f and nbytes == len(want_f) * (4 * 32 * 60 + 4 * 32) * 4
        for f, (e, dsc) in got.items():
            assert float(e[0, 0, 0]) == float(f) and float(dsc[0, 0]) == float(f) + 0.5
        have = set(mine_f) | set(got)
        for i in plan.pairs_of(rank):
            assert set(pair_ids[i]) <= have                      # every pair finds both fragments' PartI outputs on its rank
        mine_p = plan.pairs_of(rank)
        rows = torch.stack([torch.full((24,), float(i), dtype=torch.float64) for i in mine_p]) if mine_p else torch.zeros((0, 24), dtype=torch.float64)
        allrows = yd.gather_rows(rows, mine_p, len(pair_ids))
        assert [int(allrows[i, 0]) for i in range(len(pair_ids))] == list(range(len(pair_ids)))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_unshard_roundtrip():
    sys.path.insert(0, ROOT)
    from yoho_b200 import dist as yd
    for n in (0, 1, 5, 16):
        for w in (1, 2, 3, 8):
            items = list(range(n))
            assert yd.unshard([yd.shard(items, r, w) for r in range(w)]) == items
    k = yd.pack_key(torch.tensor([0.5, 0.25, 0.25]), torch.tensor([7, 9, 3]))
    assert int(torch.argmin(k)) == 2
    d, i = yd.unpack_key(k)
    assert d.tolist() == [0.5, 0.25, 0.25] and i.tolist() == [7, 9, 3]


def test_scene_plan_config3_shape_is_balanced_and_scene_aware():
    """The 3DMatch-test shape (433 fragments in 8 scenes, 1623 pairs): the contiguous scene-ordered cut keeps every rank within a
    few percent of the mean estimated cost at 2, 4 and 8 ranks, every pair lands on a rank that owns one of its fragments, pairs
    never cross scenes, and the exchanged fragments are a small fraction of what an all-gather would move."""
    sys.path.insert(0, ROOT)
    from yoho_b200 import synth, batch as yb
    S = synth.SceneSet(synth.THREEDMATCH_SCENE_SIZES, 1623, K=8, seed=0)
    assert len(S.frag_ids) == 433 and len(S.pair_ids) == 1623
    assert all(S.scene_of[a] == S.scene_of[b] for a, b in S.pair_ids)
    for w in (1, 2, 4, 8):
        p = yb.plan_scene(S.frag_ids, S.pair_ids, w, scene_of=S.scene_of)
        assert p.balance >= 0.95, (w, p.balance, p.cost)
        for (a, b), o in zip(S.pair_ids, p.owner_p):
            assert o in (p.owner_f[a], p.owner_f[b])
        assert len(p.transfers) <= 0.25 * 433 * max(w - 1, 0) or w == 1     # all-gather: every fragment to every other rank
        assert len(set(p.transfers)) == len(p.transfers)
    # without scene labels the connected components of the pair graph give the same grouping
    p2 = yb.plan_scene(S.frag_ids, S.pair_ids, 8)
    assert p2.owner_f == yb.plan_scene(S.frag_ids, S.pair_ids, 8, scene_of=S.scene_of).owner_f
