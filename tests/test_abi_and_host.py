"""CPU (no GPU needed): the C-ABI library loads and exports every symbol include/yoho_b200.h declares, fails
loudly without a device, and the host-side logic (group tables, draw order, registries, synthetic generators)
behaves like the reference's."""
import ctypes
import os
import re
import numpy as np
import pytest
import torch

import yoho_oracle as O
from yoho_b200 import _lib, group, synth

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    hdr = open(os.path.join(ROOT, "include", "yoho_b200.h")).read()
    declared = set(re.findall(r"\b(yoho_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"yoho_status"}
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.yoho_abi_version() == 3


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_device_is_a_loud_error_not_a_fallback():
    lib = _lib.load_library()
    t = group.load()
    rot = np.ascontiguousarray(t.R); perm = np.ascontiguousarray(t.P, np.int32); nei = np.ascontiguousarray(t.N, np.int32)
    h = ctypes.c_void_p()
    rc = lib.yoho_ctx_create(0, rot.ctypes.data, perm.ctypes.data, nei.ctypes.data, ctypes.byref(h))
    assert rc == -1 and b"no CPU fallback" in lib.yoho_last_error()
    from yoho_b200.engine import get_engine
    with pytest.raises(_lib.YohoError):
        get_engine()
    from yoho_b200.network import PartI_test
    class C: SO3_related_files = None
    net = PartI_test(C())
    net.load_state_dict(synth.to_torch_state_dict(synth.synth_state_dict("PartI", 0)))
    with pytest.raises(_lib.YohoError):
        net(torch.zeros(2, 32, 60))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yoho_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/estimator_oracle.c", "").replace("oracle/ref_shim", "") \
                    or f in ("build.py",) or "import" not in [l for l in src.splitlines() if "oracle" in l and "import" in l][:1], f
                for line in src.splitlines():
                    if re.match(r"\s*(from|import)\s+.*oracle", line):
                        raise AssertionError(f"{f}: {line}")


def test_group_table_identities():
    t = group.load()
    R, P, N = t.R, t.P, t.N
    assert np.allclose(R[0], np.eye(3)) and np.array_equal(P[0], np.arange(60)) and np.array_equal(P[:, 0], np.arange(60))
    def idx(M):
        d = np.abs(R - M[None]).reshape(60, -1).max(1)
        return int(np.argmin(d))
    rs = np.random.RandomState(0)
    for a, b in rs.randint(0, 60, (40, 2)):
        assert P[a][b] == idx(R[b] @ R[a])                       # P[a][b] = idx(R_b R_a)
    h = N[0]
    for g, k in rs.randint(0, 60, (40, 2)) % np.array([60, 13]):
        assert N[g][k] == idx(R[h[k]] @ R[g])                    # N[g][k] = idx(R_{h_k} R_g)
    assert np.array_equal(N[:, 0], np.arange(60))
    for k in range(13):
        assert sorted(N[:, k]) == list(range(60))                # every tap is a permutation of the group
    assert len(t.hop1) == 13 and len(t.hop2) == 45 and t.hop2_pos_of_zero() == 0
    a = t.idx_p2_a()
    assert a.min() >= 0 and a.max() < 45 and t.idx_p2_init().shape == (45, 13)


def test_draw_order_matches_reference_loop():
    """yoho_b200.estimator.yohoc draws from the global numpy stream exactly as the reference loop does."""
    from yoho_b200.estimator import yohoc
    class C: ransac_c_inlinerdist = 0.07; SO3_related_files = None
    est = yohoc(C())
    rs = np.random.RandomState(4)
    dr = np.concatenate([np.full(40, 3), np.full(9, 17), np.full(2, 5), rs.randint(20, 60, 30)])
    rs.shuffle(dr)
    stat, prob = est.DR_statictic(dr)
    members, prob_o = O.dr_statistic(dr)
    assert np.array_equal(prob, prob_o) and all(stat[i] == members[i] for i in range(60))
    np.random.seed(9); a = est.draw_hypotheses(stat, prob, 500)
    np.random.seed(9); b = O.draw_yohoc_hypotheses(members, prob_o, 500)
    assert np.array_equal(a, b)
    assert est.DR_statictic(np.arange(60)) == (None, None)
    out = est.estimate(np.zeros((60, 3)), np.zeros((60, 3)), np.arange(60), 10)   # degenerate statistics: no device call
    assert np.array_equal(out["trans"], np.eye(4)) and out["recalltime"] == 50001


def test_registries_and_checkpoint_keys():
    from yoho_b200.network import name2network, PartI_test, PartII_test
    from yoho_b200.extractor import name2extractor
    from yoho_b200.matcher import name2matcher
    from yoho_b200.estimator import name2estimator
    assert set(name2network) == {"PartI_train", "PartI_test", "PartII_train", "PartII_test"}
    assert set(name2extractor) == {"PartI", "PartII"} and set(name2matcher) == {"Match"}
    assert set(name2estimator) == {"yohoc", "yohoc_mul", "yohoo"}
    class C: SO3_related_files = None
    n1, n2 = PartI_test(C()), PartII_test(C())
    assert set(n1.state_dict()) == set(synth.synth_state_dict("PartI", 0))
    assert set(n2.state_dict()) == set(synth.synth_state_dict("PartII", 0))
    sd2 = synth.to_torch_state_dict(synth.synth_state_dict("PartII", 0))
    sd2["PartI_net.PartI_net.Conv_in.0.weight"] = torch.zeros(1)        # nested PartI copy in the real checkpoint
    n2.load_state_dict(sd2, strict=False)
    with pytest.raises(RuntimeError):
        n2.load_state_dict(sd2, strict=True)
    with pytest.raises(NotImplementedError):
        name2network["PartII_train"](C())
    from yoho_b200.train import PartI_train                              # SURVEY.md §8f-4 twin: the reference's keys, strict
    t1 = name2network["PartI_train"](C())
    assert isinstance(t1, PartI_train)
    t1.load_state_dict(synth.to_torch_state_dict(synth.synth_state_dict("PartI", 0)), strict=True)


def test_synth_is_reproducible():
    a = synth.make_fragment_pair(64, seed=3)
    b = synth.make_fragment_pair(64, seed=3)
    assert all(np.array_equal(a[k], b[k]) for k in a if isinstance(a[k], np.ndarray))
    assert np.allclose(np.linalg.norm(a["feat_A"], axis=1), 1, atol=1e-6)
    t = group.load()
    i = a["ids_A"][0]; j = a["ids_B"][0]
    cor = np.einsum("fag,fg->a", a["feat_B"][j][:, t.P.reshape(-1)].reshape(32, 60, 60), a["feat_A"][i])
    assert int(np.argmax(cor)) == a["r"]
    assert np.allclose(a["kps_A"][a["ids_A"]], a["kps_B"][a["ids_B"]] @ a["R_gt"].T + a["t_gt"], atol=0.06)


def test_dropin_aliases():
    import sys
    from yoho_b200 import dropin
    saved = {k: sys.modules.get(k) for k in dropin._ALIASES}
    try:
        names = dropin.install()
        assert "tests.estimator" in names
        import yoho_b200.estimator as ours
        assert sys.modules["tests.estimator"] is ours and sys.modules["utils.network"].name2network
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_make_scene_plants_consistent_pairwise_transforms():
    """yoho_b200.synth.make_scene (scene-level driver input, BASELINE.json configs 3-4): the planted pair transform maps the
    shared points of fragment j onto fragment i, and the group features of shared points are related by ONE group element."""
    import numpy as np
    from yoho_b200 import synth, group
    fr, pairs, gt = synth.make_scene(8, 200, seed=3)
    assert len(fr) == 8 and len(pairs) == 28
    P = group.load().P
    for (i, j) in [pairs[0], pairs[9], pairs[27]]:
        R, t = gt[(i, j)]
        ki, kj = fr[i][1], fr[j][1] @ R.T + t
        d = np.linalg.norm(ki[:, None, :] - kj[None, :, :], axis=2)
        a, b = np.nonzero(d < 0.06)                                   # shared points (1 cm noise on both sides)
        assert len(a) >= 0.25 * 200                                   # overlap rho_i * rho_j >= 0.55^2
        fi, fj = fr[i][0][a], fr[j][0][b]
        cor = np.array([(fi * fj[:, :, P[r]]).sum() for r in range(60)])
        r = int(cor.argmax())
        assert cor[r] > 0.9 * len(a) * 60 and np.sort(cor)[-2] < 0.5 * cor[r]


def test_numa_local_restores_affinity():
    """hostutil.numa_local is a best-effort context manager: whatever NVML answers (or fails to), the affinity is back afterwards."""
    import os
    from yoho_b200.hostutil import numa_local
    before = os.sched_getaffinity(0)
    with numa_local(0):
        assert len(os.sched_getaffinity(0)) >= 1
    assert os.sched_getaffinity(0) == before
    try:
        with numa_local(0):
            raise KeyError("x")
    except KeyError:
        pass
    assert os.sched_getaffinity(0) == before


def test_pair_output_layout_is_aligned_and_disjoint():
    """Engine._pair_layout (the single allocation behind yoho_register_pair's outputs): 256-byte aligned, non-overlapping
    entries, T_c / T_o back to back in one [2,3,4] block, PartI outputs only when they are not supplied."""
    import types
    import torch
    from yoho_b200.engine import Engine
    me = types.SimpleNamespace(_ESIZE=Engine._ESIZE)
    for Ka, Kb, iters, have in [(5000, 4000, 1000, False), (300, 300, 0, True), (0, 7, 10, False)]:
        ent, total = Engine._pair_layout(me, Ka, Kb, iters, have)
        spans = sorted((o, o + n) for (o, n, _, _) in ent.values())
        assert all(o % 256 == 0 for o, _ in spans) and spans[-1][1] <= total
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
        assert ent["T_co"][2] == (2, 3, 4) and ent["T_co"][3] == torch.float64 and ent["T_co"][1] == 192
        cap = max(1, min(Ka, Kb))
        assert ent["pairs"][2] == (cap, 2) and ent["trans"][2] == (cap, 3, 4)
        assert ("eqvA" in ent) == (not have)
    assert Engine._pair_layout(me, 300, 300, 0, True) is Engine._pair_layout(me, 300, 300, 0, True)      # cached


def test_lapack_replay_reproduces_the_reference_hypotheses():
    """Host half of the reference-identical YOHO-C run (yoho_b200/estimator.py): the batched LAPACK call gives the
    per-hypothesis transforms the unmodified reference computed one by one (goldens recorded by subclassing its
    Threepps2Tran), bit for bit on the authoring host; rank-deficient triplets are flagged for pass-through."""
    import importlib
    est = importlib.import_module("yoho_b200.estimator")
    from yoho_b200 import synth
    from conftest import load_golden
    for name in ("pipeline_synth.npz", "pipeline_realckpt.npz"):
        g = load_golden(name)
        pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
        m = g["matches"]
        k0, k1 = pair["kps_A"][m[:, 0]], pair["kps_B"][m[:, 1]]
        trans, signs = est.yohoc.lapack_replay(k0, k1, g["c_hyp"])
        ok = signs != 2
        assert np.abs(trans[ok] - g["c_hyp_trans"][ok]).max() <= 1e-9
        h = g["c_hyp"]
        dup = (h[:, 0] == h[:, 1]) | (h[:, 0] == h[:, 2]) | (h[:, 1] == h[:, 2])
        assert np.array_equal(~ok, dup)
        assert np.array_equal(signs[ok], g["c_sign"][ok])
