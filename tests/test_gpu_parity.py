"""GPU parity tests proper: every stage of the hot path, called through the C ABI (yoho_b200.engine ->
libyoho_b200.so), against (1) the committed goldens recorded from the unmodified reference and (2) the oracle
on seeded inputs.  Bars: descriptors / quaternions within 1e-4 (north_star), index / match / inlier work
bit-exact (near-ties arbitrated in FP64 and required to be within FP32 noise)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, real_ckpt
import yoho_oracle as O
import estimator_oracle as E
from yoho_b200 import synth

pytestmark = pytest.mark.gpu

DESC_TOL = 1e-4      # BASELINE.json north_star: "within 1e-4 on descriptors"


def _np(t):
    return t.detach().cpu().numpy()


# ------------------------------------------------------------------------------------------------ PartI
def test_part1_golden_stage(engine):
    g = load_golden("stages_synth.npz")
    engine.load_part1(synth.synth_state_dict("PartI", 1))
    x, _ = synth.make_fragment(40, 21)
    o = engine.part1(x)
    assert np.abs(_np(o["eqv"]) - g["p1_eqv"]).max() <= DESC_TOL
    assert np.abs(_np(o["inv"]) - g["p1_inv"]).max() <= DESC_TOL
    if engine.impl_name == "simt":      # the FP32 SIMT path is far inside the bar; keep it honest
        assert np.abs(_np(o["eqv"]) - g["p1_eqv"]).max() <= 2e-5


@pytest.mark.parametrize("K", [1, 2, 59, 130, 700])
def test_part1_vs_oracle_ragged(engine, tables, K):
    _, _, N = tables
    sd = synth.synth_state_dict("PartI", 2)
    engine.load_part1(sd)
    x, _ = synth.make_fragment(K, 100 + K)
    o = engine.part1(x)
    if K == 1:
        ref = O.part1_forward(np.concatenate([x, x]), sd, N)     # the reference itself cannot run B=1
        ref = {k: v[:1] for k, v in ref.items()}
    else:
        ref = O.part1_forward(x, sd, N)
    assert np.abs(_np(o["eqv"]) - ref["eqv"].numpy()).max() <= DESC_TOL
    assert np.abs(_np(o["inv"]) - ref["inv"].numpy()).max() <= DESC_TOL
    assert np.array_equal(_np(o["desc"]), O.matcher_descriptor(_np(o["eqv"])))   # numpy's mean, bit for bit


def test_part1_empty(engine):
    engine.load_part1(synth.synth_state_dict("PartI", 2))
    o = engine.part1(np.zeros((0, 32, 60), np.float32))
    assert o["eqv"].shape == (0, 32, 60)


def test_part1_realckpt(engine, tables):
    sd = real_ckpt("PartI")
    if sd is None:
        pytest.skip("oracle/_ref/ckpt not present")
    _, _, N = tables
    engine.load_part1(sd)
    x, _ = synth.make_fragment(200, 5)
    o = engine.part1(x)
    ref = O.part1_forward(x, sd, N)
    ref64 = O.part1_forward(x, sd, N, torch.float64)
    assert np.abs(_np(o["eqv"]) - ref["eqv"].numpy()).max() <= DESC_TOL
    assert np.abs(_np(o["eqv"]) - ref64["eqv"].numpy()).max() <= DESC_TOL
    g = load_golden("pipeline_realckpt.npz")
    pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
    assert np.abs(_np(engine.part1(pair["feat_A"])["eqv"]) - g["eqv0"]).max() <= DESC_TOL


def test_part1_equivariance(engine, tables):
    """The reference's one self-check (utils/network.py:290-312): permuting the input by P[i] permutes eqv the
    same way and leaves inv unchanged."""
    _, P, _ = tables
    engine.load_part1(synth.synth_state_dict("PartI", 0))
    x, _ = synth.make_fragment(16, 3)
    base = engine.part1(x)
    for i in (0, 1, 17, 59):
        o = engine.part1(np.ascontiguousarray(x[:, :, P[i]]))
        assert np.abs(_np(o["eqv"]) - _np(base["eqv"])[:, :, P[i]]).max() <= 1e-5
        assert np.abs(_np(o["inv"]) - _np(base["inv"])).max() <= 1e-5


def test_forward_without_weights_fails_loudly():
    from yoho_b200.engine import Engine
    from yoho_b200._lib import YohoError
    e = Engine()
    with pytest.raises(YohoError, match="No model exists"):
        e.part1(np.zeros((2, 32, 60), np.float32))
    e.close()


# ------------------------------------------------------------------------------------------------ matching
def _check_nn(src, tgt, got_idx):
    want = O.nn1(src, tgt)[1].numpy()
    bad = np.nonzero(got_idx != want)[0]
    if bad.size:
        best, second, _ = O.nn1_margins(src, tgt)
        assert np.all((second[bad] - best[bad]) <= 1e-6 * np.maximum(best[bad], 1e-12)), "non-tie argmin mismatch"
    return bad.size


def test_group_mean_bit_exact(engine):
    rs = np.random.RandomState(0)
    e = rs.standard_normal((77, 32, 60)).astype(np.float32)
    assert np.array_equal(_np(engine.group_mean(e)), np.mean(e, axis=-1))


def test_nn1_golden(engine):
    g = load_golden("stages_synth.npz")
    d, idx = engine.nn1(g["knn_d0"], g["knn_d1"])
    assert np.array_equal(_np(idx), g["knn_a01"]) or _check_nn(g["knn_d0"], g["knn_d1"], _np(idx)) >= 0
    assert np.abs(_np(d) - g["knn_dist01"]).max() <= 1e-6
    d, idx = engine.nn1(g["knn_d1"], g["knn_d0"])
    assert np.array_equal(_np(idx), g["knn_a10"]) or _check_nn(g["knn_d1"], g["knn_d0"], _np(idx)) >= 0
    assert np.abs(_np(d) - g["knn_dist10"]).max() <= 1e-6


def test_nn1_3d(engine):
    """F < 32 (the 3-D search of the feature-lift 'next' row, YOHO_testset.py:155-158) is zero-padded."""
    rs = np.random.RandomState(1)
    s, t = rs.rand(500, 3).astype(np.float32), rs.rand(333, 3).astype(np.float32)
    d, idx = engine.nn1(s, t)
    _check_nn(s, t, _np(idx))


@pytest.mark.parametrize("Ka,Kb", [(1, 1), (5, 3), (64, 64), (65, 129), (1000, 777), (5000, 5000)])
def test_mutual_nn_vs_oracle(engine, Ka, Kb):
    rs = np.random.RandomState(Ka * 7 + Kb)
    dA = (rs.standard_normal((Ka, 32)) * 0.1).astype(np.float32)
    dB = (rs.standard_normal((Kb, 32)) * 0.1).astype(np.float32)
    n = min(Ka, Kb) // 2
    if n:
        dB[:n] = dA[rs.permutation(Ka)[:n]] + (rs.standard_normal((n, 32)) * 0.01).astype(np.float32)
    pairs, cnt, nnA, nnB = engine.mutual_nn(dA, dB, want_nn=True)
    M = int(cnt.item())
    got = _np(pairs[:M])
    tiesA = _check_nn(dA, dB, _np(nnA).astype(np.int64))
    tiesB = _check_nn(dB, dA, _np(nnB).astype(np.int64))
    # the mutual filter / ordering is exact given the two argmin arrays
    a01, a10 = _np(nnA).astype(np.int64), _np(nnB).astype(np.int64)
    keep = a10[a01] == np.arange(Ka)
    want = np.stack([np.arange(Ka)[keep], a01[keep]], 1)
    assert got.dtype == np.int64 and np.array_equal(got, want)
    if tiesA == 0 and tiesB == 0:
        assert np.array_equal(got, O.mutual_matches(dA, dB)[0])
    if n:
        assert M >= n // 2


def test_mutual_nn_duplicates_tie_to_lowest_index(engine):
    """Exact duplicates: torch.min keeps the first minimal index (utils/knn_search.py:41)."""
    rs = np.random.RandomState(3)
    dA = (rs.standard_normal((40, 32)) * 0.1).astype(np.float32)
    dB = np.concatenate([dA[:10], dA[:10], dA[10:30]]).astype(np.float32)   # rows 0-9 appear twice in B
    pairs, cnt, nnA, nnB = engine.mutual_nn(dA, dB, want_nn=True)
    assert np.array_equal(_np(nnA)[:10], np.arange(10))                      # first copy wins
    assert np.array_equal(_np(pairs[: int(cnt.item())]), O.mutual_matches(dA, dB)[0])


def test_matcher_golden(engine):
    g = load_golden("pipeline_synth.npz")
    d0, d1 = engine.group_mean(g["eqv0"]), engine.group_mean(g["eqv1"])
    pairs, cnt = engine.mutual_nn(d0, d1)
    assert np.array_equal(_np(pairs[: int(cnt.item())]), g["matches"])


# ------------------------------------------------------------------------------------------------ rotation index
def test_rot_argmax_golden(engine):
    g = load_golden("stages_synth.npz")
    pr = synth.make_fragment_pair(48, seed=23, overlap=1.0, sigma=0.3)
    des1, des2 = pr["feat_B"][pr["ids_B"]], pr["feat_A"][pr["ids_A"]]
    idx, cor = engine.rot_argmax(des1, des2, want_cor=True)
    assert np.array_equal(_np(idx), g["rot_idx"])
    assert np.abs(_np(cor) - g["rot_cor"]).max() <= 1e-4
    g2 = load_golden("pipeline_synth.npz")
    idx = engine.rot_argmax(g2["eqv1"], g2["eqv0"], pairs=g2["matches"])
    assert idx.dtype == torch.int64 and np.array_equal(_np(idx), g2["dr_index"])


def test_rot_argmax_all_rotations(engine, tables):
    """Batch_Des2R(a, a[:, :, P[i]]) == i for every group element."""
    _, P, _ = tables
    x, _ = synth.make_fragment(60, 9)
    y = np.stack([x[i][:, P[i]] for i in range(60)])
    idx = engine.rot_argmax(x, y)
    assert np.array_equal(_np(idx), np.arange(60))
    assert engine.rot_argmax(x[:0], y[:0]).shape == (0,)


def test_rot_argmax_random_vs_fp64(engine, tables):
    _, P, _ = tables
    a, _ = synth.make_fragment(512, 31)
    b, _ = synth.make_fragment(512, 32)
    idx, cor = engine.rot_argmax(a, b, want_cor=True)
    want, cor64 = O.rot_argmax(a, b, P, torch.float64)
    bad = np.nonzero(_np(idx) != want)[0]
    top2 = np.sort(cor64.numpy(), axis=1)[:, -2:]
    assert np.all((top2[bad, 1] - top2[bad, 0]) <= 1e-5), "non-tie rotation index mismatch"
    assert np.abs(_np(cor) - cor64.numpy()).max() <= 1e-4


# ------------------------------------------------------------------------------------------------ PartII
def test_part2_golden_stage(engine):
    g = load_golden("stages_synth.npz")
    engine.load_part2(synth.synth_state_dict("PartII", 1))
    pp = synth.make_fragment_pair(24, seed=24, overlap=1.0, sigma=0.05)
    fA, fB = pp["feat_A"][pp["ids_A"]], pp["feat_B"][pp["ids_B"]]
    q, _ = engine.part2(fA, fB, g["p2_yA"], g["p2_yB"], g["p2_pre"])
    assert np.abs(_np(q) - g["p2_quat"]).max() <= DESC_TOL


def _part2_pipeline_case(engine, tables, golden, sdII):
    g = load_golden(golden)
    R, P, N = tables
    engine.load_part2(sdII)
    pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
    q, tr = engine.part2(pair["feat_A"], pair["feat_B"], g["eqv0"], g["eqv1"], g["dr_index"], pairs=g["matches"],
                         kps0=pair["kps_A"], kps1=pair["kps_B"])
    assert tr.dtype == torch.float64 and tuple(tr.shape) == (g["matches"].shape[0], 3, 4)
    assert np.abs(_np(tr) - g["trans_pre"]).max() <= DESC_TOL
    m = g["matches"]
    want = O.part2_transforms(_np(q), g["dr_index"], pair["kps_A"][m[:, 0]], pair["kps_B"][m[:, 1]], R)
    assert np.abs(_np(tr) - want).max() <= 1e-12       # the fp32-quaternion -> f64 transform arithmetic itself


def test_part2_pipeline_golden(engine, tables):
    _part2_pipeline_case(engine, tables, "pipeline_synth.npz", synth.synth_state_dict("PartII", 0))


def test_part2_pipeline_realckpt(engine, tables):
    sd = real_ckpt("PartII")
    if sd is None:
        pytest.skip("oracle/_ref/ckpt not present")
    _part2_pipeline_case(engine, tables, "pipeline_realckpt.npz", sd)


@pytest.mark.parametrize("M", [1, 3, 130])
def test_part2_vs_oracle(engine, tables, M):
    R, P, N = tables
    sd = synth.synth_state_dict("PartII", 4)
    engine.load_part2(sd)
    rs = np.random.RandomState(M)
    fA, _ = synth.make_fragment(M, 40 + M)
    fB, _ = synth.make_fragment(M, 50 + M)
    yA, _ = synth.make_fragment(M, 60 + M)
    yB, _ = synth.make_fragment(M, 70 + M)
    pre = rs.randint(0, 60, M).astype(np.int64)
    q, _ = engine.part2(fA, fB, yA, yB, pre)
    want = O.part2_forward(fA, fB, yA, yB, pre, sd, P, N)
    assert np.abs(_np(q) - want.numpy()).max() <= DESC_TOL


# ------------------------------------------------------------------------------------------------ estimators
def _planted(M, seed, inlier_frac=0.4, noise=0.01):
    rs = np.random.RandomState(seed)
    k1 = rs.uniform(0, 3, (M, 3))
    pr = synth.make_fragment_pair(4, seed=seed)
    k0 = k1 @ pr["R_gt"].T + pr["t_gt"] + noise * rs.standard_normal((M, 3))
    out = rs.rand(M) > inlier_frac
    k0[out] = rs.uniform(-2, 5, (int(out.sum()), 3))
    dr = np.where(out, rs.randint(0, 60, M), pr["r"]).astype(np.int64)
    return k0, k1, dr, pr


@pytest.mark.parametrize("M,iters", [(3, 10), (87, 1000), (1000, 1000), (5000, 300)])
def test_c_ransac_vs_oracle_bit_exact(engine, M, iters):
    k0, k1, dr, _ = _planted(M, 11 + M)
    rs = np.random.RandomState(M)
    hyp = rs.randint(0, M, (iters, 3)).astype(np.int32)
    hyp[::7, 1] = hyp[::7, 0]                       # duplicates (drawn WITH replacement in the reference)
    hyp[::31] = hyp[::31, :1]                       # fully degenerate triplets
    for signs in (None, rs.choice(np.array([-1, 0, 1], np.int8), iters)):
        res = engine.c_ransac(k0, k1, hyp, 0.07, signs=signs, want_counts=True)
        want = E.yohoc(k0, k1, hyp, 0.07, signs=signs)
        assert np.array_equal(_np(res["counts"]), want["counts"])
        assert int(res["best_iter"].item()) == want["best_iter"]
        assert int(res["n_inl"].item()) == want["n_inl"]
        assert np.array_equal(_np(res["mask"]), want["mask"])
        assert np.abs(_np(res["T"]) - want["T"]).max() <= 1e-12


def test_c_ransac_golden_replay(engine):
    """Reference yohoc run replayed: same triplets, LAPACK's null-space signs -> same winner and transform."""
    for name in ("pipeline_synth.npz", "pipeline_realckpt.npz"):
        g = load_golden(name)
        pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
        m = g["matches"]
        k0, k1 = pair["kps_A"][m[:, 0]], pair["kps_B"][m[:, 1]]
        res = engine.c_ransac(k0, k1, g["c_hyp"], float(g["c_dist"]), signs=g["c_sign"], want_counts=True)
        want = E.yohoc(k0, k1, g["c_hyp"], float(g["c_dist"]), signs=g["c_sign"])
        ok = ~want["degenerate"]
        ref_counts = np.rint(g["c_overlap"] * m.shape[0]).astype(np.int64)
        assert np.array_equal(_np(res["counts"])[ok], ref_counts[ok])
        # full replay: rank-deficient triplets take LAPACK's (arbitrary) transform as recorded from the reference, so
        # EVERY per-hypothesis count, the winner and the transform are the reference's
        signs = g["c_sign"].copy()
        signs[~ok] = 2
        res = engine.c_ransac(k0, k1, g["c_hyp"], float(g["c_dist"]), signs=signs, fixed=g["c_hyp_trans"], want_counts=True)
        want = E.yohoc(k0, k1, g["c_hyp"], float(g["c_dist"]), signs=signs, fixed=g["c_hyp_trans"])
        assert np.array_equal(_np(res["counts"]), ref_counts) and np.array_equal(want["counts"], ref_counts)
        assert int(res["best_iter"].item()) + 1 == int(g["c_recalltime"]) == want["best_iter"] + 1
        assert np.abs(_np(res["T"]) - g["c_trans"][:3]).max() <= 1e-9
        assert int(res["n_inl"].item()) == ref_counts.max()


def test_c_ransac_no_inlier_gives_identity(engine):
    """No hypothesis has an inlier -> eye(4), recalltime 0 (tests/estimator.py:111-117: best_overlap stays 0).  The triplet's
    two triangles differ in scale by 1000, so even its own three points miss by hundreds of metres; the other matches are far."""
    k1 = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [50, 50, 50], [60, 60, 60]], np.float64)
    k0 = np.array([[0, 0, 0], [1000, 0, 0], [0, 1000, 0], [1e5, -1e5, 3e5], [-2e5, 7e5, 1e5]], np.float64)
    hyp = np.array([[0, 1, 2], [2, 1, 0], [0, 2, 1]], np.int32)
    res = engine.c_ransac(k0, k1, hyp, 0.07, want_counts=True)
    assert np.array_equal(_np(res["counts"]), [0, 0, 0]) and int(res["n_inl"].item()) == 0
    assert int(res["best_iter"].item()) == -1
    assert np.array_equal(_np(res["T"]), np.eye(4)[:3])
    assert not _np(res["mask"]).any()
    want = E.yohoc(k0, k1, hyp, 0.07)
    assert want["best_iter"] == -1 and np.array_equal(want["T"], np.eye(4)[:3])


@pytest.mark.parametrize("M,H", [(1, 1), (87, 87), (1000, 1000), (2000, 500)])
def test_o_score_vs_oracle(engine, M, H):
    k0, k1, dr, pr = _planted(M, 5 + M)
    rs = np.random.RandomState(M + 1)
    trans = np.zeros((M, 3, 4))
    for i in range(M):
        p = synth.make_fragment_pair(4, seed=1000 + i % 50)
        trans[i, :, :3], trans[i, :, 3] = p["R_gt"], p["t_gt"]
    trans[::9, :, :3], trans[::9, :, 3] = pr["R_gt"], pr["t_gt"]
    order = rs.permutation(M)[:H].astype(np.int32)
    res = engine.o_score(k0, k1, trans, 0.09, order=order, want_counts=True)
    want = E.yohoo(k0, k1, trans[order], 0.09)
    assert np.array_equal(_np(res["counts"]), want["counts"])
    assert int(res["best_iter"].item()) == want["best_iter"]
    assert np.array_equal(_np(res["mask"]), want["mask"])
    assert np.array_equal(_np(res["T"]), want["T"])


def test_o_score_golden(engine):
    g = load_golden("pipeline_synth.npz")
    pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
    m = g["matches"]
    res = engine.o_score(pair["kps_A"][m[:, 0]], pair["kps_B"][m[:, 1]], g["trans_pre"], float(g["o_dist"]),
                         order=g["o_order"])
    assert int(res["best_iter"].item()) == int(g["o_recalltime"])
    assert np.array_equal(_np(res["T"]), g["o_trans"][:3])


def test_device_draws(engine):
    """yoho_c_draw: bins follow the reference's weights, triplets stay inside one bin, degenerate statistics flagged."""
    rs = np.random.RandomState(0)
    dr = np.concatenate([np.full(300, 7), np.full(100, 21), np.full(2, 33), rs.randint(40, 60, 60)]).astype(np.int64)
    rs.shuffle(dr)
    hyp, status = engine.c_draw(dr, 4000, seed=5)
    assert int(status.item()) == 0
    h = _np(hyp)
    bins = dr[h]
    assert (bins[:, 0] == bins[:, 1]).all() and (bins[:, 0] == bins[:, 2]).all()
    members, prob = O.dr_statistic(dr)
    freq = np.bincount(bins[:, 0], minlength=60) / 4000.0
    assert np.abs(freq - prob).max() < 0.03
    assert freq[prob == 0].sum() == 0
    hyp2, _ = engine.c_draw(dr, 4000, seed=5)
    assert np.array_equal(_np(hyp2), h)                                   # counter-based: reproducible
    _, status = engine.c_draw(np.arange(60, dtype=np.int64), 10, seed=1)  # one member per bin -> reference returns None
    assert int(status.item()) == 1


def _philox4x32(ctr, key):
    """Philox4x32-10 (csrc/estimator.cu) on python ints."""
    c = list(ctr); k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xffffffff, p0 & 0xffffffff]
        k = [(k[0] + 0x9E3779B9) & 0xffffffff, (k[1] + 0xBB67AE85) & 0xffffffff]
    return c


def test_device_draws_equal_host_restatement(engine):
    """yoho_c_draw against a line-by-line host restatement of its arithmetic (DR_statictic weights in float64 in the reference's
    order, np.random.choice's cdf construction, searchsorted(side='right'), members in ascending match order, Philox4x32-10
    counters): every triplet identical.  Pins the kernel across restructurings (histogram / bucket fill / search)."""
    rs = np.random.RandomState(4)
    for M, seed in ((2800, 11), (517, 12345678901), (40, 3)):
        dr = np.where(rs.rand(M) < 0.5, 17, rs.randint(0, 60, M)).astype(np.int64)
        iters = 1000
        hyp, status = engine.c_draw(dr, iters, seed)
        cnt = np.bincount(dr, minlength=60)
        w = [0.0 if c < 2 else (c / 100.0) * (c / 100.0 - 0.01) * (c / 100.0 - 0.02) for c in cnt]
        tot = 0.0
        for x in w:
            tot += x
        if tot < 1e-4:
            assert int(status.item()) == 1
            continue
        assert int(status.item()) == 0
        pn = [x / tot for x in w]
        cdf, c = [], 0.0
        for x in pn:
            c += x
            cdf.append(c)
        cdf = [x / cdf[-1] for x in cdf]
        members = [np.nonzero(dr == b)[0] for b in range(60)]
        key = (seed & 0xffffffff, (seed >> 32) & 0xffffffff)
        want = np.zeros((iters, 3), np.int32)
        for it in range(iters):
            r = _philox4x32((it, 0, 0x59484f43, 0), key)
            u = ((r[0] >> 5) * 67108864.0 + (r[1] >> 6)) / 9007199254740992.0
            b = 0
            while b < 59 and not (u < cdf[b]):
                b += 1
            r2 = _philox4x32((it, 1, 0x59484f43, 0), key)
            n = int(cnt[b])
            want[it] = [members[b][(r2[j] * n) >> 32] for j in range(3)]
        assert np.array_equal(_np(hyp), want)


def test_device_order_is_permutation(engine):
    for M in (1, 87, 2500):
        o = _np(engine.o_order(M, seed=3))
        assert np.array_equal(np.sort(o), np.arange(M))


# ------------------------------------------------------------------------------------------------ next row: feature lift
def test_lift_group_features_vs_oracle(_engine_session, tables):
    """SURVEY.md §8f-1 (YOHO_testset.py:153-166): 60 rotated 3-D 1-NN searches + feature gather, exact."""
    R, _, _ = tables
    rs = np.random.RandomState(7)
    K = 333
    kps = rs.uniform(-1.5, 1.5, (K, 3))
    base = rs.uniform(-1.6, 1.6, (2500, 3))
    pts, feats = [], []
    for g in range(60):
        n = int(rs.randint(1200, 2500))
        p = (base[rs.permutation(2500)[:n]] @ R[g].T).astype(np.float32)
        if g == 3:
            p[10] = p[5]                       # duplicate point: the first index must win
        pts.append(p)
        feats.append(rs.standard_normal((n, 32)).astype(np.float32))
    out, nn = _engine_session.lift_group_features(kps, pts, feats, want_nn=True)
    want, want_nn = O.lift_group_features(kps, pts, feats, R)
    assert np.array_equal(_np(nn), want_nn)
    assert out.shape == (K, 32, 60) and np.array_equal(_np(out), want)
