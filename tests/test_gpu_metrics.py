"""GPU: the device evaluation metrics (csrc/metrics.cu through the C ABI and yoho_b200/rr_cal.py) against the
oracle and the golden vectors recorded from the reference.  Integer results (match counts, flags, precision / recall) are
exact; float64 errors agree within 1e-9 relative (Jacobi eigenvectors / Gauss-Jordan inverse instead of LAPACK)."""
import os
import sys
import numpy as np
import pytest
import torch

from conftest import load_golden
import metrics_oracle as MO
from yoho_b200 import rr_cal

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

RTOL = 1e-9


@pytest.fixture(scope="module")
def eng(_engine_session):
    return _engine_session


@pytest.mark.parametrize("tag", ["a", "b"])
def test_evaluate_registration_golden(eng, tag):
    g = load_golden("metrics_synth.npz")
    prec, rec, flags, errors = rr_cal.evaluate_registration(int(g[f"{tag}_n_frag"]), g[f"{tag}_est"], g[f"{tag}_est_pairs"], g[f"{tag}_gt_pairs"],
                                                            g[f"{tag}_gt"], g[f"{tag}_info"], err2=0.2,
                                                            nonconsecutive=bool(g[f"{tag}_nonconsecutive"]), engine=eng)
    assert np.array_equal(np.array(flags), g[f"{tag}_flags"])
    assert prec == float(g[f"{tag}_precision"]) and rec == float(g[f"{tag}_recall"])
    assert np.allclose(np.array(errors), g[f"{tag}_errors"], rtol=RTOL, atol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_errors_golden(eng, tag):
    g = load_golden("metrics_synth.npz")
    est, ext, info = g[f"{tag}_est"], g[f"{tag}_ext_gt"], g[f"{tag}_info"]
    inf = info[np.minimum(np.arange(len(est)), len(info) - 1)]
    p, re, te = eng.registration_errors(est, ext, inf)
    assert np.allclose(p.cpu().numpy(), g[f"{tag}_p_all"], rtol=RTOL, atol=1e-12)
    assert np.allclose(re.cpu().numpy(), g[f"{tag}_re"][:, 0], rtol=RTOL, atol=1e-7)     # acos near 1 amplifies last-bit differences
    assert np.allclose(te.cpu().numpy(), g[f"{tag}_te"], rtol=1e-13, atol=0)
    # the reference-named wrappers
    re2 = rr_cal.rotation_error(torch.from_numpy(ext[:, :3, :3]), torch.from_numpy(est[:, :3, :3]), engine=eng)
    te2 = rr_cal.translation_error(torch.from_numpy(ext[:, :3, 3:4]), torch.from_numpy(est[:, :3, 3:4]), engine=eng)
    assert re2.shape == (len(est), 1) and torch.equal(re2[:, 0], re) and torch.equal(te2, te)
    k = 3
    one = rr_cal.computeTransformationErr(np.linalg.inv(ext[k]) @ est[k], inf[k], engine=eng)
    assert abs(one - g[f"{tag}_p_all"][k]) <= RTOL * abs(g[f"{tag}_p_all"][k]) + 1e-12


def test_fmr_golden_and_batch(eng):
    g = load_golden("metrics_synth.npz")
    thr = float(g["fmr_threshold"])
    k0, k1, gts = [], [], []
    for k in range(6):
        m = g[f"fmr{k}_matches"]
        k0.append(g[f"fmr{k}_kps0"][m[:, 0]])
        k1.append(g[f"fmr{k}_kps1"][m[:, 1]])
        gts.append(g[f"fmr{k}_gt"])                      # 4x4 and 3x4 alternate
    ratios = rr_cal.pair_match_ratios(k0, k1, gts, thr, engine=eng)
    assert np.array_equal(ratios, g["fmr_ratios"])
    fmr, pr = rr_cal.feature_match_recall(k0, k1, gts, thr, 0.2, engine=eng)
    assert fmr == float(np.mean(g["fmr_ratios"] > 0.2)) and np.array_equal(pr, ratios)
    one = rr_cal.evaluate_the_match(g["fmr2_kps0"], g["fmr2_kps1"], g["fmr2_matches"], g["fmr2_gt"], thr, engine=eng)
    assert one == float(g["fmr_ratios"][2])


def test_random_scenes_against_oracle(eng):
    import make_golden_metrics as G
    for seed, noncons in ((201, True), (202, False), (203, True), (204, True)):
        n_frag, est, est_pairs, gt_pairs, gt, info = G.make_scene(seed, n_frag=20)
        want = MO.evaluate_registration(n_frag, est, est_pairs, gt_pairs, gt, info, err2=0.2, nonconsecutive=noncons)
        got = rr_cal.evaluate_registration(n_frag, est, est_pairs, gt_pairs, gt, info, err2=0.2, nonconsecutive=noncons, engine=eng)
        assert got[0] == want[0] and got[1] == want[1] and list(got[2]) == list(want[2])
        assert np.allclose(np.array(got[3]), np.array(want[3]), rtol=RTOL, atol=1e-12)
    # rotations by ~0 and ~180 degrees, identical transforms, large translations
    rs = np.random.RandomState(3)
    gt = np.stack([G.rand_rigid(rs, angle_deg=a, trans=t) for a, t in ((0, 0), (180, 1), (179.999, 50), (1e-6, 1e-3), (90, 1))])
    est = np.stack([gt[0], gt[1] @ G.rand_rigid(rs, angle_deg=180, trans=0), gt[2], G.rand_rigid(rs, angle_deg=180, trans=1), gt[4]])
    info = np.stack([G.rand_info(rs, 50) for _ in range(5)])
    p, re, te = eng.registration_errors(est, gt, info)
    pw, rew, tew = MO.registration_errors(est, gt, info)
    assert np.allclose(p.cpu().numpy(), pw, rtol=1e-8, atol=1e-12)
    assert np.allclose(re.cpu().numpy(), rew, rtol=1e-9, atol=2e-6)
    assert np.allclose(te.cpu().numpy(), tew, rtol=1e-13, atol=1e-15)


def test_large_batch_and_empty(eng):
    rs = np.random.RandomState(8)
    n = 1623                                              # pairs of the 3DMatch test split (BASELINE.json configs[2])
    sizes = rs.randint(0, 900, n)
    sizes[:3] = 0                                         # pairs without matches
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    k1 = rs.uniform(0, 3, (int(off[-1]), 3))
    import make_golden_metrics as G
    gts = np.stack([G.rand_rigid(rs) for _ in range(n)])
    k0 = np.empty_like(k1)
    for p in range(n):
        s = slice(off[p], off[p + 1])
        k0[s] = k1[s] @ gts[p, :3, :3].T + gts[p, :3, 3] + rs.standard_normal((sizes[p], 3)) * 0.07
    counts = eng.fmr_counts(k0, k1, off, gts, 0.1).cpu().numpy()
    want = np.array([MO.match_ok_count(k0[off[p]:off[p + 1]], k1[off[p]:off[p + 1]], gts[p], 0.1) for p in range(n)])
    assert np.array_equal(counts, want)
    ratios = rr_cal.pair_match_ratios([k0[off[p]:off[p + 1]] for p in range(8)], [k1[off[p]:off[p + 1]] for p in range(8)], list(gts[:8]), 0.1, engine=eng)
    assert np.isnan(ratios[:3]).all() and np.array_equal(ratios[3:], want[3:8] / sizes[3:8])
    assert rr_cal.pair_match_ratios([], [], [], 0.1, engine=eng).shape == (0,)
    p, re, te = eng.registration_errors(np.zeros((0, 4, 4)), np.zeros((0, 4, 4)), np.zeros((0, 6, 6)))
    assert p.numel() == 0 and re.numel() == 0 and te.numel() == 0
    prec, rec, flags, errors = rr_cal.evaluate_registration(4, np.zeros((0, 4, 4)), np.zeros((0, 3)), np.array([[0, 2, 4], [1, 3, 4]]),
                                                            np.tile(np.eye(4), (2, 1, 1)), np.tile(np.eye(6), (2, 1, 1)), engine=eng)
    assert prec == 0.0 and rec == 0.0 and flags == [] and errors == []
