"""CPU: the metrics oracle (oracle/metrics_oracle.py) against the golden vectors recorded from the reference's
own functions (tests/golden/make_golden_metrics.py), plus the known-answer check of the restated nibabel `mat2quat`."""
import numpy as np
import pytest

from conftest import load_golden
import metrics_oracle as MO

TOL = 1e-10          # float64 arithmetic replayed on another host (BLAS / libm last-bit differences)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_evaluate_registration_golden(tag):
    g = load_golden("metrics_synth.npz")
    prec, rec, flags, errors = MO.evaluate_registration(int(g[f"{tag}_n_frag"]), g[f"{tag}_est"], g[f"{tag}_est_pairs"], g[f"{tag}_gt_pairs"],
                                                        g[f"{tag}_gt"], g[f"{tag}_info"], err2=0.2, nonconsecutive=bool(g[f"{tag}_nonconsecutive"]))
    assert np.array_equal(np.array(flags), g[f"{tag}_flags"])
    assert prec == float(g[f"{tag}_precision"]) and rec == float(g[f"{tag}_recall"])
    assert np.allclose(np.array(errors), g[f"{tag}_errors"], rtol=TOL, atol=TOL)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_errors_golden(tag):
    g = load_golden("metrics_synth.npz")
    est, ext, info = g[f"{tag}_est"], g[f"{tag}_ext_gt"], g[f"{tag}_info"]
    re = MO.rotation_error(ext[:, :3, :3], est[:, :3, :3])
    te = MO.translation_error(ext[:, :3, 3:4], est[:, :3, 3:4])
    assert re.shape == g[f"{tag}_re"].shape
    assert np.allclose(re, g[f"{tag}_re"], rtol=TOL, atol=1e-9) and np.allclose(te, g[f"{tag}_te"], rtol=TOL, atol=TOL)
    inf = info[np.minimum(np.arange(len(est)), len(info) - 1)]
    p, re2, te2 = MO.registration_errors(est, ext, inf)
    assert np.allclose(p, g[f"{tag}_p_all"], rtol=1e-9, atol=1e-12)
    assert np.allclose(re2, g[f"{tag}_re"][:, 0], rtol=TOL, atol=1e-9) and np.allclose(te2, g[f"{tag}_te"], rtol=TOL, atol=TOL)


def test_fmr_golden():
    g = load_golden("metrics_synth.npz")
    thr = float(g["fmr_threshold"])
    for k in range(6):
        m = g[f"fmr{k}_matches"]
        r = MO.pair_fmr(g[f"fmr{k}_kps0"][m[:, 0]], g[f"fmr{k}_kps1"][m[:, 1]], g[f"fmr{k}_gt"], thr)
        assert r == float(g["fmr_ratios"][k])
    fmr, ratios = MO.scene_fmr(g["fmr_ratios"], 0.2)
    assert fmr == np.mean(g["fmr_ratios"] > 0.2)


def test_mat2quat_known_answers():
    """nibabel's mat2quat of an exact rotation matrix is the unit quaternion (w>=0) of that rotation."""
    from scipy.spatial.transform import Rotation
    rs = np.random.RandomState(0)
    for _ in range(200):
        rot = Rotation.from_rotvec(rs.standard_normal(3) * rs.uniform(0, 1.0))
        q = MO.mat2quat(rot.as_matrix())
        x, y, z, w = rot.as_quat()
        want = np.array([w, x, y, z]) * (1 if w >= 0 else -1)
        assert np.allclose(q, want, atol=1e-12)
    assert np.allclose(MO.mat2quat(np.eye(3)), [1, 0, 0, 0])
    assert np.allclose(np.abs(MO.mat2quat(np.diag([1.0, -1.0, -1.0]))), [0, 1, 0, 0])


def test_empty_and_degenerate():
    assert np.isnan(MO.pair_fmr(np.zeros((0, 3)), np.zeros((0, 3)), np.eye(4), 0.1))
    prec, rec, flags, errors = MO.evaluate_registration(4, np.zeros((0, 4, 4)), np.zeros((0, 3)), np.array([[0, 2, 4], [1, 3, 4]]),
                                                        np.tile(np.eye(4), (2, 1, 1)), np.tile(np.eye(6), (2, 1, 1)))
    assert prec == 0.0 and rec == 0.0 and flags == [] and errors == []
