"""CPU, authoring container only (skipped where /root/reference is absent): the oracle against the UNMODIFIED reference
imported under oracle/ref_shim.py, stage by stage, with the shipped checkpoints and on seeds the goldens do not use."""
import os
import numpy as np
import pytest
import torch

import ref_shim
import yoho_oracle as O
from yoho_b200 import synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present on this box")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def _ckpt(ref, part):
    fn = {"PartI": "PartI_train", "PartII": "PartII_train"}[part]
    return torch.load(os.path.join(ref.root, "model", fn, "model_best.pth"))["network_state_dict"]


def test_part1_bit_identical(ref, tables):
    _, _, N = tables
    sd = _ckpt(ref, "PartI")
    net = ref.network.PartI_test(ref.cfgI)
    net.load_state_dict(sd)
    net.eval()
    x, _ = synth.make_fragment(64, 77)
    with torch.no_grad():
        want = net(torch.from_numpy(x))
    got = O.part1_forward(x, sd, N)
    assert torch.equal(got["eqv"], want["eqv"]) and torch.equal(got["inv"], want["inv"])
    got_cost = O.part1_forward(x, sd, N, faithful_cost=True)          # BN/ReLU after the gather, as the reference orders it
    assert torch.equal(got_cost["eqv"], want["eqv"])


def test_matcher_and_rotation_index(ref, tables):
    _, P, _ = tables
    rs = np.random.RandomState(5)
    d0 = (rs.standard_normal((700, 32)) * 0.1).astype(np.float32)
    d1 = (rs.standard_normal((650, 32)) * 0.1).astype(np.float32)
    d1[:300] = d0[rs.permutation(700)[:300]] + (rs.standard_normal((300, 32)) * 0.01).astype(np.float32)
    knn = ref.knn_search.knn_module.KNN(1)
    _, a01 = knn(torch.from_numpy(d1.T.copy())[None], torch.from_numpy(d0.T.copy())[None])
    _, a10 = knn(torch.from_numpy(d0.T.copy())[None], torch.from_numpy(d1.T.copy())[None])
    pps, o01, o10 = O.mutual_matches(d0, d1)
    assert np.array_equal(o01, a01[0, 0].numpy()) and np.array_equal(o10, a10[0, 0].numpy())
    pr = synth.make_fragment_pair(40, seed=91, overlap=1.0, sigma=0.2)
    des1, des2 = pr["feat_B"][pr["ids_B"]], pr["feat_A"][pr["ids_A"]]
    want = ref.extractor.extractor_dr_index(ref.cfgI).Batch_Des2R_torch(torch.from_numpy(des1), torch.from_numpy(des2)).numpy()
    assert np.array_equal(O.rot_argmax(des1, des2, P)[0], want)


def test_part2_and_transforms(ref, tables):
    R, P, N = tables
    sd = _ckpt(ref, "PartII")
    net = ref.network.PartII_test(ref.cfgII)
    net.load_state_dict(sd, strict=False)
    net.eval()
    M = 20
    fA, kA = synth.make_fragment(M, 1)
    fB, kB = synth.make_fragment(M, 2)
    yA, _ = synth.make_fragment(M, 3)
    yB, _ = synth.make_fragment(M, 4)
    pre = np.random.RandomState(0).randint(0, 60, M).astype(np.int64)
    batch = {"before_eqv0": torch.from_numpy(fB.copy()), "before_eqv1": torch.from_numpy(fA.copy()),
             "after_eqv0": torch.from_numpy(yB.copy()), "after_eqv1": torch.from_numpy(yA.copy()), "pre_idx": torch.from_numpy(pre)}
    with torch.no_grad():
        want = net(batch)["quaternion_pre"].numpy()
    got = O.part2_forward(fA, fB, yA, yB, pre, sd, P, N).numpy()
    assert np.abs(got - want).max() <= 1e-6
    Rg = R.astype(np.float32)
    for i in range(M):       # tests/extractor.py:187-199
        Rm = ref.r_eval.matrix_from_quaternion(want[i]) @ Rg[int(pre[i])]
        t = kA[i] - kB[i] @ Rm.T
        T = O.part2_transforms(want[i:i + 1], pre[i:i + 1], kA[i:i + 1], kB[i:i + 1], R)[0]
        assert np.array_equal(T[:, :3], Rm) and np.array_equal(T[:, 3], t)


def test_metrics_against_reference(ref):
    """the metrics oracle against the reference's utils/RR_cal.py and utils/utils.py on seeds the golden file
    does not use (nibabel's mat2quat is supplied by the oracle's restatement: the package is not installed here)."""
    import sys
    import importlib
    import metrics_oracle as MO
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_metrics as G
    sys.modules["nibabel.quaternions"].mat2quat = MO.mat2quat
    sys.modules["nibabel"].quaternions = sys.modules["nibabel.quaternions"]
    RR = importlib.import_module("utils.RR_cal")
    for seed, noncons in ((101, True), (102, False), (103, True)):
        n_frag, est, est_pairs, gt_pairs, gt, info = G.make_scene(seed, n_frag=9)
        want = RR.evaluate_registration(n_frag, est, est_pairs, gt_pairs, gt, info, err2=0.2, nonconsecutive=noncons)
        got = MO.evaluate_registration(n_frag, est, est_pairs, gt_pairs, gt, info, err2=0.2, nonconsecutive=noncons)
        assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
        assert np.array_equal(np.array(got[3]), np.array(want[3]))
        m = min(len(est), len(gt))
        re = RR.rotation_error(torch.from_numpy(gt[:m, :3, :3]), torch.from_numpy(est[:m, :3, :3])).numpy()
        te = RR.translation_error(torch.from_numpy(gt[:m, :3, 3:4]), torch.from_numpy(est[:m, :3, 3:4])).numpy()
        assert np.allclose(MO.rotation_error(gt[:m, :3, :3], est[:m, :3, :3]), re, rtol=1e-12, atol=1e-10)
        assert np.allclose(MO.translation_error(gt[:m, :3, 3:4], est[:m, :3, 3:4]), te, rtol=1e-13, atol=0)
    rs = np.random.RandomState(9)
    k0, k1 = rs.uniform(0, 3, (200, 3)), rs.uniform(0, 3, (210, 3))
    mt = np.stack([rs.randint(0, 200, 120), rs.randint(0, 210, 120)], 1)
    T = G.rand_rigid(rs)
    k0[mt[:40, 0]] = k1[mt[:40, 1]] @ T[:3, :3].T + T[:3, 3] + rs.standard_normal((40, 3)) * 0.05
    for gtm in (T, T[:3]):
        assert MO.pair_fmr(k0[mt[:, 0]], k1[mt[:, 1]], gtm, 0.1) == ref.utils.evaluate_the_match(k0, k1, mt, gtm, 0.1)
