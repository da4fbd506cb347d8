"""Golden vectors for the evaluation metrics, produced by the UNMODIFIED reference functions imported from
/root/reference under oracle/ref_shim.py.  Run in the authoring container only:

    python tests/golden/make_golden_metrics.py        ->  tests/golden/metrics_synth.npz

`nibabel` is not installed here, so `nibabel.quaternions.mat2quat` is supplied by the oracle's restatement of nibabel's
published algorithm (oracle/metrics_oracle.py) — the Redwood-error golden values are pinned to the reference's code AROUND
that call, not to nibabel itself.  Everything else (rotation / translation error, the pair bookkeeping of
evaluate_registration, evaluate_the_match) is the reference's arithmetic end to end.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import ref_shim  # noqa: E402
import metrics_oracle as MO  # noqa: E402


def rand_rigid(rs, angle_deg=None, trans=1.0):
    axis = rs.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(angle_deg if angle_deg is not None else rs.uniform(0, 180))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = rs.standard_normal(3) * trans
    return T


def rand_info(rs, n=400):
    """Redwood-style information matrix: sum over correspondences of G^T G, G = [I | -[p]x]."""
    p = rs.uniform(-1.5, 1.5, (n, 3))
    info = np.zeros((6, 6))
    for q in p:
        G = np.zeros((3, 6))
        G[:, :3] = np.eye(3)
        G[:, 3:] = -np.array([[0, -q[2], q[1]], [q[2], 0, -q[0]], [-q[1], q[0], 0]])
        info += G.T @ G
    return info


def make_scene(seed, n_frag=12):
    rs = np.random.RandomState(seed)
    gt_pairs, gt, info = [], [], []
    for i in range(n_frag):
        for j in range(i + 1, n_frag):
            if j == i + 1 or rs.rand() < 0.45:
                gt_pairs.append([i, j, n_frag])
                gt.append(rand_rigid(rs))
                info.append(rand_info(rs))
    gt_pairs, gt, info = np.array(gt_pairs), np.array(gt), np.array(info)
    est_pairs, est = [], []
    for i in range(n_frag):
        for j in range(i + 1, n_frag):
            if rs.rand() < 0.7:
                est_pairs.append([i, j, n_frag])
                hit = np.where((gt_pairs[:, 0] == i) & (gt_pairs[:, 1] == j))[0]
                if len(hit) and rs.rand() < 0.7:       # a good estimate: small perturbation of the ground truth
                    est.append(gt[hit[0]] @ rand_rigid(rs, angle_deg=rs.uniform(0, 6), trans=0.05))
                else:
                    est.append(rand_rigid(rs))
    return n_frag, np.array(est), np.array(est_pairs), gt_pairs, gt, info


def main():
    ref_shim.install()
    sys.modules["nibabel.quaternions"].mat2quat = MO.mat2quat
    sys.modules["nibabel"].quaternions = sys.modules["nibabel.quaternions"]
    import importlib
    RR = importlib.import_module("utils.RR_cal")
    U = importlib.import_module("utils.utils")
    out = {}
    for tag, seed, noncons in (("a", 11, True), ("b", 12, False)):
        n_frag, est, est_pairs, gt_pairs, gt, info = make_scene(seed)
        prec, rec, flags, errors = RR.evaluate_registration(n_frag, est, est_pairs, gt_pairs, gt, info, err2=0.2, nonconsecutive=noncons)
        # aligned ground truth per estimated pair (what benchmark() feeds rotation_error / translation_error, RR_cal.py:351-354)
        ext = np.zeros((len(est_pairs), 4, 4))
        for k, pr in enumerate(est_pairs):
            hit = np.where((gt_pairs[:, 0] == pr[0]) & (gt_pairs[:, 1] == pr[1]))[0]
            ext[k] = gt[hit[0]] if len(hit) else np.eye(4)
        re = RR.rotation_error(torch.from_numpy(ext[:, 0:3, 0:3]), torch.from_numpy(est[:, 0:3, 0:3])).numpy()
        te = RR.translation_error(torch.from_numpy(ext[:, 0:3, 3:4]), torch.from_numpy(est[:, 0:3, 3:4])).numpy()
        p_all = np.array([RR.computeTransformationErr(np.linalg.inv(ext[k]) @ est[k], info[min(k, len(info) - 1)]) for k in range(len(est))])
        out.update({f"{tag}_n_frag": n_frag, f"{tag}_est": est, f"{tag}_est_pairs": est_pairs, f"{tag}_gt_pairs": gt_pairs, f"{tag}_gt": gt,
                    f"{tag}_info": info, f"{tag}_nonconsecutive": noncons, f"{tag}_precision": prec, f"{tag}_recall": rec,
                    f"{tag}_flags": np.array(flags), f"{tag}_errors": np.array(errors), f"{tag}_ext_gt": ext, f"{tag}_re": re, f"{tag}_te": te,
                    f"{tag}_p_all": p_all})
        print(tag, "pairs", len(est_pairs), "precision", prec, "recall", rec, "flags", np.bincount(np.array(flags), minlength=3))
    # feature-match ratios (utils/utils.py:221-228 == tests/evaluator.py:57-66)
    rs = np.random.RandomState(5)
    ratios = []
    for k in range(6):
        n0, n1, M = 300 + 17 * k, 280 + 13 * k, 150 + 31 * k
        T = rand_rigid(rs)
        kps1 = rs.uniform(0, 3, (n1, 3))
        kps0 = rs.uniform(0, 3, (n0, 3))
        matches = np.stack([rs.randint(0, n0, M), rs.randint(0, n1, M)], 1)
        good = rs.rand(M) < 0.1 * (k + 1)
        kps0[matches[good, 0]] = (kps1[matches[good, 1]] @ T[:3, :3].T + T[:3, 3]) + rs.standard_normal((int(good.sum()), 3)) * 0.04
        gt = T if k % 2 == 0 else T[:3]
        ratios.append(U.evaluate_the_match(kps0, kps1, matches, gt, 0.1))
        out.update({f"fmr{k}_kps0": kps0, f"fmr{k}_kps1": kps1, f"fmr{k}_matches": matches, f"fmr{k}_gt": gt})
    out["fmr_ratios"] = np.array(ratios)
    out["fmr_threshold"] = 0.1
    print("fmr ratios", ratios)
    np.savez_compressed(os.path.join(HERE, "metrics_synth.npz"), **out)


if __name__ == "__main__":
    main()
