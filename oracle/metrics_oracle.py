"""TEST INFRASTRUCTURE — CPU restatement (numpy, float64) of the reference's evaluation metrics.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product never does.

Follows, function by function:
  * feature-match ratio of one pair ........ tests/evaluator.py:49-71 (`Feature_match_Recall`), utils/utils.py:221-228
    (`evaluate_the_match`), utils/utils.py:35-50 (`transform_points` and the homogeneous helpers)
  * rotation / translation error ........... utils/RR_cal.py:13-46
  * Redwood transformation error ........... utils/RR_cal.py:48-65 (`computeTransformationErr`)
  * registration precision / recall ........ utils/RR_cal.py:236-316 (`evaluate_registration`)

Third-party dependency NOT in /root/reference and not installed in this image: `nibabel` (requirements.txt:9, unpinned),
used for `nibabel.quaternions.mat2quat` (utils/RR_cal.py:10,61).  `mat2quat` below restates nibabel's published algorithm
(Bar-Itzhack: largest eigenvector of the symmetric 4x4 `K` built from the matrix, w >= 0).  PARITY UNPINNED for that one
function: there is no nibabel here to run it against; it is checked instead against the closed-form quaternion of exact
rotation matrices (tests/test_metrics_oracle.py).  Everything else in this file is pinned against the reference's own
functions run in the authoring container (tests/test_oracle_vs_reference.py::test_metrics_*).
"""
import math
import numpy as np


# ---- utils/utils.py:35-50 ---------------------------------------------------------------------------------------------
def transform_points(pts, transform):
    h, w = transform.shape
    if h == 3 and w == 3:
        return pts @ transform.T
    if h == 3 and w == 4:
        return pts @ transform[:, :3].T + transform[:, 3:].T
    if h == 4 and w == 4:
        hp = np.concatenate([pts, np.ones([pts.shape[0], 1])], 1) @ transform.T
        return hp[:, :-1] / hp[:, -1:]
    raise NotImplementedError


def match_ok_count(keys0, keys1, gt, threshold):
    """tests/evaluator.py:57-66: number of matches whose keypoints are closer than `threshold` under the ground truth.
    `keys0`, `keys1`: the matched keypoints [M,3] float64 (already gathered by the match list)."""
    k1 = transform_points(keys1, gt)
    dist = np.sqrt(np.sum(np.square(keys0 - k1), axis=-1))
    return int(np.sum(dist < threshold))


def pair_fmr(keys0, keys1, gt, threshold):
    """tests/evaluator.py:66: `np.mean(dist<threshold)`; NaN (with numpy's warning) for an empty match list, as numpy gives."""
    m = keys0.shape[0]
    if m == 0:
        return float("nan")
    return match_ok_count(keys0, keys1, gt, threshold) / m


def scene_fmr(pair_fmrs, ratio=0.05):
    """tests/evaluator.py:68-70."""
    pair_fmrs = np.asarray(pair_fmrs, dtype=np.float64)
    return float(np.mean(pair_fmrs > ratio)), pair_fmrs


# ---- utils/RR_cal.py:13-46 --------------------------------------------------------------------------------------------
PI_F32 = float(np.float32(math.pi))      # RR_cal.py:31-32: `torch.Tensor([math.pi])` is float32, then cast to the error's dtype


def rotation_error(R1, R2):
    """RR_cal.py:13-33, float64 inputs [b,3,3]: degrees, [b,1]."""
    R_ = np.matmul(np.transpose(R1, (0, 2, 1)), R2)
    e = (np.trace(R_, axis1=1, axis2=2) - 1) / 2
    e = np.clip(e, -1, 1)
    return (180.0 * np.arccos(e) / PI_F32)[:, None]


def translation_error(t1, t2):
    """RR_cal.py:35-46, [b,3,1] each: Frobenius norm over the last two axes, [b]."""
    d = t1 - t2
    return np.sqrt(np.sum(d * d, axis=(1, 2)))


# ---- nibabel.quaternions.mat2quat (published algorithm, see the header) -------------------------------------------------
def mat2quat(M):
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, dtype=np.float64).flat
    K = np.array([
        [Qxx - Qyy - Qzz, 0, 0, 0],
        [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
        [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
        [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)          # reads the lower triangle
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    if q[0] < 0:
        q = q * -1
    return q


# ---- utils/RR_cal.py:48-65 --------------------------------------------------------------------------------------------
def compute_transformation_err(trans, info):
    t = trans[:3, 3]
    r = trans[:3, :3]
    q = mat2quat(r)
    er = np.concatenate([t, q[1:]], axis=0)
    p = er.reshape(1, 6) @ info @ er.reshape(6, 1) / info[0, 0]
    return p.item()


def registration_errors(est, gt, info):
    """The numeric core of RR_cal.py:273-301 and :353-354 for aligned lists: est[n,4,4], gt[n,4,4], info[n,6,6] ->
    (p[n] Redwood error before the square root, rre_deg[n], rte[n])."""
    n = est.shape[0]
    p = np.zeros(n)
    for i in range(n):
        p[i] = compute_transformation_err(np.linalg.inv(gt[i]) @ est[i], info[i])
    re = rotation_error(gt[:, 0:3, 0:3], est[:, 0:3, 0:3])[:, 0]
    te = translation_error(gt[:, 0:3, 3:4], est[:, 0:3, 3:4])
    return p, re, te


# ---- utils/RR_cal.py:236-316 ------------------------------------------------------------------------------------------
def evaluate_registration(num_fragment, result, result_pairs, gt_pairs, gt, gt_info, err2=0.2, nonconsecutive=True):
    err2 = err2 ** 2
    gt_mask = np.zeros((num_fragment, num_fragment), dtype=int)
    flags, errors = [], []
    if nonconsecutive:
        for idx in range(gt_pairs.shape[0]):
            i, j = int(gt_pairs[idx, 0]), int(gt_pairs[idx, 1])
            if abs(j - i) > 1:
                gt_mask[i, j] = idx
        n_gt = np.sum(gt_mask > 0)
    else:
        for idx in range(gt_pairs.shape[0]):
            i, j = int(gt_pairs[idx, 0]), int(gt_pairs[idx, 1])
            gt_mask[i, j] = idx
        n_gt = np.sum(gt_mask > 0) + 1
    good, n_res = 0, 0
    if not nonconsecutive:
        start_check = 1
        n_res += 1
        p = compute_transformation_err(np.linalg.inv(gt[0]) @ result[0], gt_info[0])
        errors.append(np.sqrt(p))
        if p <= err2:
            good += 1
            flags.append(0)
        else:
            flags.append(1)
    else:
        start_check = 0
    for idx in range(start_check, result_pairs.shape[0]):
        i, j = int(result_pairs[idx, 0]), int(result_pairs[idx, 1])
        if gt_mask[i, j] > 0:
            n_res += 1
            gt_idx = gt_mask[i, j]
            p = compute_transformation_err(np.linalg.inv(gt[gt_idx]) @ result[idx], gt_info[gt_idx])
            errors.append(np.sqrt(p))
            if p <= err2:
                good += 1
                flags.append(0)
            else:
                flags.append(1)
        else:
            flags.append(2)
    if n_res == 0:
        n_res += 1e6
    return good * 1.0 / n_res, good * 1.0 / n_gt, flags, errors
