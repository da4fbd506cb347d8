"""Evaluation metrics on the B200 backend — the numeric functions of the reference's `utils/RR_cal.py`
(same names, argument meaning and return values) and the feature-match recall of `tests/evaluator.py:49-71`, with the
arithmetic on the device (`csrc/metrics.cu`) and the bookkeeping (pair masks, flags, precision / recall) on the host.

The trajectory / info file readers stay the reference's (`utils/RR_cal.py:67-213`): `yoho_b200.dropin.install(metrics=True)`
patches only the numeric entry points of that module, which also removes its run-time need for `nibabel`.
No CPU fallback: every function needs the CUDA library.
"""
import numpy as np
import torch

from .engine import get_engine


def _eng(engine=None):
    return engine if engine is not None else get_engine()


def rotation_error(R1, R2, engine=None):
    """utils/RR_cal.py:13-33.  R1, R2 [b,3,3] (tensors or arrays) -> float64 tensor [b,1], degrees."""
    eng = _eng(engine)
    R1, R2 = eng._f64(R1), eng._f64(R2)
    b = R1.shape[0]
    A = torch.zeros((b, 4, 4), dtype=torch.float64, device=eng.device)
    B = torch.zeros((b, 4, 4), dtype=torch.float64, device=eng.device)
    A[:, :3, :3], B[:, :3, :3] = R2, R1              # kernel convention: (est, gt); the error is symmetric in the trace anyway
    A[:, 3, 3] = 1.0
    B[:, 3, 3] = 1.0
    _, re, _ = eng.registration_errors(A, B, None)
    return re[:, None]


def translation_error(t1, t2, engine=None):
    """utils/RR_cal.py:35-46.  t1, t2 [b,3,1] -> float64 tensor [b]."""
    eng = _eng(engine)
    t1, t2 = eng._f64(t1), eng._f64(t2)
    b = t1.shape[0]
    A = torch.eye(4, dtype=torch.float64, device=eng.device).repeat(b, 1, 1)
    B = torch.eye(4, dtype=torch.float64, device=eng.device).repeat(b, 1, 1)
    A[:, :3, 3], B[:, :3, 3] = t2[:, :, 0], t1[:, :, 0]
    _, _, te = eng.registration_errors(A, B, None)
    return te


def transformation_errors(est, gt, info, engine=None):
    """Batched `computeTransformationErr(np.linalg.inv(gt[i]) @ est[i], info[i])` (utils/RR_cal.py:48-65 as called from
    :273,289): est, gt [n,4,4], info [n,6,6] -> numpy float64 [n] (the error BEFORE the square root)."""
    eng = _eng(engine)
    p, _, _ = eng.registration_errors(est, gt, info)
    return p.cpu().numpy()


def computeTransformationErr(trans, info, engine=None):
    """utils/RR_cal.py:48-65 for one already-composed transformation `trans` [4,4]."""
    eye = np.eye(4)[None]
    return float(transformation_errors(np.asarray(trans, dtype=np.float64)[None], eye, np.asarray(info, dtype=np.float64)[None], engine)[0])


def evaluate_registration(num_fragment, result, result_pairs, gt_pairs, gt, gt_info, err2=0.2, nonconsecutive=True, engine=None):
    """utils/RR_cal.py:236-316, same arguments and returns `(precision, recall, flags, errors)`.  The pair bookkeeping runs on
    the host exactly as the reference orders it; every Redwood error of the scene is evaluated in ONE device launch."""
    err2 = err2 ** 2
    result = np.asarray(result, dtype=np.float64)
    gt = np.asarray(gt, dtype=np.float64)
    gt_info = np.asarray(gt_info, dtype=np.float64)
    gt_mask = np.zeros((num_fragment, num_fragment), dtype=np.int64)
    for idx in range(gt_pairs.shape[0]):
        i, j = int(gt_pairs[idx, 0]), int(gt_pairs[idx, 1])
        if not nonconsecutive or abs(j - i) > 1:
            gt_mask[i, j] = idx
    n_gt = int(np.sum(gt_mask > 0)) + (0 if nonconsecutive else 1)
    # which (result row, gt row) combinations the reference evaluates, in its order
    todo = []                                             # (result idx, gt idx)
    flags = [None] * result_pairs.shape[0]
    start_check = 0
    if not nonconsecutive:
        todo.append((0, 0))
        start_check = 1
    for idx in range(start_check, result_pairs.shape[0]):
        i, j = int(result_pairs[idx, 0]), int(result_pairs[idx, 1])
        if gt_mask[i, j] > 0:
            todo.append((idx, int(gt_mask[i, j])))
        else:
            flags[idx] = 2
    errors = []
    good = 0
    if todo:
        ri = np.array([t[0] for t in todo])
        gi = np.array([t[1] for t in todo])
        p = transformation_errors(result[ri], gt[gi], gt_info[gi], engine)
        for (idx, _), pv in zip(todo, p):
            errors.append(np.sqrt(pv))
            if pv <= err2:
                good += 1
                flags[idx] = 0
            else:
                flags[idx] = 1
    n_res = len(todo)
    if n_res == 0:
        n_res += 1e6
    return good * 1.0 / n_res, good * 1.0 / n_gt, flags, errors


def pair_match_ratios(keys0_list, keys1_list, gts, threshold, engine=None):
    """tests/evaluator.py:57-66 for a list of pairs at once: keys0_list[p], keys1_list[p] are the MATCHED keypoints [M_p,3] of
    pair p (`Keys[id][matches[:,k]]`), gts[p] its ground truth (4x4 or 3x4).  Returns numpy float64 [n]: the ratio of matches
    closer than `threshold` (`np.mean(dist<threshold)`; NaN for an empty match list, as numpy's mean gives)."""
    eng = _eng(engine)
    n = len(keys0_list)
    if n == 0:
        return np.zeros((0,), np.float64)
    sizes = np.array([int(k.shape[0]) for k in keys0_list], dtype=np.int64)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(sizes)
    k0 = torch.cat([eng._f64(k).reshape(-1, 3) for k in keys0_list])
    k1 = torch.cat([eng._f64(k).reshape(-1, 3) for k in keys1_list])
    g = np.zeros((n, 4, 4))
    for p, t in enumerate(gts):
        t = np.asarray(t, dtype=np.float64)
        g[p] = np.eye(4)
        g[p, :t.shape[0], :t.shape[1]] = t
    counts = eng.fmr_counts(k0, k1, off, g, float(threshold)).cpu().numpy().astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return counts / sizes.astype(np.float64)


def evaluate_the_match(kps0, kps1, matches, transform_gt, threshold=0.1, engine=None):
    """utils/utils.py:221-228 (Demo.py:66)."""
    matches = np.asarray(matches)
    return float(pair_match_ratios([np.asarray(kps0)[matches[:, 0]]], [np.asarray(kps1)[matches[:, 1]]], [transform_gt], threshold, engine)[0])


def feature_match_recall(keys0_list, keys1_list, gts, threshold=0.1, ratio=0.05, engine=None):
    """Evaluator_PartI.Feature_match_Recall (tests/evaluator.py:49-71) on in-memory pairs -> (FMR, pair_fmrs)."""
    pair_fmrs = pair_match_ratios(keys0_list, keys1_list, gts, threshold, engine)
    return float(np.mean(pair_fmrs > ratio)), pair_fmrs
