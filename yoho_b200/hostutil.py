"""Small host helpers with the reference's names and semantics (utils/utils.py:13-15,42-50,108-136;
utils/r_eval.py:94-110).  Bookkeeping only — no hot-path arithmetic lives here."""
import os
import numpy as np
import torch


def make_non_exists_dir(fn):
    if not os.path.exists(fn):
        os.makedirs(fn)


def transform_points(pts, transform):
    """utils/utils.py:42-50: 3x3 / 3x4 / 4x4 transform of row-vector points."""
    h, w = transform.shape
    if h == 3 and w == 3:
        return pts @ transform.T
    if h == 3 and w == 4:
        return pts @ transform[:, :3].T + transform[:, 3:].T
    if h == 4 and w == 4:
        hp = np.concatenate([pts, np.ones([pts.shape[0], 1])], 1) @ transform.T
        return hp[:, :-1] / hp[:, -1:]
    raise NotImplementedError


def matrix_from_quaternion(quaternion):
    """utils/r_eval.py:94-110 (w,x,y,z)."""
    w, x, y, z = quaternion[0], quaternion[1], quaternion[2], quaternion[3]
    mat = np.eye(3)
    mat[0, 0] = 1 - 2 * y * y - 2 * z * z
    mat[0, 1] = 2 * x * y - 2 * z * w
    mat[0, 2] = 2 * x * z + 2 * y * w
    mat[1, 0] = 2 * x * y + 2 * z * w
    mat[1, 1] = 1 - 2 * x * x - 2 * z * z
    mat[1, 2] = 2 * y * z - 2 * x * w
    mat[2, 0] = 2 * x * z - 2 * y * w
    mat[2, 1] = 2 * y * z + 2 * x * w
    mat[2, 2] = 1 - 2 * x * x - 2 * y * y
    return mat


def to_cuda(data):
    """utils/utils.py:108-136: move a tensor / list / dict of tensors to the GPU."""
    if isinstance(data, torch.Tensor):
        return data.cuda()
    if isinstance(data, (list, tuple)):
        return [to_cuda(d) for d in data]
    if isinstance(data, dict):
        return {k: to_cuda(v) for k, v in data.items()}
    raise NotImplementedError


def feature_set_name(dataset_name):
    """'3dLo...' datasets reuse the features/keypoints of '3d...' (tests/extractor.py:84-88,
    tests/matcher.py:24-28, tests/estimator.py:84-88)."""
    if dataset_name[0:4] == '3dLo':
        return f'3d{dataset_name[4:]}'
    return dataset_name


class numa_local:
    """Context manager: run the enclosed block on the CPUs that are NUMA-local to GPU `index` (NVML's ideal affinity for the
    device), then restore the previous affinity.  Pinned host buffers allocated inside are first-touched on the memory node the
    GPU's PCIe root hangs off, which is what the H2D DMA reads fastest.  A no-op when NVML or the affinity call is missing."""

    def __init__(self, index=0):
        self.index, self.prev, self.note = int(index), None, "not attempted"

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            n_words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
            cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus and cpus != allowed:
                self.prev = allowed
                os.sched_setaffinity(0, cpus)
                self.note = f"pinned-buffer first touch bound to {len(cpus)} GPU-local CPUs of {len(allowed)} allowed"
            else:
                self.note = f"no binding needed: NVML's affinity for GPU {self.index} covers all {len(allowed)} allowed CPUs"
        except Exception as e:  # noqa: BLE001
            self.prev = None
            self.note = f"no-op ({type(e).__name__})"
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:
                pass
        return False
