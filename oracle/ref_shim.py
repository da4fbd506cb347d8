"""TEST INFRASTRUCTURE — process-level shim that lets the UNMODIFIED reference be imported on a
CUDA-less box with current numpy/torch.  Used only by `tests/golden/make_golden.py` and by the
`not gpu` oracle-pinning tests, and only when /root/reference exists (authoring container).

Nothing in the product (`yoho_b200/`) imports this file.

What is patched (reference files are never touched), see SURVEY.md §8c:
  * np.int / np.float aliases          (utils/network.py:72, tests/extractor.py:67,131 …)
  * Tensor.cuda / Module.cuda identity (utils/network.py:72-74, tests/extractor.py:21,54)
  * torch.load -> map_location=cpu, weights_only=False (tests/extractor.py:29,117)
  * stub modules tensorboardX, open3d, nibabel (utils/utils.py:11, utils/dataset.py:20, utils/RR_cal.py:10)
  * sys.argv reset before the module-level argparse in parses/*.py
"""
import os
import sys
import types
import contextlib

# /root/reference in the authoring container; on the GPU box the git-ignored copy of the hot-path .py files that
# `__graft_entry__.build()` makes under oracle/_ref/src (never committed; travels with the gpurun snapshot like the checkpoints)
_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "src")
REF_ROOT = os.environ.get("YOHO_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/utils/network.py") else _LOCAL_COPY)

# what build() copies: the path's own modules + what they import at module level (none of it is product source)
HOT_PATH_FILES = [
    "utils/__init__.py", "utils/network.py", "utils/knn_search.py", "utils/r_eval.py", "utils/utils.py", "utils/dataset.py",
    "utils/RR_cal.py", "utils/misc.py", "utils/utils_o3d.py",
    "tests/__init__.py", "tests/extractor.py", "tests/matcher.py", "tests/estimator.py", "tests/evaluator.py",
    "parses/parses_partI.py", "parses/parses_partII.py", "train/loss_val.py",
    "group_related/Rotation.npy", "group_related/60_60.npy", "group_related/Nei_Index_in_SO3_ordered_13.npy",
]


def export_hot_path(src_root="/root/reference", dst_root=_LOCAL_COPY):
    """Copy the reference's hot-path files verbatim into oracle/_ref/src (git-ignored).  Called by build() only."""
    import shutil
    n = 0
    for rel in HOT_PATH_FILES:
        s = os.path.join(src_root, rel)
        if not os.path.isfile(s):
            continue
        d = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copyfile(s, d)
        n += 1
    return n


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "utils")) and os.path.isfile(
        os.path.join(REF_ROOT, "utils", "network.py"))


_installed = False


def install():
    """Apply the shim and put the reference root on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference not present at {REF_ROOT}")
    import numpy as np
    import torch

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    _orig_load = torch.load

    def _load(f, *a, **k):
        if not torch.cuda.is_available():
            k.setdefault("map_location", "cpu")
        k.setdefault("weights_only", False)
        return _orig_load(f, *a, **k)

    torch.load = _load

    for name in ("tensorboardX", "open3d", "nibabel", "nibabel.quaternions"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "tensorboardX":
                m.SummaryWriter = object
            sys.modules[name] = m

    if REF_ROOT not in sys.path:
        # the reference's packages are literally called `utils`, `tests`, `parses`
        sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def _argv(argv):
    old = sys.argv
    sys.argv = argv
    try:
        yield
    finally:
        sys.argv = old


def load_reference():
    """Import the reference's hot-path modules.  Returns a namespace with
    network / extractor / matcher / estimator / knn_search / r_eval / utils / cfgI / cfgII."""
    install()
    # `tests` collides with this repo's own tests/ package name when pytest has it imported:
    # evict any non-reference `tests`/`utils` first.
    for pkg in ("tests", "utils", "parses"):
        m = sys.modules.get(pkg)
        if m is not None and not str(getattr(m, "__file__", "") or getattr(m, "__path__", [""])[0]).startswith(REF_ROOT):
            for key in [k for k in sys.modules if k == pkg or k.startswith(pkg + ".")]:
                del sys.modules[key]
    import importlib
    with _argv(["ref"]):
        pI = importlib.import_module("parses.parses_partI")
        pII = importlib.import_module("parses.parses_partII")
        cfgI, _ = pI.get_config()
        cfgII, _ = pII.get_config()
    ns = types.SimpleNamespace()
    ns.network = importlib.import_module("utils.network")
    ns.knn_search = importlib.import_module("utils.knn_search")
    ns.r_eval = importlib.import_module("utils.r_eval")
    ns.utils = importlib.import_module("utils.utils")
    ns.extractor = importlib.import_module("tests.extractor")
    ns.matcher = importlib.import_module("tests.matcher")
    ns.estimator = importlib.import_module("tests.estimator")
    for cfg in (cfgI, cfgII):
        cfg.SO3_related_files = os.path.join(REF_ROOT, "group_related")
        cfg.model_fn = os.path.join(REF_ROOT, "model")
    ns.cfgI, ns.cfgII = cfgI, cfgII
    ns.root = REF_ROOT
    return ns
