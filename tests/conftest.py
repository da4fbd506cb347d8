import os
import sys
import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE = os.path.join(ROOT, "oracle")
if ORACLE not in sys.path:
    sys.path.insert(0, ORACLE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REAL_CKPT = os.path.join(ROOT, "oracle", "_ref", "ckpt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def real_ckpt(part):
    p = os.path.join(REAL_CKPT, part + ".npz")
    if not os.path.exists(p):
        return None
    return dict(np.load(p))


@pytest.fixture(scope="session")
def tables():
    import yoho_oracle as O
    return O.load_tables()


@pytest.fixture(scope="session")
def _engine_session():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from yoho_b200.engine import get_engine
    return get_engine()


GCONV_IMPLS = ["simt", "tcgen05", "tcgen05_split", "tcgen05_fourier"]


@pytest.fixture(params=GCONV_IMPLS)
def engine(request, _engine_session):
    """Every GPU test runs once per group-convolution implementation (FP32 SIMT, tcgen05, tcgen05 split-accumulator,
    tcgen05 with PartI layers 2+3 in the group-Fourier domain)."""
    _engine_session.set_gconv_impl(request.param)
    _engine_session.impl_name = request.param
    yield _engine_session
    _engine_session.set_gconv_impl("tcgen05_fourier")
