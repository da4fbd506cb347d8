"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end of oracle/estimator_oracle.c (E2-E5)."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libestimator_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "estimator_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        d, i32, i8, u8 = (ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32),
                          ctypes.POINTER(ctypes.c_int8), ctypes.POINTER(ctypes.c_uint8))
        L.yoho_oracle_kabsch3.argtypes = [d, d, i32, ctypes.c_int, d]
        L.yoho_oracle_kabsch3.restype = ctypes.c_int
        L.yoho_oracle_count_inliers.argtypes = [d, d, ctypes.c_int32, d, ctypes.c_double, u8]
        L.yoho_oracle_count_inliers.restype = ctypes.c_int32
        L.yoho_oracle_yohoc.argtypes = [d, d, ctypes.c_int32, i32, ctypes.c_int32, i8, d, ctypes.c_double,
                                        d, i32, i32, u8, i32, u8]
        L.yoho_oracle_yohoc.restype = None
        L.yoho_oracle_yohoo.argtypes = [d, d, ctypes.c_int32, d, ctypes.c_int32, ctypes.c_double,
                                        d, i32, i32, u8, i32]
        L.yoho_oracle_yohoo.restype = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def kabsch3(k0, k1, ids, sign_override=0):
    k0 = np.ascontiguousarray(k0, np.float64); k1 = np.ascontiguousarray(k1, np.float64)
    ids = np.ascontiguousarray(ids, np.int32)
    T = np.zeros(12)
    dg = lib().yoho_oracle_kabsch3(_p(k0, ctypes.c_double), _p(k1, ctypes.c_double), _p(ids, ctypes.c_int32),
                                   int(sign_override), _p(T, ctypes.c_double))
    return T.reshape(3, 4), bool(dg)


def count_inliers(k0, k1, T, dist):
    k0 = np.ascontiguousarray(k0, np.float64); k1 = np.ascontiguousarray(k1, np.float64)
    T = np.ascontiguousarray(T, np.float64).reshape(-1)
    mask = np.zeros(k0.shape[0], np.uint8)
    n = lib().yoho_oracle_count_inliers(_p(k0, ctypes.c_double), _p(k1, ctypes.c_double), k0.shape[0],
                                        _p(T, ctypes.c_double), float(dist) * float(dist), _p(mask, ctypes.c_uint8))
    return int(n), mask


def yohoc(k0, k1, hyp, dist, signs=None, fixed=None):
    """Returns dict(T[3,4], best_iter, n_inl, mask[M], counts[iters], degenerate[iters])."""
    k0 = np.ascontiguousarray(k0, np.float64); k1 = np.ascontiguousarray(k1, np.float64)
    hyp = np.ascontiguousarray(hyp, np.int32).reshape(-1, 3)
    M, iters = k0.shape[0], hyp.shape[0]
    sg = None if signs is None else np.ascontiguousarray(signs, np.int8)
    fx = None if fixed is None else np.ascontiguousarray(fixed, np.float64).reshape(-1)
    T = np.zeros(12); bi = np.zeros(1, np.int32); ni = np.zeros(1, np.int32)
    mask = np.zeros(M, np.uint8); counts = np.zeros(iters, np.int32); dg = np.zeros(iters, np.uint8)
    lib().yoho_oracle_yohoc(_p(k0, ctypes.c_double), _p(k1, ctypes.c_double), M, _p(hyp, ctypes.c_int32), iters,
                            _p(sg, ctypes.c_int8), _p(fx, ctypes.c_double), float(dist), _p(T, ctypes.c_double), _p(bi, ctypes.c_int32),
                            _p(ni, ctypes.c_int32), _p(mask, ctypes.c_uint8), _p(counts, ctypes.c_int32),
                            _p(dg, ctypes.c_uint8))
    return dict(T=T.reshape(3, 4), best_iter=int(bi[0]), n_inl=int(ni[0]), mask=mask, counts=counts,
                degenerate=dg.astype(bool))


def yohoo(k0, k1, trans, dist):
    k0 = np.ascontiguousarray(k0, np.float64); k1 = np.ascontiguousarray(k1, np.float64)
    trans = np.ascontiguousarray(trans, np.float64).reshape(-1, 12)
    M, H = k0.shape[0], trans.shape[0]
    T = np.zeros(12); bi = np.zeros(1, np.int32); ni = np.zeros(1, np.int32)
    mask = np.zeros(M, np.uint8); counts = np.zeros(H, np.int32)
    lib().yoho_oracle_yohoo(_p(k0, ctypes.c_double), _p(k1, ctypes.c_double), M, _p(trans, ctypes.c_double), H,
                            float(dist), _p(T, ctypes.c_double), _p(bi, ctypes.c_int32), _p(ni, ctypes.c_int32),
                            _p(mask, ctypes.c_uint8), _p(counts, ctypes.c_int32))
    return dict(T=T.reshape(3, 4), best_iter=int(bi[0]), n_inl=int(ni[0]), mask=mask, counts=counts)
