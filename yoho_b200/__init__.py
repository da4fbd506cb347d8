"""yoho_b200 — B200-native implementation of YOHO's descriptor + registration hot path.

Python modules mirror the reference's plugin API (network / extractor / matcher / estimator / knn_search);
the arithmetic lives in hand-written sm_100a CUDA kernels behind the C ABI of include/yoho_b200.h.
"""
__version__ = "0.1.0"
