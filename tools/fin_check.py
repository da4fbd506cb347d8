"""Bring-up check of the tensor-core output side of PartI (tuning flags 2048 / 4096, csrc/fourier_tc.cu group_finalize_tc_kernel)
against the default FP32 SIMT finalize kernel: the two tensor-core variants must agree bit for bit with each other and to a
few 1e-6 with the SIMT kernel (different summation order of the channel norms).

    python tools/fin_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                      # noqa: E402
from yoho_b200 import synth                        # noqa: E402
from yoho_b200.engine import get_engine            # noqa: E402

e = get_engine()
e.set_gconv_impl("tcgen05_fourier")
e.load_part1(synth.synth_state_dict("PartI", 2))
x, _ = synth.make_fragment(1001, 41)               # not a multiple of four keypoints: the last tile is partial
outs = {}
for f in (e.DEFAULT_TUNING, e.DEFAULT_TUNING | 2048, e.DEFAULT_TUNING | 2048 | 4096):
    e.set_tuning(0, f)
    o = e.part1(x)
    torch.cuda.synchronize()
    outs[f] = {k: v.clone() for k, v in o.items()}
e.set_tuning(0, e.DEFAULT_TUNING)
a, b, c = (outs[k] for k in sorted(outs))
for k in ("eqv", "inv", "desc"):
    print(k, "staged == unstaged:", torch.equal(c[k], b[k]), " max|tensor-core - SIMT| =", float((b[k] - a[k]).abs().max()))
