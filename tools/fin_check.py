import sys, torch
sys.path.insert(0, '.')
from yoho_b200 import synth
from yoho_b200.engine import get_engine
e = get_engine(); e.set_gconv_impl("tcgen05_fourier"); e.load_part1(synth.synth_state_dict("PartI", 2))
x, _ = synth.make_fragment(1001, 41)
outs = {}
for f in (259, 2307, 6403):
    e.set_tuning(0, f); o = e.part1(x); torch.cuda.synchronize()
    outs[f] = {k: v.clone() for k, v in o.items()}
for k in ("eqv", "inv", "desc"):
    print(k, "staged == unstaged:", torch.equal(outs[6403][k], outs[2307][k]), " max|tc - simt| =", float((outs[2307][k] - outs[259][k]).abs().max()))
