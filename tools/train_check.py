"""Per-call check of the group-convolution backward kernels on the tensors an actual training step feeds them."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import yoho_oracle as O
from yoho_b200 import synth
import yoho_b200.train as T

R, P, N = O.load_tables()
dev = torch.device("cuda")
sd = synth.to_torch_state_dict(synth.synth_state_dict("PartI", 4))
pr = synth.make_fragment_pair(24, seed=9, overlap=1.0, sigma=0.1)
f0, f1 = pr["feat_B"][pr["ids_B"]], pr["feat_A"][pr["ids_A"]]
ti = np.full((24,), pr["r"], np.int64)
calls = []
orig_bwd = T._GroupConv.backward


def spy(ctx, dy):
    out = orig_bwd(ctx, dy)
    x, w = ctx.saved_tensors
    calls.append((x.detach().clone(), w.detach().clone(), dy.detach().clone(), [None if o is None else o.detach().clone() for o in out]))
    return out


T._GroupConv.backward = staticmethod(spy)
m = T.PartI_train(None).to(dev)
m.load_state_dict(sd, strict=True)
m.train()
out = m({"feats0": torch.from_numpy(f0).to(dev), "feats1": torch.from_numpy(f1).to(dev), "true_idx": torch.from_numpy(ti).to(dev)})
T.Batch_hard_Rindex_loss()(out).backward()
torch.cuda.synchronize()
for i, (x, w, dy, (dx, dw, db)) in enumerate(calls):
    xr, wr = x.double().cpu().requires_grad_(True), w.double().cpu().requires_grad_(True)
    yr = torch.nn.functional.conv2d(O.gather13(xr, N), wr)[:, :, :, 0]
    yr.backward(dy.double().cpu())
    r = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30))
    print(f"call {i}: x{tuple(x.shape)} w{tuple(w.shape)} |dy|max {float(dy.abs().max()):.2e} contiguous {dy.is_contiguous()} "
          f"dx {'-' if dx is None else '%.2e' % r(dx, xr.grad)} dw {'-' if dw is None else '%.2e' % r(dw, wr.grad)} "
          f"db {'-' if db is None else '%.2e' % r(db, dy.double().cpu().sum((0, 2)))}")

# ---- the same step with float64 torch stand-ins, capturing the tensors at the same eight points ------------------------------
ref_calls = []


class RefConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        y = torch.nn.functional.conv2d(O.gather13(x.cpu(), N).to(x.device), w, b)[:, :, :, 0]
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        with torch.enable_grad():
            xr, wr = x.detach().requires_grad_(True), w.detach().requires_grad_(True)
            y = torch.nn.functional.conv2d(O.gather13(xr.cpu(), N).to(x.device), wr)[:, :, :, 0]
            gx, gw = torch.autograd.grad(y, (xr, wr), dy)
        ref_calls.append((x.detach().clone(), dy.detach().clone(), gx.clone(), gw.clone()))
        return gx, gw, dy.sum((0, 2))


T._GroupConv.backward = staticmethod(orig_bwd)
idx = torch.from_numpy(P.reshape(-1)).to(dev)
T.group_conv = lambda x, w, b=None: RefConv.apply(x, w, b)
T.rot_correlation = lambda a, b: torch.einsum('bfag,bfg->ba', a[:, :, idx].reshape(a.shape[0], 32, 60, 60), b)
T.Des2DR = lambda a, b: torch.argmax(T.rot_correlation(a.detach(), b.detach()), 1)
m2 = T.PartI_train(None).to(dev)
m2.load_state_dict(sd, strict=True)
m2 = m2.double()
m2.train()
out2 = m2({"feats0": torch.from_numpy(f0).to(dev).double(), "feats1": torch.from_numpy(f1).to(dev).double(), "true_idx": torch.from_numpy(ti).to(dev)})
T.Batch_hard_Rindex_loss()(out2).backward()
rr = lambda a, b: float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))
for i, ((x, w, dy, (dx, dw, db)), (x64, dy64, gx64, gw64)) in enumerate(zip(calls, ref_calls)):
    print(f"call {i}: x {rr(x, x64):.2e} dy {rr(dy, dy64):.2e} dx {'-' if dx is None else '%.2e' % rr(dx, gx64)} dw {rr(dw, gw64):.2e} "
          f"|dw|max {float(gw64.abs().max()):.2e}")
