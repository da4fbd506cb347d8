"""GPU: training-time twins against torch autograd on the oracle's formulation (the reference's own operator
sequence: gather `x[:,:,Nei]` + Conv2d(1,13); einsum correlation) in float64 on the CPU.  Bars: forward 1e-5 relative to the
output scale, gradients 1e-4 relative; the loss of the reference's Batch_hard_Rindex_loss and one SGD step of PartI_train agree
with the unmodified reference modules when its sources are available."""
import os
import sys
import numpy as np
import pytest
import torch

from conftest import ROOT
import yoho_oracle as O
from yoho_b200 import synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,C,Oc,bias", [(1, 32, 256, True), (5, 256, 512, True), (3, 512, 256, False), (4, 40, 70, True), (2, 256, 32, True)])
def test_group_conv_forward_backward_vs_autograd(_engine_session, tables, B, C, Oc, bias):
    from yoho_b200.train import group_conv
    _, _, N = tables
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.randn((B, C, 60), generator=g, dtype=torch.float64)
    w = torch.randn((Oc, C, 1, 13), generator=g, dtype=torch.float64) / np.sqrt(13 * C)
    b = torch.randn((Oc,), generator=g, dtype=torch.float64) if bias else None
    dy = torch.randn((B, Oc, 60), generator=g, dtype=torch.float64)
    xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True) if bias else None
    yr = torch.nn.functional.conv2d(O.gather13(xr, N), wr, br)[:, :, :, 0]          # utils/network.py:46-52 + :18
    yr.backward(dy)
    dev = _engine_session.device
    xg, wg = x.float().to(dev).requires_grad_(True), w.float().to(dev).requires_grad_(True)
    bg = b.float().to(dev).requires_grad_(True) if bias else None
    yg = group_conv(xg, wg, bg)
    yg.backward(dy.float().to(dev))
    assert _rel(yg.detach().cpu().double(), yr.detach()) <= 1e-5
    assert _rel(xg.grad.cpu().double(), xr.grad) <= 1e-4
    assert _rel(wg.grad.cpu().double(), wr.grad) <= 1e-4
    if bias:
        assert _rel(bg.grad.cpu().double(), br.grad) <= 1e-4
    # determinism: same bits on a second evaluation
    xg2, wg2 = xg.detach().clone().requires_grad_(True), wg.detach().clone().requires_grad_(True)
    y2 = group_conv(xg2, wg2, bg.detach() if bias else None)
    y2.backward(dy.float().to(dev))
    assert torch.equal(y2, yg) and torch.equal(xg2.grad, xg.grad) and torch.equal(wg2.grad, wg.grad)


def test_group_conv_equals_inference_layer(_engine_session, tables):
    """The training forward and the inference path's FP32 layer (yoho_debug_layer, SIMT) are the same operator."""
    from yoho_b200.train import group_conv
    sd = synth.synth_state_dict("PartI", 2)
    _engine_session.load_part1(sd)
    dev = _engine_session.device
    x, _ = synth.make_fragment(7, 3)                                      # [7,32,60]
    w = torch.from_numpy(sd["PartI_net.Conv_in.0.weight"]).to(dev)
    b = torch.from_numpy(sd["PartI_net.Conv_in.0.bias"]).to(dev)
    y = group_conv(torch.from_numpy(x).to(dev), w, b)                     # [7,256,60]
    act = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 1))).to(dev)     # [B,60,32]
    raw = _engine_session.debug_layer(0, "simt", act, 256)                # [B,60,256]
    assert (y.permute(0, 2, 1) - raw).abs().max().item() <= 2e-5


def test_rot_correlation_and_des2dr(_engine_session, tables):
    from yoho_b200.train import rot_correlation, Des2DR
    _, P, _ = tables
    dev = _engine_session.device
    a, _ = synth.make_fragment(37, 11)
    b, _ = synth.make_fragment(37, 12)
    g = torch.Generator().manual_seed(1)
    dc = torch.randn((37, 60), generator=g, dtype=torch.float64)
    ar, br = torch.from_numpy(a).double().requires_grad_(True), torch.from_numpy(b).double().requires_grad_(True)
    idx = torch.from_numpy(P.reshape(-1))
    cr = torch.einsum('bfag,bfg->ba', ar[:, :, idx].reshape(37, 32, 60, 60), br)     # train/loss_val.py:27-31, utils/network.py:115-117
    cr.backward(dc)
    ag, bg = torch.from_numpy(a).to(dev).requires_grad_(True), torch.from_numpy(b).to(dev).requires_grad_(True)
    cg = rot_correlation(ag, bg)
    cg.backward(dc.float().to(dev))
    assert _rel(cg.detach().cpu().double(), cr.detach()) <= 1e-5
    assert _rel(ag.grad.cpu().double(), ar.grad) <= 1e-4 and _rel(bg.grad.cpu().double(), br.grad) <= 1e-4
    want = torch.argmax(cr.detach(), dim=1)
    assert torch.equal(Des2DR(ag, bg).cpu(), want)
    y = np.stack([a[i][:, P[i % 60]] for i in range(37)])
    assert torch.equal(Des2DR(torch.from_numpy(a).to(dev), torch.from_numpy(y).to(dev)).cpu(), torch.arange(37) % 60)


def _train_inputs():
    pr = synth.make_fragment_pair(24, seed=9, overlap=1.0, sigma=0.1)
    f0, f1 = pr["feat_B"][pr["ids_B"]], pr["feat_A"][pr["ids_A"]]
    return f0, f1, np.full((24,), pr["r"], np.int64)


def test_every_backward_call_of_a_real_training_step(_engine_session, tables):
    """One training-mode step of PartI_train + Batch_hard_Rindex_loss on the CUDA kernels; every group-convolution backward call
    inside it (8 calls: 4 layers x 2 branches, real activations and real upstream gradients) is re-derived with float64 torch
    autograd from the SAME inputs: dx and dweight within 2e-5 of their own scale.  (This is the discontinuity-free statement; the
    end-to-end comparison below has to tolerate ReLU-mask flips.)"""
    from yoho_b200 import train as T
    _, _, N = tables
    dev = _engine_session.device
    calls = []
    orig = T._GroupConv.backward

    def spy(ctx, dy):
        out = orig(ctx, dy)
        x, w = ctx.saved_tensors
        calls.append((x.detach().clone(), w.detach().clone(), dy.detach().clone(), [None if o is None else o.detach().clone() for o in out]))
        return out
    T._GroupConv.backward = staticmethod(spy)
    try:
        m = T.PartI_train(None).to(dev)
        m.load_state_dict(synth.to_torch_state_dict(synth.synth_state_dict("PartI", 4)), strict=True)
        m.train()
        f0, f1, ti = _train_inputs()
        out = m({"feats0": torch.from_numpy(f0).to(dev), "feats1": torch.from_numpy(f1).to(dev), "true_idx": torch.from_numpy(ti).to(dev)})
        T.Batch_hard_Rindex_loss()(out).backward()
        torch.cuda.synchronize()
    finally:
        T._GroupConv.backward = staticmethod(orig)
    assert len(calls) == 8
    for x, w, dy, (dx, dw, db) in calls:
        xr, wr = x.double().cpu().requires_grad_(True), w.double().cpu().requires_grad_(True)
        torch.nn.functional.conv2d(O.gather13(xr, N), wr)[:, :, :, 0].backward(dy.double().cpu())
        if dx is not None:
            assert _rel(dx.double().cpu(), xr.grad) <= 2e-5
        assert _rel(dw.double().cpu(), wr.grad) <= 2e-5
        # a conv bias in front of a training-mode BatchNorm has zero gradient: compare on the scale of the terms that cancel
        want_b = dy.double().cpu().sum((0, 2))
        assert float((db.double().cpu() - want_b).abs().max()) <= 2e-5 * float(dy.double().abs().sum((0, 2)).max())


def test_part1_train_step_matches_reference_modules(_engine_session, tables):
    """PartI_train + Batch_hard_Rindex_loss, one forward / backward in TRAINING mode (BatchNorm batch statistics), against the
    unmodified reference modules (utils/network.py:106-138, train/loss_val.py:20-56) in float64: same loss, same rotation
    indices, same gradients.  The reference's checkpoint keys load with strict=True.
    ReLU makes the gradient a discontinuous function of the forward values: a pre-activation within FP32 rounding of zero takes
    the other sub-gradient and changes one row of the preceding layer's weight gradient by a visible amount (measured here: the
    same happens between torch's own FP32 and FP64 runs when a flip occurs), and through the batch-statistics terms of the
    BatchNorm backward a small part of it reaches every entry.  So the end-to-end bar is: median error 1e-4 of the tensor's
    scale, 97 % of the entries within 1e-3, relative L2 error 3e-2; the per-call test above is the tight one (2e-5)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference sources not available")
    from yoho_b200 import train as T
    ref = ref_shim.load_reference()
    import importlib
    loss_mod = importlib.import_module("train.loss_val")
    sd = synth.to_torch_state_dict(synth.synth_state_dict("PartI", 4))
    dev = _engine_session.device
    ours = T.PartI_train(None).to(dev)
    ours.load_state_dict(sd, strict=True)
    ours.train()
    theirs = ref.network.PartI_train(ref.cfgI)
    theirs.load_state_dict(sd, strict=True)
    theirs = theirs.double().cuda()
    theirs.train()
    f0, f1, true_idx = _train_inputs()
    d32 = {"feats0": torch.from_numpy(f0).to(dev), "feats1": torch.from_numpy(f1).to(dev), "true_idx": torch.from_numpy(true_idx).to(dev)}
    d64 = {"feats0": torch.from_numpy(f0).double().cuda(), "feats1": torch.from_numpy(f1).double().cuda(), "true_idx": torch.from_numpy(true_idx).cuda()}
    out = ours(d32)
    loss = T.Batch_hard_Rindex_loss()(out)
    loss.backward()
    rout = theirs(d64)
    rloss = loss_mod.Batch_hard_Rindex_loss(ref.cfgI)(rout)
    rloss.backward()
    assert abs(float(loss.detach()) - float(rloss.detach())) <= 2e-6 * max(1.0, abs(float(rloss.detach())))
    assert torch.equal(out["DR_pre_index"].cpu(), rout["DR_pre_index"].cpu())
    for k in ("feats0_eqv_af_conv", "feats1_eqv_af_conv", "feats0_inv", "feats1_inv"):
        assert float((out[k].detach().double() - rout[k].detach()).abs().max()) <= 1e-5
    rp = dict(theirs.named_parameters())
    for name, p in ours.named_parameters():
        assert p.grad is not None, name
        a, b = p.grad.double().cpu(), rp[name].grad.cpu()
        scale = float(b.abs().max())
        if scale < 1e-12:                           # conv bias in front of a training-mode BatchNorm: exactly zero in exact arithmetic
            assert float(a.abs().max()) <= 1e-6, name
            continue
        d = (a - b).abs()
        assert float(d.median()) <= 1e-4 * scale, name
        assert float((d > 1e-3 * scale).double().mean()) <= 0.03, name
        assert float((a - b).norm() / b.norm()) <= 3e-2, name
