"""Training-time twins of the hot kernels: differentiable wrappers over the C ABI with the reference's names.

  * `group_conv(x, weight, bias)`       Comb_Conv's gather + Conv2d(C,O,(1,13)) (utils/network.py:12-21,46-52,80-84) as ONE
                                         autograd Function: forward, backward-data and backward-weight are CUDA kernels
                                         (csrc/train.cu); the 13x gathered tensor is never materialised.
  * `rot_correlation(des1, des2)`       cor[b,a] = sum_{f,g} des1[b,f,P[a][g]] des2[b,f,g] with its gradient — the score of
                                         `Batch_hard_Rindex_loss.eqvloss` (train/loss_val.py:27-31).
  * `Des2DR(des1, des2)`                argmax_a of that correlation: `PartI_train.Des2DR` (utils/network.py:115-118).
  * `Comb_Conv`, `Residual_Comb_Conv`, `PartI_network`, `PartI_train`, `Batch_hard_Rindex_loss`
                                         torch modules with the reference's parameter names (so its checkpoints load with
                                         strict=True and an optimiser sees the same tensors); BatchNorm / ReLU / the loss's
                                         softmax arithmetic stay torch ops — only the path's own operators are replaced.

There is no CPU fallback: the Functions raise without a CUDA device, like the rest of the package.
"""
import torch
import torch.nn as nn

from . import _lib
from .engine import get_engine, _ptr, _stream


def _eng(t):
    if not t.is_cuda:
        raise _lib.YohoError("yoho_b200.train needs CUDA tensors (there is no CPU fallback)")
    return get_engine(t.device.index)


class _GroupConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        e = _eng(x)
        x = x.contiguous().float()
        w = weight.contiguous().float()
        B, C, G = x.shape
        O = w.shape[0]
        assert G == 60 and tuple(w.shape) == (O, C, 1, 13), "x [B,C,60], weight [O,C,1,13]"
        b = bias.contiguous().float() if bias is not None else None
        y = torch.empty((B, O, 60), device=x.device, dtype=torch.float32)
        _lib.check(e.lib.yoho_gconv_train_forward(e.h, _ptr(x), _ptr(w), _ptr(b), B, C, O, _ptr(y), _stream()))
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        e = _eng(x)
        dy = dy.contiguous().float()
        B, C, _ = x.shape
        O = w.shape[0]
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        dx = torch.empty_like(x) if need_x else None
        dw = torch.empty_like(w) if (need_w or need_b) else None
        db = torch.empty((O,), device=x.device, dtype=torch.float32) if need_b else None
        if B == 0:
            return (torch.zeros_like(x) if need_x else None, torch.zeros_like(w) if need_w else None,
                    torch.zeros((O,), device=x.device) if need_b else None)
        _lib.check(e.lib.yoho_gconv_train_backward(e.h, _ptr(x), _ptr(w), _ptr(dy), B, C, O, _ptr(dx), _ptr(dw), _ptr(db), _stream()))
        return dx, (dw if need_w else None), db


def group_conv(x, weight, bias=None):
    """x [B,C,60] -> [B,O,60]; weight [O,C,1,13] (nn.Conv2d(C,O,(1,13)).weight), bias [O] or None."""
    return _GroupConv.apply(x, weight, bias)


class _RotCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, des1, des2):
        e = _eng(des1)
        d1, d2 = des1.contiguous().float(), des2.contiguous().float()
        _, cor = e.rot_argmax(d1, d2, want_cor=True)
        ctx.save_for_backward(d1, d2)
        return cor

    @staticmethod
    def backward(ctx, gcor):
        d1, d2 = ctx.saved_tensors
        e = _eng(d1)
        M = d1.shape[0]
        g1 = torch.empty_like(d1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(d2) if ctx.needs_input_grad[1] else None
        if M and (g1 is not None or g2 is not None):
            _lib.check(e.lib.yoho_rot_correlation_backward(e.h, _ptr(d1), _ptr(d2), _ptr(gcor.contiguous().float()), M, _ptr(g1),
                                                           _ptr(g2), _stream()))
        return g1, g2


def rot_correlation(des1, des2):
    """des1, des2 [B,32,60] -> cor [B,60] (differentiable)."""
    return _RotCorrelation.apply(des1, des2)


def Des2DR(des1, des2):
    """PartI_train.Des2DR (utils/network.py:115-118): before-rotation / after-rotation descriptors -> rotation index [B]."""
    e = _eng(des1)
    return e.rot_argmax(des1.detach(), des2.detach())


# ---- modules with the reference's parameter names ------------------------------------------------------------------
class _GConv2d(nn.Conv2d):
    """nn.Conv2d(C,O,(1,13)) parameters (same names / shapes / initialisation as the reference's), applied to the UN-gathered
    [B,C,60] tensor through `group_conv`."""

    def forward(self, x):
        return group_conv(x, self.weight, self.bias)


class Comb_Conv(nn.Module):
    """utils/network.py:12-21: BN -> ReLU -> Conv2d(in,out,(1,13)) on the gathered tensor.  BatchNorm2d / ReLU are per-channel /
    point-wise and the gather only permutes the group axis, so they are applied before it on [B,C,60,1];
    in TRAINING mode BatchNorm2d's batch statistics over the 13x gathered tensor equal those over the un-gathered one because
    every tap column of the neighbour table is a permutation of the group (each element appears exactly 13 times)."""

    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.comb_layer = nn.Sequential(nn.BatchNorm2d(in_dim), nn.ReLU(), _GConv2d(in_dim, out_dim, (1, 13), 1))

    def forward(self, x):                       # x [B,C,60]
        h = self.comb_layer[1](self.comb_layer[0](x.unsqueeze(-1))).squeeze(-1)
        return self.comb_layer[2](h)


class Residual_Comb_Conv(nn.Module):
    """utils/network.py:23-65."""

    def __init__(self, in_dim, middle_dim, out_dim):
        super().__init__()
        self.comb_layer_in = nn.Sequential(nn.BatchNorm2d(in_dim), nn.ReLU(), _GConv2d(in_dim, middle_dim, (1, 13), 1))
        self.comb_layer_out = nn.Sequential(nn.BatchNorm2d(middle_dim), nn.ReLU(), _GConv2d(middle_dim, out_dim, (1, 13), 1))
        self.short_cut = in_dim != out_dim
        if self.short_cut:
            self.short_cut_layer = nn.Sequential(nn.BatchNorm2d(in_dim), nn.ReLU(), _GConv2d(in_dim, out_dim, (1, 13), 1))

    @staticmethod
    def _seq(seq, x):
        return seq[2](seq[1](seq[0](x.unsqueeze(-1))).squeeze(-1))

    def forward(self, x):
        y = self._seq(self.comb_layer_out, self._seq(self.comb_layer_in, x))
        if self.short_cut:
            return y + self._seq(self.short_cut_layer, x)
        return y + x


class PartI_network(nn.Module):
    """utils/network.py:67-105 (training-capable)."""

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = cfg
        self.Conv_in = nn.Sequential(_GConv2d(32, 256, (1, 13), 1))
        self.SO3_Conv_layers = nn.ModuleList([Residual_Comb_Conv(256, 512, 256)])
        self.Conv_out = Comb_Conv(256, 32)

    def forward(self, feats):
        feats_eqv = feats.reshape(-1, 32, 60)
        x = self.Conv_in[0](feats_eqv)
        for layer in self.SO3_Conv_layers:
            x = layer(x)
        x = self.Conv_out(x)
        feats_eqv = x + feats_eqv                                                                   # :98
        feats_inv = torch.mean(feats_eqv, dim=-1)                                                   # :99
        feats_eqv = feats_eqv / torch.clamp_min(torch.norm(feats_eqv, dim=1, keepdim=True), min=1e-4)
        feats_inv = feats_inv / torch.clamp_min(torch.norm(feats_inv, dim=1, keepdim=True), min=1e-4)
        return {'inv': feats_inv, 'eqv': feats_eqv}


class PartI_train(nn.Module):
    """utils/network.py:106-138."""

    def __init__(self, cfg=None):
        super().__init__()
        self.cfg = cfg
        self.PartI_net = PartI_network(cfg)

    def Des2DR(self, Des1, Des2):
        return Des2DR(Des1, Des2)

    def forward(self, data):
        feats0 = data['feats0'].reshape(-1, 32, 60)
        feats1 = data['feats1'].reshape(-1, 32, 60)
        true_idxs = data['true_idx'].reshape(-1)
        yoho_0 = self.PartI_net(feats0)
        yoho_1 = self.PartI_net(feats1)
        pre_idxs = self.Des2DR(yoho_0['eqv'], yoho_1['eqv'])
        part1_ability = torch.mean((pre_idxs == true_idxs).type(torch.float32))
        return {'feats0_eqv_bf_conv': feats0, 'feats1_eqv_bf_conv': feats1,
                'feats0_eqv_af_conv': yoho_0['eqv'], 'feats1_eqv_af_conv': yoho_1['eqv'],
                'feats0_inv': yoho_0['inv'], 'feats1_inv': yoho_1['inv'],
                'DR_pre_ability': part1_ability, 'DR_true_index': true_idxs, 'DR_pre_index': pre_idxs}


class Batch_hard_Rindex_loss:
    """train/loss_val.py:20-56: 5 * batch-hard triplet loss on the invariant descriptors + cross-entropy of the rotation score."""

    def __init__(self, cfg=None):
        self.keys = ['triplet_ranking_Rindex_loss']
        self.class_loss = torch.nn.CrossEntropyLoss()

    def eqvloss(self, eqvfeat0, eqvfeat1):
        return rot_correlation(eqvfeat0, eqvfeat1)

    def __call__(self, data_pr):
        Index = data_pr['DR_true_index'].type(torch.int64)
        feats0 = data_pr['feats0_inv']
        feats1 = data_pr['feats1_inv']
        B, L = feats1.shape
        q_vec = feats0.contiguous().view(B, 1, L)
        ans_vecs = feats1.contiguous().view(1, B, L)
        dist = ((q_vec - ans_vecs) ** 2).sum(-1)
        dist = torch.nn.functional.log_softmax(dist, 1)
        loss_true = torch.diag(dist)
        loss_false = torch.min(dist + torch.eye(B, device=dist.device), dim=1)[0]
        loss = torch.mean(torch.clamp_min(loss_true - loss_false + 0.3, 0))
        score = self.eqvloss(data_pr['feats0_eqv_af_conv'], data_pr['feats1_eqv_af_conv'])
        eqv_loss = self.class_loss(score, Index)
        return 5 * loss + eqv_loss
