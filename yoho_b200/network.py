"""Drop-in for the inference half of the reference's `utils/network.py`: `PartI_test`, `PartII_test`,
`name2network` — same constructor (`cfg`), same `.cuda()/.eval()/.load_state_dict()` protocol, same call
signature and output dicts (utils/network.py:140-147, 218-278, 282-287), backed by libyoho_b200.so.

The `nn.Module` tree below exists only so that the reference's unmodified checkpoints load with the same key
names and `strict` semantics; it is never executed.  After `load_state_dict` the tensors are
packed once into the C library (BN folded, weights re-laid out) and every forward is a C-ABI call.
There is no torch/CPU fallback: without the CUDA library a forward raises.
"""
import torch
import torch.nn as nn

from .engine import get_engine


def _comb(in_dim, out_dim):
    # parameter container mirroring Comb_Conv.comb_layer (utils/network.py:12-21): indices 0 (BN) and 2 (conv)
    return nn.Sequential(nn.BatchNorm2d(in_dim), nn.ReLU(), nn.Conv2d(in_dim, out_dim, (1, 13), 1))


class _CombConv(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.comb_layer = _comb(in_dim, out_dim)


class _ResidualCombConv(nn.Module):
    def __init__(self, in_dim, middle_dim, out_dim):
        super().__init__()
        self.comb_layer_in = _comb(in_dim, middle_dim)
        self.comb_layer_out = _comb(middle_dim, out_dim)


class _PartINet(nn.Module):
    def __init__(self):
        super().__init__()
        self.Conv_in = nn.Sequential(nn.Conv2d(32, 256, (1, 13), 1))
        self.SO3_Conv_layers = nn.ModuleList([_ResidualCombConv(256, 512, 256)])
        self.Conv_out = _CombConv(256, 32)


class _EngineModule(nn.Module):
    """Common plumbing: engine handle, lazy weight upload, no-op device moves."""
    _part = None

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self._so3 = getattr(cfg, "SO3_related_files", None)
        self._engine = None
        self._token = object()          # identifies THIS module's current weights inside the shared engine

    @property
    def engine(self):
        if self._engine is None:
            self._engine = get_engine(so3_dir=self._so3)
        return self._engine

    def cuda(self, device=None):       # weights live inside the C library; nothing to move
        return self

    def load_state_dict(self, state_dict, strict=True):
        res = super().load_state_dict(state_dict, strict=strict)
        self._token = object()
        return res

    def _ensure_uploaded(self):
        """The engine (one per device) holds one weight set per network: upload ours unless they are the ones in there.  Another
        module with a different checkpoint, or a direct Engine.load_part* call, changes the owner and triggers a re-upload here,
        so two PartI_test instances never silently run with each other's weights (the reference keeps weights per module)."""
        if self.engine.weights_owner[self._part] is not self._token:
            sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            if self._part == 1:
                self.engine.load_part1(sd, owner=self._token)
            else:
                self.engine.load_part2(sd, owner=self._token)


class PartI_test(_EngineModule):
    """utils/network.py:140-147.  forward(group_feat [B,32,60]) -> {'inv': [B,32], 'eqv': [B,32,60]}.
    (Unlike the reference, a batch of one keypoint works: utils/network.py:80-84 squeezes it away.)"""
    _part = 1

    def __init__(self, cfg):
        super().__init__(cfg)
        self.PartI_net = _PartINet()

    def forward(self, group_feat):
        self._ensure_uploaded()
        x = group_feat
        if x.dim() == 2:
            x = x[None]
        out = self.engine.part1(x, want_inv=True, want_desc=False)
        return {"inv": out["inv"], "eqv": out["eqv"]}


class PartII_test(_EngineModule):
    """utils/network.py:218-278.  forward({'before_eqv0','before_eqv1','after_eqv0','after_eqv1','pre_idx'})
    -> {'quaternion_pre': [b,4] (w,x,y,z), 'pre_idxs': [b]}.  Inputs are NOT modified (the reference permutes
    'before_eqv0'/'after_eqv0' in place, utils/network.py:266-268)."""
    _part = 2

    def __init__(self, cfg):
        super().__init__(cfg)
        self.Conv_init = _CombConv(32 * 4, 256)
        self.PartII_SO3_Conv_layers = nn.ModuleList([_ResidualCombConv(256, 512, 256)])
        self.PartII_To_R_FC = nn.Sequential(
            nn.Conv2d(256, 512, 1, 1), nn.BatchNorm2d(512), nn.ReLU(),
            nn.Conv2d(512, 128, 1, 1), nn.BatchNorm2d(128), nn.ReLU(),
            nn.Conv2d(128, 4, 1, 1))

    def forward(self, data):
        self._ensure_uploaded()
        pre = data["pre_idx"]
        # batch_create stores fragment id1 in the "*_eqv0" slots and id0 in "*_eqv1" (tests/extractor.py:132-137)
        quat, _ = self.engine.part2(fcgf0=data["before_eqv1"], fcgf1=data["before_eqv0"],
                                    yoho0=data["after_eqv1"], yoho1=data["after_eqv0"], pre_idx=pre)
        return {"quaternion_pre": quat, "pre_idxs": pre}


def _train_only(name):
    class _Unavailable(nn.Module):
        def __init__(self, cfg):
            super().__init__()
            raise NotImplementedError(f"{name}: the PartII training network is outside the B200 hot path "
                                      "(SURVEY.md §2.1, 'Training' row)")
    _Unavailable.__name__ = name
    return _Unavailable


def _part1_train(cfg):
    from .train import PartI_train          # training-time twin (csrc/train.cu)
    return PartI_train(cfg)


name2network = {
    "PartI_train": _part1_train,
    "PartI_test": PartI_test,
    "PartII_train": _train_only("PartII_train"),
    "PartII_test": PartII_test,
}
