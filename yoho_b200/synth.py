"""Seeded synthetic inputs and weights.  Pure numpy, bit-reproducible on any box
(legacy `np.random.RandomState` streams are stable across numpy versions), so the goldens generated
in the authoring container and the tensors regenerated on the GPU box are identical.

* `synth_state_dict(part, seed)` — a state-dict with the reference's key names and shapes
  (SURVEY.md §8a "Checkpoint key names"), variance-preserving random conv weights and BatchNorm
  statistics drawn analytically (no calibration pass, so no dependence on a BLAS build).
* `make_fragment_pair(K, seed, ...)` — FCGF-like unit-norm group features [K,32,60] for two fragments
  with a planted group rotation r, residual rotation, translation and overlap (BASELINE.json configs).
"""
import numpy as np
from . import group as _group


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def _conv(rs, o, c, taps, gain):
    std = np.sqrt(gain / (c * taps))
    w = (rs.standard_normal((o, c, 1, taps)) * std).astype(np.float32)
    b = (rs.standard_normal((o,)) * 0.02).astype(np.float32)
    return w, b


def _bn(rs, c, mean_scale=0.3, var_lo=0.5, var_hi=2.0):
    return {
        "weight": rs.uniform(0.5, 1.5, (c,)).astype(np.float32),
        "bias": (rs.standard_normal((c,)) * 0.3).astype(np.float32),
        "running_mean": (rs.standard_normal((c,)) * mean_scale).astype(np.float32),
        "running_var": rs.uniform(var_lo, var_hi, (c,)).astype(np.float32),
        "num_batches_tracked": np.array(1000, dtype=np.int64),
    }


def _put_bn(sd, prefix, bn):
    for k, v in bn.items():
        sd[f"{prefix}.{k}"] = v


def synth_state_dict(part: str, seed: int = 0):
    """part in {'PartI','PartII'} -> dict name -> np.ndarray with the reference's key names."""
    rs = np.random.RandomState(1000 + seed + (0 if part == "PartI" else 500))
    sd = {}
    if part == "PartI":
        p = "PartI_net."
        w, b = _conv(rs, 256, 32, 13, 32.0)            # inputs have variance 1/32 per channel
        sd[p + "Conv_in.0.weight"], sd[p + "Conv_in.0.bias"] = w, b
        blk = p + "SO3_Conv_layers.0."
        _put_bn(sd, blk + "comb_layer_in.0", _bn(rs, 256))
        w, b = _conv(rs, 512, 256, 13, 2.0)
        sd[blk + "comb_layer_in.2.weight"], sd[blk + "comb_layer_in.2.bias"] = w, b
        _put_bn(sd, blk + "comb_layer_out.0", _bn(rs, 512))
        w, b = _conv(rs, 256, 512, 13, 2.0)
        sd[blk + "comb_layer_out.2.weight"], sd[blk + "comb_layer_out.2.bias"] = w, b
        _put_bn(sd, p + "Conv_out.comb_layer.0", _bn(rs, 256, var_lo=1.0, var_hi=3.0))
        w, b = _conv(rs, 32, 256, 13, 0.05)            # keeps the residual branch ~ the input scale
        sd[p + "Conv_out.comb_layer.2.weight"], sd[p + "Conv_out.comb_layer.2.bias"] = w, b
    elif part == "PartII":
        _put_bn(sd, "Conv_init.comb_layer.0", _bn(rs, 128, mean_scale=0.02, var_lo=0.02, var_hi=0.05))
        w, b = _conv(rs, 256, 128, 13, 2.0)
        sd["Conv_init.comb_layer.2.weight"], sd["Conv_init.comb_layer.2.bias"] = w, b
        blk = "PartII_SO3_Conv_layers.0."
        _put_bn(sd, blk + "comb_layer_in.0", _bn(rs, 256))
        w, b = _conv(rs, 512, 256, 13, 2.0)
        sd[blk + "comb_layer_in.2.weight"], sd[blk + "comb_layer_in.2.bias"] = w, b
        _put_bn(sd, blk + "comb_layer_out.0", _bn(rs, 512))
        w, b = _conv(rs, 256, 512, 13, 2.0)
        sd[blk + "comb_layer_out.2.weight"], sd[blk + "comb_layer_out.2.bias"] = w, b
        fc = "PartII_To_R_FC."
        w, b = _conv(rs, 512, 256, 1, 1.0)
        sd[fc + "0.weight"], sd[fc + "0.bias"] = w, b
        _put_bn(sd, fc + "1", _bn(rs, 512, var_lo=1.0, var_hi=3.0))
        w, b = _conv(rs, 128, 512, 1, 2.0)
        sd[fc + "3.weight"], sd[fc + "3.bias"] = w, b
        _put_bn(sd, fc + "4", _bn(rs, 128))
        w, b = _conv(rs, 4, 128, 1, 2.0)
        sd[fc + "6.weight"] = w
        sd[fc + "6.bias"] = (b + np.array([0.9, 0.0, 0.0, 0.0], np.float32)).astype(np.float32)
    else:
        raise ValueError(part)
    return sd


def to_torch_state_dict(sd):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}


# --------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------
def _unit(x, axis):
    n = np.sqrt((x * x).sum(axis=axis, keepdims=True))
    return x / np.maximum(n, 1e-12)


def _small_rotation(rs, max_deg):
    axis = rs.standard_normal(3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rs.uniform(0, max_deg))
    kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * kx + (1 - np.cos(ang)) * (kx @ kx)


def make_fragment(K, seed):
    """One fragment: FCGF-like group feature [K,32,60] f32 (unit-norm over the 32 channels for every
    (keypoint, g), as the backbone's normalize_feature=True produces) and keypoints [K,3] f64."""
    rs = np.random.RandomState(seed)
    feat = _unit(rs.standard_normal((K, 32, 60)), 1).astype(np.float32)
    kps = rs.uniform(0.0, 3.0, (K, 3))
    return feat, kps


def make_fragment_pair(K, seed=0, overlap=0.5, sigma=0.05, max_residual_deg=15.0, kp_noise=0.01,
                       so3_dir=None):
    """Pair (A,B) with planted transform  pts_A = R_gt pts_B + t_gt  on an overlapping subset.

    Returns dict: feat_A, feat_B [K,32,60] f32; kps_A, kps_B [K,3] f64; r (planted group index, such
    that F_A[:, :, g] ~= F_B[:, :, P[r][g]], i.e. pts_A ~= R_r pts_B); R_gt, t_gt; subset ids.
    """
    gt = _group.load(so3_dir)
    rs = np.random.RandomState(seed)
    feat_B = _unit(rs.standard_normal((K, 32, 60)), 1).astype(np.float32)
    kps_B = rs.uniform(0.0, 3.0, (K, 3))
    r = int(rs.randint(0, 60))
    dR = _small_rotation(rs, max_residual_deg)
    R_gt = dR @ gt.R[r]
    t_gt = rs.uniform(-1.0, 1.0, 3)
    n_ov = int(round(overlap * K))
    ids_B = rs.permutation(K)[:n_ov]
    ids_A = rs.permutation(K)[:n_ov]
    feat_A = _unit(rs.standard_normal((K, 32, 60)), 1).astype(np.float32)
    kps_A = rs.uniform(-2.0, 5.0, (K, 3))
    fa = feat_B[ids_B][:, :, gt.P[r]] + sigma * rs.standard_normal((n_ov, 32, 60))
    feat_A[ids_A] = _unit(fa, 1).astype(np.float32)
    kps_A[ids_A] = kps_B[ids_B] @ R_gt.T + t_gt + kp_noise * rs.standard_normal((n_ov, 3))
    return dict(feat_A=feat_A, feat_B=feat_B, kps_A=kps_A, kps_B=kps_B, r=r, R_gt=R_gt, t_gt=t_gt,
                ids_A=ids_A, ids_B=ids_B)


def make_scene(n_fragments, K, seed=0, cluster=8, overlap_lo=0.55, overlap_hi=0.95, sigma=0.05, max_residual_deg=15.0,
               kp_noise=0.01, so3_dir=None):
    """A 3DMatch-shaped scene (BASELINE.json configs 3-4): clusters of `cluster` fragments cut from one base cloud, every
    fragment carrying its own rigid motion (group element r_j times a small residual rotation, translation t_j) and its own
    subset S_j (fraction rho_j ~ U[overlap_lo, overlap_hi]) of the base points; the rest of each fragment is unrelated clutter.
    Two fragments of a cluster overlap on S_i & S_j (fraction ~ rho_i rho_j), so all C(cluster, 2) pairs of a cluster are
    genuine registration problems: 3.5 pairs per fragment at cluster = 8 (3DMatch: 1623 / 433 = 3.75).

    Returns (fragments, pair_ids, gt): fragments {id: (feat [K,32,60] f32, kps [K,3] f64)}, pair_ids [(i, j)], gt {(i, j): (R, t)}
    with pts_i = R pts_j + t on the overlap.
    """
    gt = _group.load(so3_dir)
    rs = np.random.RandomState(seed)
    fragments, pair_ids, gts = {}, [], {}
    fid = 0
    while fid < n_fragments:
        n_c = min(cluster, n_fragments - fid)
        base_f = _unit(rs.standard_normal((K, 32, 60)), 1)
        base_k = rs.uniform(0.0, 3.0, (K, 3))
        motions = []
        for j in range(n_c):
            r = int(rs.randint(0, 60))
            R = _small_rotation(rs, max_residual_deg) @ gt.R[r]
            t = rs.uniform(-1.0, 1.0, 3)
            rho = rs.uniform(overlap_lo, overlap_hi)
            n_ov = int(round(rho * K))
            src = rs.permutation(K)[:n_ov]                 # base points seen by this fragment
            dst = rs.permutation(K)[:n_ov]                 # where they sit in the fragment
            feat = _unit(rs.standard_normal((K, 32, 60)), 1).astype(np.float32)
            kps = rs.uniform(-2.0, 5.0, (K, 3))
            # F(R_r pc)[:, :, g] = F(pc)[:, :, P[r][g]]  
            feat[dst] = _unit(base_f[src][:, :, gt.P[r]] + sigma * rs.standard_normal((n_ov, 32, 60)), 1).astype(np.float32)
            kps[dst] = base_k[src] @ R.T + t + kp_noise * rs.standard_normal((n_ov, 3))
            fragments[fid + j] = (feat, kps)
            motions.append((R, t))
        for i in range(n_c):
            for j in range(i + 1, n_c):
                Ri, ti = motions[i]
                Rj, tj = motions[j]
                Rij = Ri @ Rj.T                            # pts_i = Ri p + ti, pts_j = Rj p + tj  ->  pts_i = Rij pts_j + (ti - Rij tj)
                pair_ids.append((fid + i, fid + j))
                gts[(fid + i, fid + j)] = (Rij, ti - Rij @ tj)
        fid += n_c
    return fragments, pair_ids, gts


# --------------------------------------------------------------------------------------------
# dataset-shaped scene sets (BASELINE.json configs 3-4), generated fragment by fragment on the device
# --------------------------------------------------------------------------------------------
THREEDMATCH_SCENE_SIZES = [60, 60, 60, 55, 57, 37, 66, 38]      # fragments per 3DMatch test scene (utils/dataset.py:167): 433


class SceneSet:
    """Host-side description of a multi-scene set: which scene a fragment belongs to, its planted motion and visible fraction,
    the pair list (pairs never cross scenes) and the ground truth of every pair.  The tensors themselves are produced per
    fragment by `fragment_torch` — on whichever rank needs them, bit-identically (seeded device generators) — so a 433-fragment
    set (16.6 GB of group features) never exists on the host."""

    def __init__(self, sizes, n_pairs, K, seed=0, overlap_lo=0.55, overlap_hi=0.95, sigma=0.05, max_residual_deg=15.0,
                 kp_noise=0.01, so3_dir=None):
        gt = _group.load(so3_dir)
        rs = np.random.RandomState(seed)
        self.K, self.seed, self.sigma, self.kp_noise = int(K), int(seed), float(sigma), float(kp_noise)
        self.P = gt.P
        self.frag_ids, self.scene_of, self.motion, self.rho, self.r = [], {}, {}, {}, {}
        pairs = []
        base = 0
        for s, n in enumerate(sizes):
            ids = list(range(base, base + n))
            for f in ids:
                r = int(rs.randint(0, 60))
                self.r[f] = r
                self.motion[f] = (_small_rotation(rs, max_residual_deg) @ gt.R[r], rs.uniform(-1.0, 1.0, 3))
                self.rho[f] = float(rs.uniform(overlap_lo, overlap_hi))
                self.scene_of[f] = s
            self.frag_ids += ids
            # a fragment overlaps its temporal neighbours ...
            for i in range(n):
                for d in (1, 2, 3):
                    if i + d < n:
                        pairs.append((ids[i], ids[i + d]))
            base += n
        # ... plus loop closures, until the set has the requested number of pairs (3DMatch: 1623, 3DLoMatch: 1781)
        have = set(pairs)
        guard = 0
        while len(pairs) < n_pairs and guard < 100 * n_pairs:
            guard += 1
            s = int(rs.randint(len(sizes)))
            off = int(sum(sizes[:s]))
            i, j = sorted(int(v) for v in rs.randint(0, sizes[s], 2))
            if j - i > 3 and (off + i, off + j) not in have:
                have.add((off + i, off + j))
                pairs.append((off + i, off + j))
        self.pair_ids = sorted(pairs[:n_pairs])
        self.gt = {}
        for (i, j) in self.pair_ids:
            Ri, ti = self.motion[i]
            Rj, tj = self.motion[j]
            Rij = Ri @ Rj.T                      # pts_i = Ri p + ti, pts_j = Rj p + tj  ->  pts_i = Rij pts_j + (ti - Rij tj)
            self.gt[(i, j)] = (Rij, ti - Rij @ tj)
        self._base = {}

    def _scene_base(self, s, device):
        import torch
        key = (s, str(device))
        if key not in self._base:
            if len(self._base) >= 2:
                self._base.pop(next(iter(self._base)))
            g = torch.Generator(device=device)
            g.manual_seed(self.seed * 1000003 + 7919 * s + 1)
            f = torch.randn((self.K, 32, 60), generator=g, device=device)
            f = f / f.norm(dim=1, keepdim=True).clamp_min(1e-12)
            k = torch.rand((self.K, 3), generator=g, device=device, dtype=torch.float64) * 3.0
            self._base[key] = (f, k)
        return self._base[key]

    def fragment_torch(self, fid, device):
        """-> (feat [K,32,60] f32, kps [K,3] f64) on `device`; the same values on every rank."""
        import torch
        K = self.K
        base_f, base_k = self._scene_base(self.scene_of[fid], device)
        g = torch.Generator(device=device)
        g.manual_seed(self.seed * 1000003 + 104729 * (fid + 1))
        n_ov = int(round(self.rho[fid] * K))
        src = torch.randperm(K, generator=g, device=device)[:n_ov]
        dst = torch.randperm(K, generator=g, device=device)[:n_ov]
        feat = torch.randn((K, 32, 60), generator=g, device=device)
        feat = feat / feat.norm(dim=1, keepdim=True).clamp_min(1e-12)
        kps = torch.rand((K, 3), generator=g, device=device, dtype=torch.float64) * 7.0 - 2.0
        perm = torch.as_tensor(np.asarray(self.P[self.r[fid]], np.int64), device=device)
        # F(R_r pc)[:, :, g] = F(pc)[:, :, P[r][g]]  
        v = base_f[src][:, :, perm] + self.sigma * torch.randn((n_ov, 32, 60), generator=g, device=device)
        feat[dst] = v / v.norm(dim=1, keepdim=True).clamp_min(1e-12)
        R, t = self.motion[fid]
        Rt = torch.as_tensor(R, device=device, dtype=torch.float64)
        tt = torch.as_tensor(t, device=device, dtype=torch.float64)
        kps[dst] = base_k[src] @ Rt.T + tt + self.kp_noise * torch.randn((n_ov, 3), generator=g, device=device, dtype=torch.float64)
        return feat.contiguous(), kps.contiguous()

    def success(self, pair, T, max_rot_deg=5.0, max_trans=0.3):
        R, t = self.gt[pair]
        cosang = np.clip((np.trace(T[:, :3].T @ R) - 1) / 2, -1, 1)
        return bool(np.degrees(np.arccos(cosang)) < max_rot_deg and np.linalg.norm(T[:, 3] - t) < max_trans)
