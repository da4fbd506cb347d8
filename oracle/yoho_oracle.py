"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; nothing under `yoho_b200/` does.  The product path has no CPU fallback.

Every function cites the reference file:line it follows.  Floating-point stages are restated with the
same torch CPU operators the reference calls (conv2d / batch_norm / einsum / min), so in FP32 they are
the reference's own arithmetic; each also runs in FP64 (`dtype=torch.float64`) as the arbiter for
tolerance disputes and near-tie decisions.  Pinning status: PINNED — checked against the
unmodified reference imported under `oracle/ref_shim.py` (tests/test_oracle_vs_reference.py, runs when
/root/reference exists) and against the committed goldens in tests/golden/ (generated from the reference
by tests/golden/make_golden.py).

The estimator part (E1-E5) is restated in C (`oracle/estimator_oracle.c`, loaded through ctypes by
`oracle/estimator_oracle.py`) because bit-exactness with the CUDA kernels needs a fixed FP64 operation
order that numpy/LAPACK does not define.
"""
import os
import numpy as np
import torch
import torch.nn.functional as F

G, TAPS = 60, 13
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "yoho_b200", "data", "group_related")


def load_tables(so3_dir=None):
    d = so3_dir or _DATA
    R = np.load(os.path.join(d, "Rotation.npy")).astype(np.float64)
    P = np.load(os.path.join(d, "60_60.npy")).astype(np.int64)
    N = np.load(os.path.join(d, "Nei_Index_in_SO3_ordered_13.npy")).astype(np.int64)
    return R, P, N


def _t(sd, key, dtype):
    v = sd[key]
    if isinstance(v, np.ndarray):
        v = torch.from_numpy(np.ascontiguousarray(v))
    return v.to(dtype)


# ---------------------------------------------------------------------------------------------
# A1-A4: group convolution pieces
# ---------------------------------------------------------------------------------------------
def gather13(x, N):
    """utils/network.py:46-52, 80-84 (`data_process`): x[B,C,60] -> [B,C,60,13] via x[:,:,N.flat]."""
    B, C, _ = x.shape
    idx = torch.from_numpy(N.reshape(-1))
    return x[:, :, idx].reshape(B, C, G, TAPS)


def bn_relu(x, sd, prefix, dtype):
    """Eval-mode BatchNorm2d + ReLU (utils/network.py:16-17,28-29,33-34), eps=1e-5.
    BN/ReLU are per-channel / point-wise, so applying them before or after `gather13` gives
    bit-identical values; `faithful_cost` callers apply them after, like the reference."""
    y = F.batch_norm(x, _t(sd, prefix + ".running_mean", dtype), _t(sd, prefix + ".running_var", dtype),
                     _t(sd, prefix + ".weight", dtype), _t(sd, prefix + ".bias", dtype), False, 0.0, 1e-5)
    return F.relu(y)


def gconv(x, sd, prefix, N, dtype, bn_prefix=None, faithful_cost=False):
    """[BN->ReLU->] Conv2d(C,O,(1,13)) on the gathered tensor (utils/network.py:12-21,18,30,35,76).
    x [B,C,60] -> [B,O,60]."""
    w = _t(sd, prefix + ".weight", dtype)
    b = _t(sd, prefix + ".bias", dtype)
    if bn_prefix is not None and not faithful_cost:
        x = bn_relu(x[:, :, :, None], sd, bn_prefix, dtype)[:, :, :, 0]
    xg = gather13(x, N)
    if bn_prefix is not None and faithful_cost:
        xg = bn_relu(xg, sd, bn_prefix, dtype)
    return F.conv2d(xg, w, b)[:, :, :, 0]


# ---------------------------------------------------------------------------------------------
# A5-A6: PartI
# ---------------------------------------------------------------------------------------------
def part1_forward(x, sd, N, dtype=torch.float32, faithful_cost=False):
    """PartI_test / PartI_network.forward (utils/network.py:86-105,140-147).
    x: [B,32,60] (numpy or tensor).  Returns dict eqv [B,32,60], inv [B,32] as torch tensors."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    x = x.to(dtype)
    p = "PartI_net."
    blk = p + "SO3_Conv_layers.0."
    with torch.no_grad():
        y1 = gconv(x, sd, p + "Conv_in.0", N, dtype, None, faithful_cost)                                   # :87-88
        y2 = gconv(y1, sd, blk + "comb_layer_in.2", N, dtype, blk + "comb_layer_in.0", faithful_cost)       # :55-56
        y3 = gconv(y2, sd, blk + "comb_layer_out.2", N, dtype, blk + "comb_layer_out.0", faithful_cost) + y1  # :57-65
        y4 = gconv(y3, sd, p + "Conv_out.comb_layer.2", N, dtype, p + "Conv_out.comb_layer.0", faithful_cost)  # :91-92
        e = y4 + x                                                                                           # :98
        inv = torch.mean(e, dim=-1)                                                                          # :99
        eqv = e / torch.clamp_min(torch.norm(e, dim=1, keepdim=True), min=1e-4)                              # :102
        inv = inv / torch.clamp_min(torch.norm(inv, dim=1, keepdim=True), min=1e-4)                          # :103
    return {"eqv": eqv, "inv": inv}


def part1_extract(x, sd, N, batch=900, dtype=torch.float32, faithful_cost=False):
    """extractor_PartI.Extract batching (tests/extractor.py:51-59): keeps eqv only."""
    outs = []
    for s in range(0, x.shape[0], batch):
        outs.append(part1_forward(x[s:s + batch], sd, N, dtype, faithful_cost)["eqv"])
    return torch.cat(outs, 0)


def matcher_descriptor(eqv):
    """tests/matcher.py:35-36: mean over the 60 group elements of the saved eqv, NOT re-normalised.
    numpy float32 mean, exactly as the reference computes it."""
    e = eqv.numpy() if isinstance(eqv, torch.Tensor) else eqv
    return np.mean(e, axis=-1).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# B1-B2: matcher
# ---------------------------------------------------------------------------------------------
def nn1(source, target, dtype=torch.float32, chunk=500):
    """modified_knn_matcher.find_nn_gpu / pdist with dist_type='L2', nn_max_n=500
    (utils/knn_search.py:17-24,26-66,138-154): for every source row the (distance, index) of the nearest
    target row; dist = sqrt(sum((a-b)^2)+1e-7); torch.min -> first minimal index."""
    s = torch.as_tensor(source).to(dtype)
    t = torch.as_tensor(target).to(dtype)
    ds, ids = [], []
    for i in range(0, s.shape[0], chunk):
        d2 = torch.sum((s[i:i + chunk].unsqueeze(1) - t.unsqueeze(0)).pow(2), 2)
        d, ind = torch.sqrt(d2 + 1e-7).min(dim=1)
        ds.append(d)
        ids.append(ind)
    return torch.cat(ds), torch.cat(ids)


def mutual_matches(desc0, desc1, dtype=torch.float32):
    """matcher_dual.match (tests/matcher.py:37-48): mutual 1-NN pairs, ascending in fragment-0 index.
    Returns int64 [M,2] and the two argmin arrays."""
    _, a01 = nn1(desc0, desc1, dtype)      # KNN(feats1, feats0): for each 0-point its NN in 1
    _, a10 = nn1(desc1, desc0, dtype)
    a01 = a01.numpy()
    a10 = a10.numpy()
    i0 = np.arange(a01.shape[0])
    keep = a10[a01] == i0
    pps = np.stack([i0[keep], a01[keep]], 1).astype(np.int64).reshape(-1, 2)
    return pps, a01, a10


def nn1_margins(source, target):
    """FP64 arbiter: best and second-best squared distance per source row (near-tie flagging)."""
    s = torch.as_tensor(source).double()
    t = torch.as_tensor(target).double()
    best, second, idx = [], [], []
    for i in range(0, s.shape[0], 500):
        d2 = torch.sum((s[i:i + 500].unsqueeze(1) - t.unsqueeze(0)).pow(2), 2)
        v, ind = torch.topk(d2, 2, dim=1, largest=False)
        best.append(v[:, 0]); second.append(v[:, 1]); idx.append(ind[:, 0])
    return torch.cat(best).numpy(), torch.cat(second).numpy(), torch.cat(idx).numpy()


# ---------------------------------------------------------------------------------------------
# C1: rotation-correlation argmax
# ---------------------------------------------------------------------------------------------
def rot_correlation(des1, des2, P, dtype=torch.float32):
    """extractor_dr_index.Batch_Des2R_torch (tests/extractor.py:74-78):
    cor[b,a] = sum_{f,g} des1[b,f,P[a][g]] * des2[b,f,g]."""
    d1 = torch.as_tensor(des1).to(dtype)
    d2 = torch.as_tensor(des2).to(dtype)
    B, Fd, _ = d1.shape
    idx = torch.from_numpy(P.reshape(-1))
    cors = []
    for s in range(0, B, 256):
        x = d1[s:s + 256][:, :, idx].reshape(-1, Fd, G, G)
        cors.append(torch.einsum('bfag,bfg->ba', x, d2[s:s + 256]))
    return torch.cat(cors) if cors else torch.zeros((0, G), dtype=dtype)


def rot_argmax(des1, des2, P, dtype=torch.float32):
    cor = rot_correlation(des1, des2, P, dtype)
    return torch.argmax(cor, dim=1).numpy().astype(np.int64), cor


# ---------------------------------------------------------------------------------------------
# D1-D3: PartII
# ---------------------------------------------------------------------------------------------
def part2_forward(fcgf_A, fcgf_B, yoho_A, yoho_B, pre_idx, sd, P, N, dtype=torch.float32,
                  faithful_cost=False):
    """extractor_PartII.batch_create swap (tests/extractor.py:125-138) + PartII_test.forward
    (utils/network.py:259-278).  Inputs are the per-match rows [M,32,60] of fragment A (id0) and B (id1).
    before_eqv0 <- FCGF_B, before_eqv1 <- FCGF_A, after_eqv0 <- YOHO_B, after_eqv1 <- YOHO_A.
    Returns quaternion [M,4] (w,x,y,z), unit norm."""
    tt = lambda a: torch.as_tensor(a).to(dtype)
    b0, b1, a0, a1 = tt(fcgf_B).clone(), tt(fcgf_A), tt(yoho_B).clone(), tt(yoho_A)
    pre = torch.as_tensor(pre_idx).long()
    Pt = torch.from_numpy(P)
    with torch.no_grad():
        perm = Pt[pre]                                              # [M,60]   :266-268
        b0 = torch.gather(b0, 2, perm[:, None, :].expand(-1, 32, -1))
        a0 = torch.gather(a0, 2, perm[:, None, :].expand(-1, 32, -1))
        z0 = torch.cat([b0, b1, a0, a1], dim=1)                     # [M,128,60]  :269
        z1 = gconv(z0, sd, "Conv_init.comb_layer.2", N, dtype, "Conv_init.comb_layer.0", faithful_cost)   # :252-253
        blk = "PartII_SO3_Conv_layers.0."
        z2 = gconv(z1, sd, blk + "comb_layer_in.2", N, dtype, blk + "comb_layer_in.0", faithful_cost)
        z3 = gconv(z2, sd, blk + "comb_layer_out.2", N, dtype, blk + "comb_layer_out.0", faithful_cost) + z1  # :254-255
        fc = "PartII_To_R_FC."
        h = z3.unsqueeze(-1)                                        # [M,256,60,1]  :273-274
        h = F.conv2d(h, _t(sd, fc + "0.weight", dtype), _t(sd, fc + "0.bias", dtype))
        h = bn_relu(h, sd, fc + "1", dtype)
        h = F.conv2d(h, _t(sd, fc + "3.weight", dtype), _t(sd, fc + "3.bias", dtype))
        h = bn_relu(h, sd, fc + "4", dtype)
        h = F.conv2d(h, _t(sd, fc + "6.weight", dtype), _t(sd, fc + "6.bias", dtype))
        q = h[:, :, 0, 0]                                           # g=0 only   :276
        q = q / torch.norm(q, dim=1)[:, None]                       # :277
    return q


def quat_to_matrix_f32(q):
    """utils/r_eval.py:94-110 evaluated on numpy float32 scalars (as tests/extractor.py:187 passes them):
    the arithmetic is float32, the result is stored into a float64 matrix."""
    w, x, y, z = (np.float32(q[0]), np.float32(q[1]), np.float32(q[2]), np.float32(q[3]))
    two = np.float32(2)
    one = np.float32(1)
    m = np.eye(3)
    m[0, 0] = one - two * y * y - two * z * z
    m[0, 1] = two * x * y - two * z * w
    m[0, 2] = two * x * z + two * y * w
    m[1, 0] = two * x * y + two * z * w
    m[1, 1] = one - two * x * x - two * z * z
    m[1, 2] = two * y * z - two * x * w
    m[2, 0] = two * x * z - two * y * w
    m[2, 1] = two * y * z + two * x * w
    m[2, 2] = one - two * x * x - two * y * y
    return m


def part2_transforms(quat, pre_idx, keys0, keys1, Rgroup):
    """tests/extractor.py:185-201: R = R(q) @ Rgroup_f32[idx]; t = key0 - key1 @ R.T; [M,3,4] f64."""
    Rg32 = Rgroup.astype(np.float32)
    q = np.asarray(quat, dtype=np.float32)
    out = np.zeros((q.shape[0], 3, 4))
    for i in range(q.shape[0]):
        R = quat_to_matrix_f32(q[i]) @ Rg32[int(pre_idx[i])]
        t = keys0[i] - keys1[i] @ R.T
        out[i, :, :3] = R
        out[i, :, 3] = t
    return out


# ---------------------------------------------------------------------------------------------
# E1: YOHO-C bin statistics (pure integer/fp64 bookkeeping; the loops live in estimator_oracle)
# ---------------------------------------------------------------------------------------------
def dr_statistic(dr_index):
    """yohoc.DR_statictic (tests/estimator.py:34-51): per-bin member lists and sampling probabilities.
    Returns (members: list of 60 lists, p[60] f64) or (None, None)."""
    members = [[] for _ in range(G)]
    for t in range(len(dr_index)):
        members[int(dr_index[t])].append(t)
    prob = []
    for i in range(G):
        c = len(members[i])
        if c < 2:
            prob.append(0)
        else:
            num = float(c) / 100.0
            prob.append(num * (num - 0.01) * (num - 0.02))
    prob = np.array(prob)
    if np.sum(prob) < 1e-4:
        return None, None
    return members, prob / np.sum(prob)


def draw_yohoc_hypotheses(members, prob, max_iter, rng=np.random):
    """The reference's draw order inside the RANSAC loop (tests/estimator.py:119-126): one categorical
    draw for the bin, then three member draws WITH replacement.  Returns int32 [n,3] match ids."""
    hyp = []
    it = 0
    exec_time = 0
    while it < max_iter:
        if exec_time > 50000:
            break
        exec_time += 1
        r = rng.choice(range(G), p=prob)
        if len(members[r]) < 2:
            continue
        it += 1
        hyp.append(rng.choice(np.array(members[r]), 3))
    return np.array(hyp, dtype=np.int32).reshape(-1, 3)


# ---------------------------------------------------------------------------------------------
# "next" row: group-feature lift tail
# ---------------------------------------------------------------------------------------------
def lift_group_features(kps, pts_list, feats_list, Rgroup):
    """YOHO_testset.py:153-166: per rotation g, Keys @ R_g^T (float64), KNN(1) of the float64 keypoints into the float32
    down-sampled cloud (float64 distances by type promotion, utils/knn_search.py:17-24), gather the backbone feature,
    stack to [K,32,60] in g order.  Returns (features f32, nn int64 [60,K])."""
    out, nns = [], []
    for g in range(G):
        keys = np.asarray(kps, np.float64) @ Rgroup[g].T
        src = torch.from_numpy(keys)                                # float64 [K,3]
        tgt = torch.from_numpy(np.asarray(pts_list[g], np.float32)) # float32 [n,3]
        ids = []
        for i in range(0, src.shape[0], 500):
            d2 = torch.sum((src[i:i + 500].unsqueeze(1) - tgt.unsqueeze(0)).pow(2), 2)
            ids.append(torch.sqrt(d2 + 1e-7).min(dim=1)[1])
        nn = torch.cat(ids).numpy()
        nns.append(nn)
        out.append(np.asarray(feats_list[g], np.float32)[nn][:, :, None])
    return np.concatenate(out, axis=-1), np.stack(nns)
