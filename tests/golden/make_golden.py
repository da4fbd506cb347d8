"""Generate the committed golden vectors by running the UNMODIFIED reference (imported from /root/reference
under oracle/ref_shim.py) on seeded synthetic inputs.  Run in the authoring container only:

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  pipeline_synth.npz     the reference's whole file-based pipeline on one synthetic 128-keypoint pair with the
                         seeded synthetic weights of yoho_b200.synth (regenerable anywhere from the seed):
                         Extract -> match -> PartI_Rindex -> yohoc.ransac -> PartII_R_pre -> yohoo.ransac,
                         plus the per-iteration triplets / SVD signs / scores the reference's yohoc loop saw, the
                         `center` array of its YOHO_C npz and the bytes of both `pre.log` files.
  stages_synth.npz       stage-level vectors on other seeds: PartI (eqv, inv), KNN both ways, rotation
                         correlation, PartII quaternion.
  pipeline_realckpt.npz  same as pipeline_synth with the reference's shipped checkpoints (the test that uses it
                         needs oracle/_ref/ckpt/*.npz, which `__graft_entry__.build()` extracts when
                         /root/reference is present; git-ignored, travels with gpurun).
"""
import os
import sys
import shutil
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import ref_shim  # noqa: E402
from yoho_b200 import synth  # noqa: E402

torch.set_num_threads(max(1, os.cpu_count() or 1))


class StubDataset:
    """10-line duck type of utils/dataset.py's ThrDMatchPartDataset."""

    def __init__(self, name, kps, gt):
        self.name = name
        self.pc_ids = ['0', '1']
        self.pair_ids = [('0', '1')]
        self._kps = kps
        self._gt = gt

    def get_transform(self, a, b):
        return self._gt

    def get_kps(self, i):
        return self._kps[int(i)]


def write_ckpt(path, sd):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({'best_para': 0.0, 'step': 0, 'network_state_dict': sd}, path)


def run_pipeline(ref, tmp, sdI, sdII, pair, seed, max_iter=1000):
    cfgI, cfgII = ref.cfgI, ref.cfgII
    for cfg in (cfgI, cfgII):
        cfg.output_cache_fn = os.path.join(tmp, 'cache')
        cfg.origin_data_dir = os.path.join(tmp, 'origin')
        cfg.model_fn = os.path.join(tmp, 'model')
        cfg.SO3_related_files = os.path.join(ref.root, 'group_related')
    write_ckpt(os.path.join(tmp, 'model', 'PartI_train', 'model_best.pth'), sdI)
    write_ckpt(os.path.join(tmp, 'model', 'PartII_train', 'model_best.pth'), sdII)
    name = 'synth/scene'
    base = os.path.join(tmp, 'cache', 'Testset', name)
    os.makedirs(os.path.join(base, 'FCGF_Input_Group_feature'))
    np.save(os.path.join(base, 'FCGF_Input_Group_feature', '0.npy'), pair['feat_A'])
    np.save(os.path.join(base, 'FCGF_Input_Group_feature', '1.npy'), pair['feat_B'])
    kdir = os.path.join(tmp, 'origin', name, 'Keypoints_PC')
    os.makedirs(kdir)
    np.save(os.path.join(kdir, 'cloud_bin_0Keypoints.npy'), pair['kps_A'])
    np.save(os.path.join(kdir, 'cloud_bin_1Keypoints.npy'), pair['kps_B'])
    gt = np.concatenate([pair['R_gt'], pair['t_gt'][:, None]], 1)
    ds = StubDataset(name, [pair['kps_A'], pair['kps_B']], gt)

    ref.extractor.extractor_PartI(cfgI).Extract(ds)
    ref.matcher.matcher_dual(cfgI).match(ds)
    ref.extractor.extractor_dr_index(cfgI).PartI_Rindex(ds)

    # YOHO-C with every hypothesis recorded (subclass of the reference class; its loop is untouched)
    log = dict(ids=[], sign=[], trans=[], overlap=[])

    class Recorder(ref.estimator.yohoc):
        def Threepps2Tran(self, k0, k1):
            T = super().Threepps2Tran(k0, k1)
            log['trans'].append(T.copy())
            log['sign'].append(1 if np.linalg.det(T[:, :3]) > 0 else -1)
            return T

        def overlap_cal(self, m0, m1, T):
            ov = super().overlap_cal(m0, m1, T)
            log['overlap'].append(ov)
            return ov

    # the triplet ids are recovered by replaying the same draws on a copy of the RNG state
    np.random.seed(seed)
    state = np.random.get_state()
    est = Recorder(cfgI)
    est.ransac(ds, max_iter)
    dr = np.load(os.path.join(base, 'Match', 'DR_index', '0-1.npy'))
    stat, prob = est.DR_statictic(dr)
    ids = []
    if prob is not None:
        np.random.set_state(state)
        it = 0
        while it < max_iter:
            r = np.random.choice(range(60), p=prob)
            if len(stat[r]) < 2:
                continue
            it += 1
            ids.append(np.random.choice(np.array(stat[r]), 3))
    ids = np.array(ids, dtype=np.int32).reshape(-1, 3)

    ref.extractor.extractor_PartII(cfgII).PartII_R_pre(ds)
    np.random.seed(seed + 1)
    state_o = np.random.get_state()
    ref.estimator.yohoo(cfgII).ransac(ds, max_iter)
    trans_pre = np.load(os.path.join(base, 'Match', 'Trans_pre', '0-1.npy'))
    np.random.set_state(state_o)
    index = np.arange(trans_pre.shape[0])
    np.random.shuffle(index)

    m = os.path.join(base, 'Match')
    c = np.load(os.path.join(m, 'YOHO_C', f'{max_iter}iters', '0-1.npz'), allow_pickle=True)
    o = np.load(os.path.join(m, 'YOHO_O', f'{max_iter}iters', '0-1.npz'), allow_pickle=True)
    def _bytes(fn):
        with open(fn, 'rb') as f:
            return np.frombuffer(f.read(), dtype=np.uint8).copy()
    out = dict(
        c_center=np.asarray(c['center']), c_prelog=_bytes(os.path.join(m, 'YOHO_C', f'{max_iter}iters', 'pre.log')),
        o_prelog=_bytes(os.path.join(m, 'YOHO_O', f'{max_iter}iters', 'pre.log')),
        eqv0=np.load(os.path.join(base, 'YOHO_Output_Group_feature', '0.npy')),
        eqv1=np.load(os.path.join(base, 'YOHO_Output_Group_feature', '1.npy')),
        matches=np.load(os.path.join(m, '0-1.npy')),
        dr_index=dr,
        trans_pre=trans_pre,
        c_seed=np.int64(seed), c_hyp=ids, c_sign=np.array(log['sign'], np.int8),
        c_hyp_trans=np.array(log['trans']), c_overlap=np.array(log['overlap']),
        c_trans=c['trans'], c_recalltime=np.int64(c['recalltime']),
        o_seed=np.int64(seed + 1), o_order=index[:max_iter].astype(np.int32),
        o_trans=o['trans'], o_recalltime=np.int64(o['recalltime']),
        c_dist=np.float64(cfgI.ransac_c_inlinerdist), o_dist=np.float64(cfgII.ransac_o_inlinerdist),
    )
    assert len(log['trans']) == ids.shape[0]
    return out


def stage_vectors(ref):
    sdI = synth.to_torch_state_dict(synth.synth_state_dict('PartI', 1))
    sdII = synth.to_torch_state_dict(synth.synth_state_dict('PartII', 1))
    out = {}
    # PartI on 40 keypoints (seed 21)
    x, _ = synth.make_fragment(40, 21)
    net = ref.network.PartI_test(ref.cfgI)
    net.load_state_dict(sdI)
    net.eval()
    with torch.no_grad():
        o = net(torch.from_numpy(x))
    out['p1_eqv'] = o['eqv'].numpy()
    out['p1_inv'] = o['inv'].numpy()
    # KNN(1) both ways on 300 x 257 random 32-d descriptors (seed 22), the reference's callable
    rs = np.random.RandomState(22)
    d0 = (rs.standard_normal((300, 32)) * 0.1).astype(np.float32)
    d1 = (rs.standard_normal((257, 32)) * 0.1).astype(np.float32)
    d1[:100] = d0[100:200] + (rs.standard_normal((100, 32)) * 0.01).astype(np.float32)
    knn = ref.knn_search.knn_module.KNN(1)
    t0 = torch.from_numpy(d0.T.copy())[None]
    t1 = torch.from_numpy(d1.T.copy())[None]
    dd01, a01 = knn(t1, t0)
    dd10, a10 = knn(t0, t1)
    out['knn_d0'], out['knn_d1'] = d0, d1
    out['knn_a01'], out['knn_a10'] = a01[0, 0].numpy(), a10[0, 0].numpy()
    out['knn_dist01'], out['knn_dist10'] = dd01[0, 0].numpy(), dd10[0, 0].numpy()
    # rotation correlation on 48 planted matches (seed 23)
    pr = synth.make_fragment_pair(48, seed=23, overlap=1.0, sigma=0.3)
    des1 = pr['feat_B'][pr['ids_B']]
    des2 = pr['feat_A'][pr['ids_A']]
    ex = ref.extractor.extractor_dr_index(ref.cfgI)
    B = des1.shape[0]
    x1 = torch.from_numpy(des1)[:, :, ex.Nei_in_SO3].reshape([B, 32, 60, 60])
    out['rot_cor'] = torch.einsum('bfag,bfg->ba', x1, torch.from_numpy(des2)).numpy()
    out['rot_idx'] = ex.Batch_Des2R_torch(torch.from_numpy(des1), torch.from_numpy(des2)).numpy()
    out['rot_planted'] = np.int64(pr['r'])
    # PartII on 24 matches (seed 24): feed the reference's forward exactly what batch_create would
    pp = synth.make_fragment_pair(24, seed=24, overlap=1.0, sigma=0.05)
    fA, fB = pp['feat_A'][pp['ids_A']], pp['feat_B'][pp['ids_B']]
    with torch.no_grad():
        yA = net(torch.from_numpy(fA))['eqv'].numpy()
        yB = net(torch.from_numpy(fB))['eqv'].numpy()
    pre = np.full((24,), pp['r'], dtype=np.int64)
    pre[::5] = (pre[::5] + 7) % 60
    net2 = ref.network.PartII_test(ref.cfgII)
    net2.load_state_dict(sdII, strict=False)
    net2.eval()
    batch = {'before_eqv0': torch.from_numpy(fB.copy()), 'before_eqv1': torch.from_numpy(fA.copy()),
             'after_eqv0': torch.from_numpy(yB.copy()), 'after_eqv1': torch.from_numpy(yA.copy()),
             'pre_idx': torch.from_numpy(pre)}
    with torch.no_grad():
        q = net2(batch)['quaternion_pre'].numpy()
    out['p2_yA'], out['p2_yB'], out['p2_pre'], out['p2_quat'] = yA, yB, pre, q
    return out


def extract_real_ckpt(ref, dst):
    os.makedirs(dst, exist_ok=True)
    for part, fn in (('PartI', 'PartI_train'), ('PartII', 'PartII_train')):
        sd = torch.load(os.path.join(ref.root, 'model', fn, 'model_best.pth'))['network_state_dict']
        keep = {k: v.numpy() for k, v in sd.items() if part == 'PartI' or not k.startswith('PartI_net.')}
        np.savez(os.path.join(dst, part + '.npz'), **keep)


def main():
    ref = ref_shim.load_reference()
    np.savez(os.path.join(HERE, 'stages_synth.npz'), **stage_vectors(ref))
    print('stages_synth.npz written')
    sdI = synth.to_torch_state_dict(synth.synth_state_dict('PartI', 0))
    sdII = synth.to_torch_state_dict(synth.synth_state_dict('PartII', 0))
    pair = synth.make_fragment_pair(128, seed=7, overlap=0.6, sigma=0.05)
    tmp = tempfile.mkdtemp(prefix='yoho_golden_')
    try:
        out = run_pipeline(ref, tmp, sdI, sdII, pair, seed=123)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez(os.path.join(HERE, 'pipeline_synth.npz'), **out)
    print('pipeline_synth.npz: M =', out['matches'].shape[0], 'c_recall', out['c_recalltime'], 'o_recall', out['o_recalltime'])
    # real checkpoints
    ck = os.path.join(ROOT, 'oracle', '_ref', 'ckpt')
    extract_real_ckpt(ref, ck)
    realI = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(ck, 'PartI.npz')).items()}
    realII = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(ck, 'PartII.npz')).items()}
    tmp = tempfile.mkdtemp(prefix='yoho_golden_')
    try:
        out = run_pipeline(ref, tmp, realI, realII, pair, seed=321)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez(os.path.join(HERE, 'pipeline_realckpt.npz'), **out)
    print('pipeline_realckpt.npz: M =', out['matches'].shape[0], 'c_recall', out['c_recalltime'], 'o_recall', out['o_recalltime'])


if __name__ == '__main__':
    main()
