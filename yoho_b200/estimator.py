"""Drop-in for the reference's `tests/estimator.py`: `yohoc`, `yohoc_mul`, `yohoo`, `R_pre_log`,
`name2estimator` (tests/estimator.py:12-347) — same constructors, `.ransac(dataset, max_iter)` protocol and
`.npz` / `pre.log` artefacts.

Randomness.  The reference draws from the global `np.random` inside its loops (tests/estimator.py:122,126,322).
These classes draw the SAME values from the SAME global stream in the SAME order on the host (so
`np.random.seed(s)` selects the same hypotheses as the reference), then score all hypotheses in one device
pass.  `rng='device'` switches to the counter-based device generator (same distribution, no host loop).

SVD sign.  With three matches the cross-covariance has rank 2 and the sign LAPACK gives the null-space singular
pair is rounding noise (det(V U^T) = +1 about half of the time): the reference's hypothesis is "whatever
np.linalg.svd returned".  In the default numpy-RNG mode `lapack_replay` therefore makes the reference's own
LAPACK call on the [iters,3,3] cross-covariance batch (tests/estimator.py:55-59; the batched gufunc runs the same
gesdd per matrix as the reference's scalar call), and hands the device sign(det(V U^T)) per hypothesis plus
LAPACK's transform for the rank-deficient triplets (a match drawn twice: the completion is arbitrary).  The device
solves and scores every hypothesis (csrc/estimator.cu); winner, `recalltime`, `center` and — through the
replayed LAPACK transform of the winning triplet — `trans` equal the reference's artefacts bit for bit
(tests/test_gpu_pipeline.py).  `rng='device'` uses the device's own sign rule (DESIGN.md "Estimator arithmetic").

`yohoc_mul` (one forked process per pair in the reference, :255-275) is the same device batch here — a CUDA
process must not fork.
"""
import numpy as np
import torch
from tqdm import tqdm

from .hostutil import make_non_exists_dir, feature_set_name
from .engine import get_engine


def R_pre_log(dataset, save_dir):
    # tests/estimator.py:12-24 — Redwood-format log consumed by RR_cal
    writer = open(f'{save_dir}/pre.log', 'w')
    pair_num = int(len(dataset.pc_ids))
    for pair in dataset.pair_ids:
        pc0, pc1 = pair
        ransac_result = np.load(f'{save_dir}/{pc0}-{pc1}.npz', allow_pickle=True)
        transform_pr = ransac_result['trans']
        writer.write(f'{int(pc0)}\t{int(pc1)}\t{pair_num}\n')
        writer.write(f'{transform_pr[0][0]}\t{transform_pr[0][1]}\t{transform_pr[0][2]}\t{transform_pr[0][3]}\n')
        writer.write(f'{transform_pr[1][0]}\t{transform_pr[1][1]}\t{transform_pr[1][2]}\t{transform_pr[1][3]}\n')
        writer.write(f'{transform_pr[2][0]}\t{transform_pr[2][1]}\t{transform_pr[2][2]}\t{transform_pr[2][3]}\n')
        writer.write(f'{0.0}\t{0.0}\t{0.0}\t{1.0}\n')
    writer.close()


class yohoc:
    def __init__(self, cfg, rng='numpy'):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_c_inlinerdist
        self.rng = rng
        self._so3 = getattr(cfg, "SO3_related_files", None)

    # ---- E1 -----------------------------------------------------------------------------------------
    def DR_statictic(self, DR_indexs):
        """tests/estimator.py:34-51: per-bin member lists and sampling probabilities; (None, None) when
        sum_bins n(n-.01)(n-.02), n = count/100, is below 1e-4."""
        DR_indexs = np.asarray(DR_indexs).astype(np.int64)
        order = np.argsort(DR_indexs, kind='stable')
        counts = np.bincount(DR_indexs, minlength=60)
        starts = np.concatenate([[0], np.cumsum(counts)])
        stat = {i: order[starts[i]:starts[i + 1]].tolist() for i in range(60)}
        prob = []
        for i in range(60):
            if counts[i] < 2:
                prob.append(0)
            else:
                num = float(counts[i]) / 100.0
                prob.append(num * (num - 0.01) * (num - 0.02))
        prob = np.array(prob)
        if np.sum(prob) < 1e-4:
            return None, None
        prob = prob / np.sum(prob)
        return stat, prob

    def draw_hypotheses(self, stat, prob, max_iter):
        """The reference's draw order (tests/estimator.py:119-126) on the global numpy stream."""
        hyp = np.empty((max_iter, 3), dtype=np.int32)
        members = {i: np.array(stat[i]) for i in range(60)}
        it = 0
        exec_time = 0
        while it < max_iter:
            if exec_time > 50000:
                break
            exec_time += 1
            R_index = np.random.choice(range(60), p=prob)
            if len(stat[R_index]) < 2:
                continue
            hyp[it] = np.random.choice(members[R_index], 3)
            it += 1
        return hyp[:it]

    @staticmethod
    def lapack_replay(Keys_m0, Keys_m1, hyp):
        """Threepps2Tran (tests/estimator.py:55-63) for every triplet at once, with the reference's operations:
        returns (trans [iters,3,4] f64 = the reference's per-hypothesis transforms, signs int8[iters]:
        +-1 = det(V U^T) of a rank-2 triplet, 2 = rank-deficient triplet -> the device takes trans[i] as given)."""
        hyp = np.asarray(hyp, np.int64).reshape(-1, 3)
        a = np.asarray(Keys_m0, np.float64)[hyp]
        b = np.asarray(Keys_m1, np.float64)[hyp]
        c0 = np.mean(a, 1, keepdims=True)
        c1 = np.mean(b, 1, keepdims=True)
        H = np.matmul((b - c1).transpose(0, 2, 1), a - c0)
        U, S, VT = np.linalg.svd(H)
        R = np.matmul(VT.transpose(0, 2, 1), U.transpose(0, 2, 1))
        t = c0 - np.matmul(c1, R.transpose(0, 2, 1))
        trans = np.concatenate([R, t.transpose(0, 2, 1)], 2)
        signs = np.where(np.linalg.det(R) > 0, 1, -1).astype(np.int8)
        dup = (hyp[:, 0] == hyp[:, 1]) | (hyp[:, 0] == hyp[:, 2]) | (hyp[:, 1] == hyp[:, 2])
        signs[dup | ~(S[:, 1] > 1e-12 * S[:, 0])] = 2
        return trans, signs

    # ---- E2-E4 on matched keypoints ---------------------------------------------------------------------
    def estimate(self, Keys_m0, Keys_m1, Index, max_iter=1000):
        """Matched keypoints [M,3] f64 x2 and rotation index [M] -> dict(trans, center, recalltime).
        Same outputs as the body of yohoc.ransac for one pair (tests/estimator.py:103-139)."""
        stat, prob = self.DR_statictic(Index)
        if prob is None:
            return dict(trans=np.eye(4), center=0, axis=0, recalltime=50001)
        eng = get_engine(so3_dir=self._so3)
        ref_trans = None
        if self.rng == 'device':
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            hyp_d, _ = eng.c_draw(np.asarray(Index, np.int64), max_iter, seed)
            res = eng.c_ransac(Keys_m0, Keys_m1, hyp_d, self.inliner_dist)
        else:
            hyp_d = self.draw_hypotheses(stat, prob, max_iter)
            ref_trans, signs = self.lapack_replay(Keys_m0, Keys_m1, hyp_d)
            res = eng.c_ransac(Keys_m0, Keys_m1, hyp_d, self.inliner_dist, signs=signs, fixed=ref_trans)
        bi = int(res['best_iter'].item())
        if bi < 0:
            return dict(trans=np.eye(4), center=np.ones([6, 3]), recalltime=0)
        # numpy mode: the winner's transform as LAPACK computed it (what the reference saves); the device's own solve of
        # the same triplet with the same sign agrees to ~1e-9 (tests/test_gpu_parity.py::test_c_ransac_golden_replay)
        trans = ref_trans[bi] if ref_trans is not None else res['T'].cpu().numpy()
        ids = (hyp_d[bi].cpu().numpy() if isinstance(hyp_d, torch.Tensor) else hyp_d[bi]).astype(np.int64)
        center = np.concatenate([np.asarray(Keys_m0)[ids], np.asarray(Keys_m1)[ids]], axis=0)
        return dict(trans=trans, center=center, recalltime=bi + 1)

    def _pair_inputs(self, dataset, match_dir, Index_dir, Keys_dir, pair):
        id0, id1 = pair
        Keys0 = np.load(f'{Keys_dir}/cloud_bin_{id0}Keypoints.npy')
        Keys1 = np.load(f'{Keys_dir}/cloud_bin_{id1}Keypoints.npy')
        pps = np.load(f'{match_dir}/{id0}-{id1}.npy').reshape(-1, 2)
        Index = np.load(f'{Index_dir}/{id0}-{id1}.npy')
        return Keys0[pps[:, 0]], Keys1[pps[:, 1]], Index

    def ransac(self, dataset, max_iter=1000):
        # tests/estimator.py:78-141
        match_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/Match'
        Index_dir = f'{match_dir}/DR_index'
        Save_dir = f'{match_dir}/YOHO_C/{max_iter}iters'
        make_non_exists_dir(Save_dir)
        datasetname = feature_set_name(dataset.name)
        Keys_dir = f'{self.cfg.origin_data_dir}/{datasetname}/Keypoints_PC'
        print(f'Ransac with YOHO-C on {dataset.name}:')
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            Keys_m0, Keys_m1, Index = self._pair_inputs(dataset, match_dir, Index_dir, Keys_dir, pair)
            out = self.estimate(Keys_m0, Keys_m1, Index, max_iter)
            np.savez(f'{Save_dir}/{id0}-{id1}.npz', **out)
        R_pre_log(dataset, Save_dir)


class yohoc_mul(yohoc):
    """tests/estimator.py:145-275.  Same results contract as `yohoc`; the per-pair process pool of the reference
    is replaced by device parallelism over hypotheses (no fork after CUDA initialisation).

    Randomness: the reference forks one worker per pair (`Pool(len(pair_ids))`, :269-273), so every worker starts
    from a COPY of the parent's global numpy state and the parent's own stream does not advance.  That is what this
    class reproduces: each pair draws from the state the caller had on entry, and the state is restored on exit.
    (When the reference's pool hands two pairs to one worker, the second continues the first one's stream; that
    assignment is scheduler-dependent and not reproducible in the reference either.)"""

    def estimate(self, Keys_m0, Keys_m1, Index, max_iter=1000):
        if self.rng != 'device' and getattr(self, '_fork_state', None) is not None:
            np.random.set_state(self._fork_state)
        return super().estimate(Keys_m0, Keys_m1, Index, max_iter)

    def ransac(self, dataset, max_iter=1000):
        self._fork_state = np.random.get_state()
        try:
            super().ransac(dataset, max_iter)
        finally:
            np.random.set_state(self._fork_state)
            self._fork_state = None
        print('Done')


class yohoo:
    def __init__(self, cfg, rng='numpy'):
        self.cfg = cfg
        self.inliner_dist = cfg.ransac_o_inlinerdist
        self.rng = rng
        self._so3 = getattr(cfg, "SO3_related_files", None)

    def estimate(self, Keys_m0, Keys_m1, Trans, max_iter=1000):
        """tests/estimator.py:321-336 for one pair: shuffle, keep the first max_iter hypotheses, score, keep the
        first strictly-best.  Returns dict(trans, recalltime)."""
        eng = get_engine(so3_dir=self._so3)
        n = Trans.shape[0]
        if self.rng == 'device':
            order = eng.o_order(n, int(np.random.randint(0, 2 ** 31 - 1)))[:max_iter]
        else:
            index = np.arange(n)
            np.random.shuffle(index)
            order = index[0:max_iter].astype(np.int32)
        res = eng.o_score(Keys_m0, Keys_m1, Trans, self.inliner_dist, order=order)
        bi = int(res['best_iter'].item())
        if bi < 0:
            return dict(trans=np.eye(4), recalltime=0)
        return dict(trans=res['T'].cpu().numpy(), recalltime=bi)

    def ransac(self, dataset, max_iter=1000):
        # tests/estimator.py:298-340
        match_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/Match'
        Trans_dir = f'{match_dir}/Trans_pre'
        Save_dir = f'{match_dir}/YOHO_O/{max_iter}iters'
        make_non_exists_dir(Save_dir)
        print(f'Ransac with YOHO-O on {dataset.name}:')
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            Keys0 = dataset.get_kps(id0)
            Keys1 = dataset.get_kps(id1)
            pps = np.load(f'{match_dir}/{id0}-{id1}.npy').reshape(-1, 2)
            Keys_m0 = Keys0[pps[:, 0]]
            Keys_m1 = Keys1[pps[:, 1]]
            Trans = np.load(f'{Trans_dir}/{id0}-{id1}.npy')
            out = self.estimate(Keys_m0, Keys_m1, Trans, max_iter)
            np.savez(f'{Save_dir}/{id0}-{id1}.npz', **out)
        R_pre_log(dataset, Save_dir)


name2estimator = {
    'yohoc': yohoc,
    'yohoc_mul': yohoc_mul,
    'yohoo': yohoo,
}
