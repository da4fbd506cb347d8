"""TEST INFRASTRUCTURE — drives the REFERENCE's own evaluator (tests/evaluator.py:28-47,103-117: Evaluator_PartI /
Evaluator_PartII .run_onescene) on a one-pair synthetic stub dataset, in a process of its own:

    python oracle/run_ref_evaluator.py --backend yoho_b200 --work DIR [--weights synth|real] [--K 128]
    python oracle/run_ref_evaluator.py --backend reference --work DIR [--device cpu|cuda]

--backend yoho_b200   `yoho_b200.dropin.install()` first: the reference's evaluator, registries, parsers and on-disk
                      protocol run UNCHANGED, with name2extractor / name2matcher / name2estimator / name2network
                      resolving to this package (the drop-in claim of INTEGRATION.md, executed).
--backend reference   the unmodified reference end to end (torch CPU, or its own .cuda() path with --device cuda).

The reference's sources are read from /root/reference when present (authoring container), else from the git-ignored copy
`oracle/_ref/src` that `__graft_entry__.build()` makes so that they travel to the GPU box (never committed).
Artefacts land under DIR/cache/Testset/synth/scene/... exactly as the reference writes them; tests compare them with
tests/golden/*.npz.  Used by tests/test_gpu_reference_evaluator.py and bench.py's reference legs only.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


class StubDataset:
    """Duck type of utils/dataset.py's per-scene dataset object."""

    def __init__(self, name, kps, gt):
        self.name, self.pc_ids, self.pair_ids = name, ['0', '1'], [('0', '1')]
        self._kps, self._gt = kps, gt

    def get_transform(self, a, b):
        return self._gt

    def get_kps(self, i):
        return self._kps[int(i)]


def setup(backend, device="cuda", tf32=-1):
    """Import the reference's evaluator stack (unmodified) with the chosen backend underneath.  Returns a namespace
    (ev = the reference's tests.evaluator module, cfgI, cfgII, ref_root, extractor_file)."""
    import types
    import importlib
    import torch
    import ref_shim
    if backend == "reference" and device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""           # before CUDA is initialised: the reference's .cuda() calls become no-ops
    if backend == "reference" and device == "cuda" and tf32 >= 0:
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    ref_shim.install()
    if backend == "yoho_b200":
        import yoho_b200.dropin as dropin
        dropin.install()
    argv, sys.argv = sys.argv, [sys.argv[0]]              # parses/*.py parse sys.argv at import time
    try:
        pI = importlib.import_module("parses.parses_partI")
        pII = importlib.import_module("parses.parses_partII")
        cfgI, _ = pI.get_config()
        cfgII, _ = pII.get_config()
    finally:
        sys.argv = argv
    ev = importlib.import_module("tests.evaluator")       # the reference's module, whatever the backend
    assert os.path.abspath(ev.__file__).startswith(os.path.abspath(ref_shim.REF_ROOT)), ev.__file__
    ext = sys.modules["tests.extractor"]
    backend_file = getattr(ext, "__file__", "")
    if backend == "yoho_b200":
        assert "yoho_b200" in backend_file, backend_file
        assert ev.name2extractor is ext.name2extractor
    else:
        assert os.path.abspath(backend_file).startswith(os.path.abspath(ref_shim.REF_ROOT)), backend_file
    return types.SimpleNamespace(ev=ev, cfgI=cfgI, cfgII=cfgII, ref_root=ref_shim.REF_ROOT, extractor_file=backend_file,
                                 backend=backend, device=device if backend == "reference" else "cuda")


def run_pair(ns, work, K=128, pair_seed=7, overlap=0.6, c_seed=123, o_seed=124, max_iter=1000, weights="synth", fmr=True):
    """One synthetic K-keypoint pair through Evaluator_PartI.run_onescene + Evaluator_PartII.run_onescene in directory `work`
    (the reference's on-disk protocol).  Returns wall-clock seconds per evaluator and bookkeeping."""
    import numpy as np
    import torch
    from yoho_b200 import synth
    ev, cfgI, cfgII = ns.ev, ns.cfgI, ns.cfgII
    tmp = work
    for cfg in (cfgI, cfgII):
        cfg.output_cache_fn = os.path.join(tmp, 'cache')
        cfg.origin_data_dir = os.path.join(tmp, 'origin')
        cfg.model_fn = os.path.join(tmp, 'model')
        cfg.SO3_related_files = os.path.join(ns.ref_root, 'group_related')
    for part, d in (('PartI', 'PartI_train'), ('PartII', 'PartII_train')):
        os.makedirs(os.path.join(tmp, 'model', d), exist_ok=True)
        if weights == "synth":
            sd = synth.synth_state_dict(part, 0)
        else:
            sd = dict(np.load(os.path.join(HERE, '_ref', 'ckpt', part + '.npz')))
        torch.save({'best_para': 0.0, 'step': 0, 'network_state_dict': synth.to_torch_state_dict(sd)},
                   os.path.join(tmp, 'model', d, 'model_best.pth'))
    pair = synth.make_fragment_pair(K, seed=pair_seed, overlap=overlap, sigma=0.05)
    name = 'synth/scene'
    base = os.path.join(tmp, 'cache', 'Testset', name)
    os.makedirs(os.path.join(base, 'FCGF_Input_Group_feature'), exist_ok=True)
    np.save(os.path.join(base, 'FCGF_Input_Group_feature', '0.npy'), pair['feat_A'])
    np.save(os.path.join(base, 'FCGF_Input_Group_feature', '1.npy'), pair['feat_B'])
    kdir = os.path.join(tmp, 'origin', name, 'Keypoints_PC')
    os.makedirs(kdir, exist_ok=True)
    np.save(os.path.join(kdir, 'cloud_bin_0Keypoints.npy'), pair['kps_A'])
    np.save(os.path.join(kdir, 'cloud_bin_1Keypoints.npy'), pair['kps_B'])
    gt = np.concatenate([pair['R_gt'], pair['t_gt'][:, None]], 1)
    ds = StubDataset(name, [pair['kps_A'], pair['kps_B']], gt)
    cfgI.ok_match_dist_threshold = cfgII.ok_match_dist_threshold = 0.1

    def sync():
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    times = {}
    sync()
    t0 = time.perf_counter()
    e1 = ev.Evaluator_PartI(cfgI, max_iter=max_iter)          # Test.py:54 -> name2evaluator['PartI']
    np.random.seed(c_seed)
    e1.run_onescene(ds)
    sync()
    times['partI_s'] = time.perf_counter() - t0
    out = dict(K=K, weights=weights)
    if fmr:
        f, pair_fmrs = e1.Feature_match_Recall(ds, ratio=0.05)
        out.update(fmr=float(f), pair_fmr=[float(v) for v in pair_fmrs])
    sync()
    t0 = time.perf_counter()
    e2 = ev.Evaluator_PartII(cfgII, max_iter=max_iter)        # Test.py:64
    np.random.seed(o_seed)
    e2.run_onescene(ds)
    sync()
    times['partII_s'] = time.perf_counter() - t0
    m = np.load(os.path.join(base, 'Match', '0-1.npy'))
    c = np.load(os.path.join(base, 'Match', 'YOHO_C', f'{max_iter}iters', '0-1.npz'), allow_pickle=True)
    R = c['trans'][:3, :3]
    rot_err = float(np.degrees(np.arccos(np.clip((np.trace(R.T @ pair['R_gt']) - 1) / 2, -1, 1))))
    out.update(times, seconds=times['partI_s'] + times['partII_s'], M=int(m.shape[0]), c_rot_err_deg=rot_err,
               estimator_class=type(e1.estimator).__module__ + "." + type(e1.estimator).__name__)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["yoho_b200", "reference"], required=True)
    ap.add_argument("--device", choices=["cpu", "cuda"], default="cuda")
    ap.add_argument("--work", required=True)
    ap.add_argument("--weights", choices=["synth", "real"], default="synth")
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--pair-seed", type=int, default=7)
    ap.add_argument("--overlap", type=float, default=0.6)
    ap.add_argument("--c-seed", type=int, default=123)
    ap.add_argument("--o-seed", type=int, default=124)
    ap.add_argument("--max-iter", type=int, default=1000)
    ap.add_argument("--tf32", type=int, default=-1, help="reference on cuda: 0/1 force torch's TF32 switches, -1 = torch defaults")
    ap.add_argument("--warmup-K", type=int, default=0, help="run a pair of this size first (untimed; model load, cuDNN selection)")
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    import torch
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    ns = setup(a.backend, a.device, a.tf32)
    if a.warmup_K > 0:
        run_pair(ns, os.path.join(a.work, 'warmup'), K=a.warmup_K, weights=a.weights, fmr=False)
    r = run_pair(ns, a.work, a.K, a.pair_seed, a.overlap, a.c_seed, a.o_seed, a.max_iter, a.weights)
    info = dict(backend=ns.backend, device=ns.device, evaluator_file=ns.ev.__file__, extractor_file=ns.extractor_file,
                threads=torch.get_num_threads(), tf32=a.tf32, **r)
    with open(os.path.join(a.work, 'run_info.json'), 'w') as f:
        json.dump(info, f)
    print(json.dumps(info))


if __name__ == '__main__':
    main()
