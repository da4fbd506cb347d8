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
    a = ap.parse_args()
    sys.argv = [sys.argv[0]]                      # parses/*.py parse sys.argv at import time

    import numpy as np
    import torch
    import ref_shim
    if a.backend == "reference" and a.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""
    if a.backend == "reference" and a.device == "cuda" and a.tf32 >= 0:
        torch.backends.cudnn.allow_tf32 = bool(a.tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(a.tf32)
    ref_shim.install()
    if a.backend == "yoho_b200":
        import yoho_b200.dropin as dropin
        dropin.install()
    import importlib
    pI = importlib.import_module("parses.parses_partI")
    pII = importlib.import_module("parses.parses_partII")
    cfgI, _ = pI.get_config()
    cfgII, _ = pII.get_config()
    ev = importlib.import_module("tests.evaluator")           # the reference's module, whatever the backend
    assert os.path.abspath(ev.__file__).startswith(os.path.abspath(ref_shim.REF_ROOT)), ev.__file__
    ext = sys.modules["tests.extractor"]
    backend_file = getattr(ext, "__file__", "")
    if a.backend == "yoho_b200":
        assert "yoho_b200" in backend_file, backend_file
        assert ev.name2extractor is ext.name2extractor

    from yoho_b200 import synth
    tmp = a.work
    for cfg in (cfgI, cfgII):
        cfg.output_cache_fn = os.path.join(tmp, 'cache')
        cfg.origin_data_dir = os.path.join(tmp, 'origin')
        cfg.model_fn = os.path.join(tmp, 'model')
        cfg.SO3_related_files = os.path.join(ref_shim.REF_ROOT, 'group_related')
    for part, d in (('PartI', 'PartI_train'), ('PartII', 'PartII_train')):
        os.makedirs(os.path.join(tmp, 'model', d), exist_ok=True)
        if a.weights == "synth":
            sd = synth.synth_state_dict(part, 0)
        else:
            sd = dict(np.load(os.path.join(HERE, '_ref', 'ckpt', part + '.npz')))
        torch.save({'best_para': 0.0, 'step': 0, 'network_state_dict': synth.to_torch_state_dict(sd)},
                   os.path.join(tmp, 'model', d, 'model_best.pth'))
    pair = synth.make_fragment_pair(a.K, seed=a.pair_seed, overlap=a.overlap, sigma=0.05)
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

    times = {}
    t0 = time.perf_counter()
    e1 = ev.Evaluator_PartI(cfgI, max_iter=a.max_iter)        # Test.py:54 -> name2evaluator['PartI']
    np.random.seed(a.c_seed)
    e1.run_onescene(ds)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    times['partI_s'] = time.perf_counter() - t0
    fmr, pair_fmrs = e1.Feature_match_Recall(ds, ratio=0.05)
    t0 = time.perf_counter()
    e2 = ev.Evaluator_PartII(cfgII, max_iter=a.max_iter)      # Test.py:64
    np.random.seed(a.o_seed)
    e2.run_onescene(ds)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    times['partII_s'] = time.perf_counter() - t0
    info = dict(backend=a.backend, device=a.device if a.backend == "reference" else "cuda", evaluator_file=ev.__file__,
                extractor_file=backend_file, estimator_class=type(e1.estimator).__module__ + "." + type(e1.estimator).__name__,
                fmr=float(fmr), pair_fmr=[float(v) for v in pair_fmrs], K=a.K, weights=a.weights, **times)
    with open(os.path.join(tmp, 'run_info.json'), 'w') as f:
        json.dump(info, f)
    print(json.dumps(info))


if __name__ == '__main__':
    main()
