"""Drop-in for the reference's `tests/extractor.py`: `extractor_PartI`, `extractor_dr_index`,
`extractor_PartII`, `name2extractor` — same constructors, methods, on-disk artefacts and skip-if-exists
behaviour (tests/extractor.py:18-207), with the arithmetic in libyoho_b200.so.

Differences from the reference that do not change results:
  * a fragment is pushed through PartI in one C-ABI call instead of `test_batch_size` slices with a host
    round-trip each (keypoints are independent, so the batching is not observable);
  * PartI_Rindex / PartII_R_pre gather matched rows on the device from the fragment tensors instead of
    fancy-indexing 38 MB arrays on the host;
  * the quaternion -> matrix and per-match translation loops (tests/extractor.py:185-199) run in the PartII
    head kernel.
"""
import os
import numpy as np
import torch
from tqdm import tqdm

from .hostutil import make_non_exists_dir, feature_set_name
from .network import name2network
from .engine import get_engine


class extractor_PartI:
    def __init__(self, cfg):
        self.cfg = cfg
        self.network = name2network[f'{self.cfg.test_network_type}'](self.cfg).cuda()
        self.model_fn = f'{self.cfg.model_fn}/{self.cfg.train_network_type}/model.pth'
        self.best_model_fn = f'{self.cfg.model_fn}/{self.cfg.train_network_type}/model_best.pth'

    def _load_model(self):
        # tests/extractor.py:26-34
        if os.path.exists(self.best_model_fn):
            checkpoint = torch.load(self.best_model_fn, map_location='cpu', weights_only=False)
            best_para = checkpoint['best_para']
            self.network.load_state_dict(checkpoint['network_state_dict'])
            print(f'Resuming best para {best_para}')
        else:
            raise ValueError("No model exists")

    def Extract(self, dataset):
        # tests/extractor.py:37-60
        self._load_model()
        self.network.eval()
        FCGF_input_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/FCGF_Input_Group_feature'
        YOHO_output_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/YOHO_Output_Group_feature'
        make_non_exists_dir(YOHO_output_dir)
        print(f'Extracting the PartI descriptors on {dataset.name}')
        for pc_id in tqdm(dataset.pc_ids):
            if os.path.exists(f'{YOHO_output_dir}/{pc_id}.npy'):
                continue
            Input_feature = np.load(f'{FCGF_input_dir}/{pc_id}.npy')           # K*32*60
            batch = torch.from_numpy(Input_feature.astype(np.float32)).cuda()
            with torch.no_grad():
                batch_output = self.network(batch)
            np.save(f'{YOHO_output_dir}/{pc_id}.npy', batch_output['eqv'].cpu().numpy())


class extractor_dr_index:
    def __init__(self, cfg):
        self.cfg = cfg
        self._so3 = getattr(cfg, "SO3_related_files", None)

    @property
    def engine(self):
        return get_engine(so3_dir=self._so3)

    def Des2R_torch(self, des1_eqv, des2_eqv):            # beforerot afterrot  [F,60] each
        return self.engine.rot_argmax(des1_eqv[None], des2_eqv[None])[0]

    def Batch_Des2R_torch(self, des1_eqv, des2_eqv):      # [B,F,60] each -> [B] int64 (tests/extractor.py:74-78)
        return self.engine.rot_argmax(des1_eqv, des2_eqv)

    def PartI_Rindex(self, dataset):
        # tests/extractor.py:80-100
        match_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/Match'
        Save_dir = f'{match_dir}/DR_index'
        make_non_exists_dir(Save_dir)
        datasetname = feature_set_name(dataset.name)
        Feature_dir = f'{self.cfg.output_cache_fn}/Testset/{datasetname}/YOHO_Output_Group_feature'
        print(f'extract the drindex of the matches on {dataset.name}')
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            if os.path.exists(f'{Save_dir}/{id0}-{id1}.npy'):
                continue
            match_pps = np.load(f'{match_dir}/{id0}-{id1}.npy')
            feats0 = torch.from_numpy(np.load(f'{Feature_dir}/{id0}.npy').astype(np.float32)).cuda()
            feats1 = torch.from_numpy(np.load(f'{Feature_dir}/{id1}.npy').astype(np.float32)).cuda()
            # Batch_Des2R_torch(feats1[m1], feats0[m0]) with the row gather done inside the kernel
            pre_idxs = self.engine.rot_argmax(feats1, feats0, pairs=match_pps.reshape(-1, 2)).cpu().numpy()
            np.save(f'{Save_dir}/{id0}-{id1}.npy', pre_idxs)


class extractor_PartII:
    def __init__(self, cfg):
        self.cfg = cfg
        self.network = name2network[f'{self.cfg.test_network_type}'](self.cfg).cuda()
        self.model_fn = f'{self.cfg.model_fn}/{self.cfg.train_network_type}/model.pth'
        self.best_model_fn = f'{self.cfg.model_fn}/{self.cfg.train_network_type}/model_best.pth'

    def _load_model(self):
        # tests/extractor.py:113-122 (strict=False: the checkpoint also carries a nested PartI copy)
        if os.path.exists(self.best_model_fn):
            print(self.best_model_fn)
            checkpoint = torch.load(self.best_model_fn, map_location='cpu', weights_only=False)
            best_para = checkpoint['best_para']
            self.network.load_state_dict(checkpoint['network_state_dict'], strict=False)
            print(f'Resuming best para {best_para}')
        else:
            raise ValueError("No model exists")

    def batch_create(self, feats0_fcgf, feats1_fcgf, feats0_yoho, feats1_yoho, index_pre, start, end):
        # tests/extractor.py:125-138 — note the 0 <-> 1 exchange ("feats0 -> feats1_in_batch for it is afterrot")
        t = lambda a: torch.from_numpy(a[start:end].astype(np.float32))
        return {
            'before_eqv0': t(feats1_fcgf),
            'before_eqv1': t(feats0_fcgf),
            'after_eqv0': t(feats1_yoho),
            'after_eqv1': t(feats0_yoho),
            'pre_idx': torch.from_numpy(index_pre[start:end].astype(np.int64)),
        }

    def PartII_R_pre(self, dataset):
        # tests/extractor.py:142-201
        self._load_model()
        self.network.eval()
        self.network._ensure_uploaded()
        eng = self.network.engine
        match_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/Match'
        DRindex_dir = f'{match_dir}/DR_index'
        Save_dir = f'{match_dir}/Trans_pre'
        make_non_exists_dir(Save_dir)
        datasetname = feature_set_name(dataset.name)
        FCGF_dir = f'{self.cfg.output_cache_fn}/Testset/{datasetname}/FCGF_Input_Group_feature'
        YOHO_dir = f'{self.cfg.output_cache_fn}/Testset/{datasetname}/YOHO_Output_Group_feature'
        print(f'extracting the PartII feature on {dataset.name}')
        for pair in tqdm(dataset.pair_ids):
            id0, id1 = pair
            if os.path.exists(f'{Save_dir}/{id0}-{id1}.npy'):
                continue
            pps = np.load(f'{match_dir}/{id0}-{id1}.npy').reshape(-1, 2)
            Index_pre = np.load(f'{DRindex_dir}/{id0}-{id1}.npy')
            load = lambda d, i: torch.from_numpy(np.load(f'{d}/{i}.npy').astype(np.float32)).cuda()
            f0, f1 = load(FCGF_dir, id0), load(FCGF_dir, id1)
            y0, y1 = load(YOHO_dir, id0), load(YOHO_dir, id1)
            Keys0 = dataset.get_kps(id0)
            Keys1 = dataset.get_kps(id1)
            _, trans = eng.part2(f0, f1, y0, y1, Index_pre, pairs=pps, kps0=np.asarray(Keys0, np.float64),
                                 kps1=np.asarray(Keys1, np.float64))
            np.save(f'{Save_dir}/{id0}-{id1}.npy', trans.cpu().numpy())


name2extractor = {
    'PartI': extractor_PartI,
    'PartII': extractor_PartII,
}
