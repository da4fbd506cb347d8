"""Drop-in for the reference's `tests/matcher.py`: `matcher_dual`, `name2matcher` (tests/matcher.py:15-53).
Descriptor = float32 mean over the 60 group elements of the SAVED eqv (not the network's `inv`), both 1-NN
searches and the mutual filter run in one device pass; `Match/{id0}-{id1}.npy` is int64 [M,2]."""
import os
import numpy as np
import torch
import tqdm

from .hostutil import make_non_exists_dir, feature_set_name
from .knn_search import knn_module
from .engine import get_engine


class matcher_dual:
    def __init__(self, cfg):
        self.cfg = cfg
        self.KNN = knn_module.KNN(1)
        self._so3 = getattr(cfg, "SO3_related_files", None)

    def match_features(self, feats0, feats1):
        """feats [K,32,60] (numpy or tensor) -> int64 [M,2] mutual matches (tests/matcher.py:35-48)."""
        eng = get_engine(so3_dir=self._so3)
        d0 = eng.group_mean(feats0)
        d1 = eng.group_mean(feats1)
        pairs, n = eng.mutual_nn(d0, d1)
        M = int(n.item())
        return pairs[:M].cpu().numpy()

    def match(self, dataset):
        print(f'match the keypoints on {dataset.name}')
        Save_dir = f'{self.cfg.output_cache_fn}/Testset/{dataset.name}/Match'
        make_non_exists_dir(Save_dir)
        datasetname = feature_set_name(dataset.name)
        Feature_dir = f'{self.cfg.output_cache_fn}/Testset/{datasetname}/YOHO_Output_Group_feature'
        for pair in tqdm.tqdm(dataset.pair_ids):
            id0, id1 = pair
            if os.path.exists(f'{Save_dir}/{id0}-{id1}.npy'):
                continue
            feats0 = np.load(f'{Feature_dir}/{id0}.npy')   # K,32,60
            feats1 = np.load(f'{Feature_dir}/{id1}.npy')
            np.save(f'{Save_dir}/{id0}-{id1}.npy', self.match_features(feats0, feats1))


name2matcher = {
    'Match': matcher_dual,
}
