"""The tail of the reference's test-set preparation (`YOHO_testset.py:153-166`) on the device: given, for each of the 60
group rotations, the down-sampled rotated cloud and its FCGF features (the backbone itself is out of scope), produce the
`FCGF_Input_Group_feature` tensor [K,32,60] that PartI consumes — without the per-rotation host round trips."""
from .engine import get_engine


def lift_group_features(kps, pts_list, feats_list, so3_dir=None):
    """kps [K,3] float64 keypoints; pts_list[g] [n_g,3], feats_list[g] [n_g,32] for g in range(60) -> torch CUDA [K,32,60]."""
    return get_engine(so3_dir=so3_dir).lift_group_features(kps, pts_list, feats_list)
