"""Icosahedral-group tables used by every stage of the hot path.

The three tables are DATA shipped by the reference (`group_related/*.npy`, all stored float64):
  Rotation.npy                      (60,3,3)  R_g, R_0 = I
  60_60.npy                         (60,60)   P[a][b] = idx(R_b . R_a)   ("R_index_permu")
  Nei_Index_in_SO3_ordered_13.npy   (60,13)   N[g][k] = idx(R_{h_k} . R_g), N[g][0] = g
They are loaded from `cfg.SO3_related_files` when the caller gives one (reference behaviour:
utils/network.py:72-74,223-226; tests/extractor.py:67,110; tests/estimator.py:283-284) and from the
packaged copy in `yoho_b200/data/group_related` otherwise.

Derived tables (new here, used by the pruned PartII evaluation):
  hop1  = N[0]                      13 elements whose outputs feed g=0
  hop2  = ordered union of N[e] for e in hop1   (45 elements)
  idx tables for the generic gather-GEMM kernel (`rows_out x 13` indices into `rows_in`).
"""
import os
import functools
import numpy as np

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "group_related")

G = 60       # group order
TAPS = 13    # kernel support of the group convolution


class GroupTables:
    def __init__(self, so3_dir=None):
        d = so3_dir if so3_dir and os.path.isdir(so3_dir) else _PKG_DIR
        self.dir = d
        self.R = np.load(os.path.join(d, "Rotation.npy")).astype(np.float64)            # [60,3,3]
        self.P = np.load(os.path.join(d, "60_60.npy")).astype(np.int64)                  # [60,60]
        self.N = np.load(os.path.join(d, "Nei_Index_in_SO3_ordered_13.npy")).astype(np.int64)  # [60,13]
        assert self.R.shape == (G, 3, 3) and self.P.shape == (G, G) and self.N.shape == (G, TAPS)
        # receptive field of output element g=0 (PartII reads only g=0: utils/network.py:272-276)
        self.hop1 = [int(v) for v in self.N[0]]
        hop2 = []
        for e in self.hop1:
            for k in range(TAPS):
                v = int(self.N[e][k])
                if v not in hop2:
                    hop2.append(v)
        self.hop2 = hop2
        assert len(self.hop1) == 13 and len(self.hop2) == 45 and self.hop2[0] == 0

    # --- index tables for the gather-GEMM (rows_out x 13 -> row in rows_in) -------------------
    def idx_full(self):
        """PartI and un-pruned layers: rows_in = rows_out = 60, idx[g][k] = N[g][k]."""
        return self.N.astype(np.int32).copy()

    def idx_p2_init(self):
        """PartII Conv_init evaluated only at the 45 two-hop elements: rows_in=60, rows_out=45."""
        return np.array([[self.N[e][k] for k in range(TAPS)] for e in self.hop2], dtype=np.int32)

    def idx_p2_a(self):
        """PartII comb_layer_in evaluated at the 13 one-hop elements, reading the 45-list."""
        pos = {e: i for i, e in enumerate(self.hop2)}
        return np.array([[pos[int(self.N[e][k])] for k in range(TAPS)] for e in self.hop1], dtype=np.int32)

    def idx_p2_b(self):
        """PartII comb_layer_out at g=0 only, reading the 13-list: idx[0][k] = k."""
        return np.arange(TAPS, dtype=np.int32)[None, :].copy()

    def hop2_pos_of_zero(self):
        return self.hop2.index(0)


@functools.lru_cache(maxsize=8)
def load(so3_dir=None) -> GroupTables:
    return GroupTables(so3_dir)
