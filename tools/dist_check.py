"""Multi-rank checks on real GPUs (run under torchrun, one rank per GPU, NCCL):
  1. config 5: cross-rank mutual 1-NN (fragment-1 descriptors sharded over the ranks, one all-gather of packed keys)
     equals the single-GPU result;
  2. configs 3-4 shape: a small scene sharded over the ranks returns, on every rank, the transforms of all pairs in order,
     equal to what one rank computes alone (the device draws are seeded per pair position).
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yoho_b200 import synth, dist as ydist
from yoho_b200.engine import get_engine
from yoho_b200.pipeline import PairPipeline
from yoho_b200.batch import register_scene

rank, local_rank, world = ydist.init_from_env()
torch.cuda.set_device(local_rank)
eng = get_engine(local_rank)
eng.load_part1(synth.synth_state_dict("PartI", 0))
eng.load_part2(synth.synth_state_dict("PartII", 0))
dev = eng.device

# ---- 1. sharded mutual NN (K = 10 000, config 5) -------------------------------------------------------------------
rs = np.random.RandomState(0)
Ka = Kb = 10000
dA = (rs.standard_normal((Ka, 32)) * 0.1).astype(np.float32)
dB = (rs.standard_normal((Kb, 32)) * 0.1).astype(np.float32)
dB[:4000] = dA[rs.permutation(Ka)[:4000]] + (rs.standard_normal((4000, 32)) * 0.01).astype(np.float32)
bounds = np.linspace(0, Kb, world + 1).astype(int)
lo, hi = int(bounds[rank]), int(bounds[rank + 1])
tA, tB = torch.from_numpy(dA).to(dev), torch.from_numpy(dB).to(dev)
got = ydist.sharded_mutual_nn(tA, tB[lo:hi].contiguous(), lo, Kb, eng.nn1)
pairs, n = eng.mutual_nn(tA, tB)
want = pairs[: int(n.item())]
ok1 = bool(torch.equal(got, want))

# ---- 2. scene sharding -----------------------------------------------------------------------------------------------
K = 600
frs = {}
base = synth.make_fragment_pair(K, seed=5, overlap=0.6)
frs[0] = (base["feat_A"], base["kps_A"])
frs[1] = (base["feat_B"], base["kps_B"])
p2 = synth.make_fragment_pair(K, seed=6, overlap=0.5)
frs[2] = (p2["feat_A"], p2["kps_A"])
frs[3] = (p2["feat_B"], p2["kps_B"])
pair_ids = [(0, 1), (2, 3), (0, 3), (2, 1), (1, 0)]
res = register_scene(PairPipeline(eng, seed=0), frs, pair_ids)
gather = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(gather, res)
else:
    gather = [res]
same_everywhere = all(bool(torch.equal(g, gather[0])) for g in gather)
Tc = res[0, 0].cpu().numpy()
cosang = (np.trace(Tc[:, :3].T @ base["R_gt"]) - 1) / 2
ok2 = same_everywhere and res.shape == (len(pair_ids), 2, 3, 4) and cosang > np.cos(np.deg2rad(3.0))
print(f"rank {rank}/{world}: sharded_mutual_nn == single-GPU: {ok1} (M={got.shape[0]}); scene gather consistent + planted rotation recovered: {ok2}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if (ok1 and ok2) else 1)
