"""Key reduction when a rank cannot map its peers (every process sees only its own GPU through CUDA_VISIBLE_DEVICES): the
peer-memory exchange must decline on ALL ranks and dist.reduce_best fall back to the NCCL all-reduce, with the same winner.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_fallback_check.py [isolate]"""
import os, sys
if len(sys.argv) > 1 and sys.argv[1] == "isolate":
    os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("LOCAL_RANK", "0")
    local = 0
else:
    local = int(os.environ.get("LOCAL_RANK", "0"))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import dist

torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = td.get_rank(), td.get_world_size()
bounds = np.array([[-2.0, 2.0], [-2.0, 2.0]])
sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=0.7, device=local)
rng = np.random.default_rng(3)
X = rng.uniform(-2, 2, (15, 2)); y = -np.sum(X ** 2, axis=1, keepdims=True)
agt = pkg.SSPAgent(X, y, sp, length_scale=0.7, gamma_c=0.0, beta_ucb=5.0)
eng = agt._scoring_engine()
cand = rng.uniform(-1.9, 1.9, (40000, 2)).astype(np.float32)
s, e = dist.shard_range(len(cand), rank, world)
on = eng.dist_connect()
for step in range(3):
    agt.best_candidate(cand[s:e], index0=s)
    score, gidx = dist.reduce_best(eng)
    agt.best_candidate(cand)                                   # the whole set on every rank: the reference answer
    want = eng.best()
    assert (score, gidx) == want, (rank, score, gidx, want)
    agt.update(cand[gidx:gidx + 1].astype(np.float64), np.atleast_2d(-np.sum(cand[gidx] ** 2)), 0.0)
if rank == 0:
    print(f"world {world}: peer-memory exchange {'on' if on else 'declined -> NCCL all-reduce'}; winners equal the single-set arg-max in 3 steps")
td.destroy_process_group()
