"""Config C4 lock-step: where the time goes (host call time and GPU time of the grouped selection and of the batched update)."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib
from ssp_bayesopt_b200.batched import BatchedRuns
from tools.configs_bench import himmelblau

R = int(os.environ.get("C4_RUNS", "256"))
b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
agents = []
for r in range(R):
    sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=b2, length_scale=1.0, rng=r)
    X = np.random.default_rng(r).uniform(-5, 5, (10, 2))
    agents.append(pkg.SSPAgent(X, himmelblau(X).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
ax = np.linspace(-5, 5, 100)
grid = np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2).astype(np.float32)
br = BatchedRuns(agents, grid)
for _ in range(3):
    br.step(himmelblau)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = {"select_wall": 0.0, "select_gpu": 0.0, "update_wall": 0.0, "update_gpu": 0.0}
steps = 10
for _ in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ev[0].record()
    scores, idx, pts = br.select()
    ev[1].record(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ys = himmelblau(pts)
    t2 = time.perf_counter(); ev[2].record()
    br.update(pts, ys)
    ev[3].record(); torch.cuda.synchronize(); t3 = time.perf_counter()
    acc["select_wall"] += t1 - t0; acc["update_wall"] += t3 - t2
    acc["select_gpu"] += ev[0].elapsed_time(ev[1]) * 1e-3; acc["update_gpu"] += ev[2].elapsed_time(ev[3]) * 1e-3
print({k: round(v / steps * 1e3, 3) for k, v in acc.items()}, f"ms per lock-step of {R} runs; rank {agents[0]._scoring_engine().lowrank_info()['rank']}")

if os.environ.get("C4_PROFILE"):
    # kernel-level view of one lock-step (CUPTI through torch.profiler): which launches make up the selection / update
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        scores, idx, pts = br.select()
        br.update(pts, himmelblau(pts))
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
    rows = [r for r in rows if r[1] > 0]
    for name, us, cnt in sorted(rows, key=lambda r: -r[1])[:14]:
        print(f"{us / 1e3:9.3f} ms  x{cnt:<5d} {name[:110]}")
