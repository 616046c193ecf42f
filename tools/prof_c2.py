"""One fused scoring launch of config C2's shape (hex n=2, d=151, 1e6-point grid, rank 70) for ncu.
Usage: python tools/prof_c2.py [--rank 70] [--gen]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg

ap = argparse.ArgumentParser()
ap.add_argument("--rank", type=int, default=70)
ap.add_argument("--gen", action="store_true")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
xs = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (args.rank, 2))
sp = pkg.HexagonalSSPSpace(2, ssp_dim=151, domain_bounds=bounds, length_scale=1.0)
agt = pkg.SSPAgent(xs[:10], branin(xs[:10]).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
agt.update(xs[10:], branin(xs[10:]).reshape(-1, 1), 0.0)
eng = agt._scoring_engine()
axes = [np.linspace(bounds[i, 0], bounds[i, 1], 1000) for i in range(2)]
grid = None if args.gen else eng.grid_points(axes)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
eng.set_timing(True)
for rep in range(args.reps):
    if rep == 1:
        eng.timing_read(); ev[0].record()
    if args.gen:
        eng.score_generated("grid", 1000000, 10.0, 0.0, axes=axes)
    else:
        eng.score(grid, 10.0, 0.0)
ev[1].record(); torch.cuda.synchronize()
print(f"C2 d={eng.d} rank {eng.lowrank_info()['rank']} engine {eng.last_engine()}: {ev[0].elapsed_time(ev[1]) / (args.reps - 1):.4f} ms per call, "
      f"timing {eng.timing_read()}, stats {eng.score_stats()}")
