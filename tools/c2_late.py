"""Config C2 late in its 250-iteration run (posterior rank 130 .. 260 at d = 151): which kernels make up a BO step once the
low-rank form flags most of a well-explored 2-D domain.  Usage: python tools/c2_late.py [rank ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from torch.profiler import profile, ProfilerActivity

bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
ranks = [int(v) for v in sys.argv[1:]] or [60, 140, 200, 250]
xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
opt = pkg.BayesianOptimization(f=branin, bounds=bounds)
done = 10
for target in ranks:
    opt2 = pkg.BayesianOptimization(f=branin, bounds=bounds)
    opt2.maximize(init_points=(xs0, branin(xs0).reshape(-1, 1)), n_iter=target - 10 + 6, agent_type="ssp-hex", ssp_dim=151,
                  length_scale=1.0, beta_ucb=10.0, gamma_c=0.0, candidates_per_dim=1000)
    eng = opt2.agt._scoring_engine()
    ft = opt2.full_times[-5:] * 1e-6
    print(f"rank {target}: {np.mean(ft):.3f} ms per step over the last 5; engine {eng.last_engine()}, stats of the last step {eng.score_stats(reset=False)}, "
          f"lowrank {eng.lowrank_info()}")
    cand = opt2._cand if hasattr(opt2, "_cand") else None
    grid = eng.grid_points([np.linspace(bounds[i, 0], bounds[i, 1], 1000) for i in range(2)], 0, 1000000)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            opt2.agt.best_candidate(grid)
            sc, idx = eng.best()
            x = grid[idx:idx + 1].double().cpu().numpy()
            opt2.agt.observe(x, np.atleast_2d(branin(x)), step_num=0)
        torch.cuda.synchronize()
    rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
    for name, us, cnt in sorted(rows, key=lambda q: -q[1])[:9]:
        print(f"   {us / 3e3:8.4f} ms/step  x{cnt // 3:<3d} {name[:100]}")
