"""Config C2 (Branin, ssp-hex d=151, 1e6-point grid): time per BO step of BayesianOptimization.maximize, with a breakdown."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 60
opt = pkg.BayesianOptimization(f=branin, bounds=bounds)
opt.maximize(init_points=(xs0, branin(xs0).reshape(-1, 1)), n_iter=n_iter, agent_type="ssp-hex", ssp_dim=151,
             length_scale=1.0, beta_ucb=10.0, gamma_c=0.0, candidates_per_dim=1000)
ft = opt.full_times[5:] * 1e-6
tt = opt.times[5:] * 1e-6
eng = opt.agt._scoring_engine()
print(f"ms per BO step mean {ft.mean():.3f} min {ft.min():.3f}  (acquisition part {tt.mean():.3f})  d={eng.d} engine {eng.last_engine()} "
      f"rank {eng.lowrank_info()} best {opt.max['target']:.4f}")
print("per-step ms:", np.round(ft[:50], 3))
