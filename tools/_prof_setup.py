import os, sys, time, cProfile, pstats
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import ssp_bayesopt_b200 as pkg
def himmelblau(x):
    return -((x[:, 0] ** 2 + x[:, 1] - 11) ** 2 + (x[:, 0] + x[:, 1] ** 2 - 7) ** 2) / 100.0
b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
def make(n):
    out = []
    for r in range(n):
        sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=b2, length_scale=1.0, rng=r)
        X = np.random.default_rng(r).uniform(-5, 5, (10, 2))
        out.append(pkg.SSPAgent(X, himmelblau(X).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
    return out
make(4)
pr = cProfile.Profile(); pr.enable(); a = make(64); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
