"""Regret check of the two acquisition optimisers of `maximize` on Himmelblau (max = 0.07 at x = (3.58, -1.85) for the
sign / scale used here): fused candidate scoring ('candidates') against the phi-space ascent ('phi-ascent')."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssp_bayesopt_b200 as pkg

bounds = np.array([[-5.0, 5.0], [-5.0, 5.0]])


def f(x):
    x = np.atleast_2d(x)
    return -((x[:, 0] ** 2 + x[:, 1] - 11) ** 2 + (x[:, 0] + x[:, 1] ** 2 - 7) ** 2) / 100.0


xs0 = np.random.default_rng(0).uniform(-5, 5, (10, 2))
for mode in ("candidates", "phi-ascent"):
    for gamma_c in (0.0, 1.0):
        opt = pkg.BayesianOptimization(f=f, bounds=bounds, random_state=1)
        opt.maximize(init_points=(xs0, f(xs0).reshape(-1, 1)), n_iter=60, num_restarts=5, agent_type="ssp-hex", ssp_dim=151,
                     length_scale=1.0, decoder_method="from-set", acq_optimizer=mode, gamma_c=gamma_c, beta_ucb=10.0)
        best = np.maximum.accumulate(opt.ys)
        uniq = len(np.unique(np.round(opt.xs[10:], 6), axis=0))
        print(f"{mode:11s} gamma_c={gamma_c}: best after 10/30/60 its {best[19]:.4f} {best[39]:.4f} {best[69]:.4f}  "
              f"distinct points {uniq}/60  ms/step {opt.full_times[5:].mean() * 1e-6:.3f}")
