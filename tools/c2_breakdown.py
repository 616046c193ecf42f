"""Config C2 BO step: wall-clock breakdown of one step into its host-visible segments, and GPU-only times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
bounds = np.array([[-5.0, 10.0], [0.0, 15.0]])
a, b, c, r, s, t = 1, 5.1 / (4 * np.pi ** 2), 5 / np.pi, 6., 10., 1 / (8 * np.pi)
branin = lambda x: -(a * (x[:, 1] - b * x[:, 0] ** 2 + c * x[:, 0] - r) ** 2 + s * (1 - t) * np.cos(x[:, 0]) + s)
xs0 = np.random.default_rng(0).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
opt = pkg.BayesianOptimization(f=branin, bounds=bounds)
opt.maximize(init_points=(xs0, branin(xs0).reshape(-1, 1)), n_iter=20, agent_type="ssp-hex", ssp_dim=151,
             length_scale=1.0, beta_ucb=10.0, gamma_c=0.0, candidates_per_dim=1000)
agt = opt.agt; eng = agt._scoring_engine()
cand = opt._build_candidates(agt, 'grid', 1000, None, 0)
T = {k: [] for k in ("select", "target", "eval", "update", "sync", "score_call", "settle", "best", "lookup")}
for it in range(40):
    t0 = time.perf_counter()
    key = agt.best_candidate(cand['x'], index0=0, reset_best=True)
    t1 = time.perf_counter()
    eng.settle()
    t2 = time.perf_counter()
    score, gidx = eng.best()
    t3 = time.perf_counter()
    x_t = np.asarray(cand['lookup'](gidx), dtype=np.float64).reshape(1, -1)
    t4 = time.perf_counter()
    y_t = np.atleast_2d(branin(x_t))
    t5 = time.perf_counter()
    mu_t, var_t, phi_t = agt.eval(x_t)
    t6 = time.perf_counter()
    agt.update(x_t, y_t, var_t, step_num=it)
    t7 = time.perf_counter()
    torch.cuda.synchronize()
    t8 = time.perf_counter()
    for k, v in zip(("score_call", "settle", "best", "lookup", "target", "eval", "update", "sync"),
                    (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6, t8 - t7)):
        T[k].append(v * 1e6)
for k in ("score_call", "settle", "best", "lookup", "target", "eval", "update", "sync"):
    print(f"{k:10s} {np.mean(T[k][5:]):8.1f} us")
print("total", sum(np.mean(T[k][5:]) for k in ("score_call", "settle", "best", "lookup", "target", "eval", "update", "sync")))
# GPU-only time of the scoring for 1e6 candidates
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record()
for _ in range(10): agt.best_candidate(cand['x'], index0=0, reset_best=True)
ev[1].record(); torch.cuda.synchronize()
print("score GPU ms per call", ev[0].elapsed_time(ev[1]) / 10, eng.last_engine(), eng.lowrank_info())
xs = np.random.default_rng(1).uniform(bounds[:, 0], bounds[:, 1], (10, 2))
torch.cuda.synchronize(); ev[0].record()
for i in range(10): agt.update(xs[i:i+1], np.atleast_2d(branin(xs[i:i+1])), np.array([0.0]))
ev[1].record(); torch.cuda.synchronize()
print("update GPU ms per call", ev[0].elapsed_time(ev[1]) / 10)
# kernel-level split of the scoring call (library events) and of the generated-grid route
eng.set_timing(True)
for _ in range(10): agt.best_candidate(cand['x'], index0=0, reset_best=True)
print("timing (10 calls, resident grid):", eng.timing_read())
axes = [np.linspace(bounds[i, 0], bounds[i, 1], 1000) for i in range(2)]
for _ in range(3): eng.score_generated("grid", 1000000, 10.0, 0.0, axes=axes)
eng.timing_read()
torch.cuda.synchronize(); ev[0].record()
for _ in range(10): eng.score_generated("grid", 1000000, 10.0, 0.0, axes=axes)
ev[1].record(); torch.cuda.synchronize()
print("generated-grid score GPU ms per call", ev[0].elapsed_time(ev[1]) / 10, "timing:", eng.timing_read())
