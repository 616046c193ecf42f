#!/usr/bin/env python
"""Side measurements of BASELINE configs C4 (batched independent runs) and C5 (trajectory agent) at reduced scale."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200.batched import BatchedRuns


def himmelblau(x):
    return -((x[:, 0] ** 2 + x[:, 1] - 11) ** 2 + (x[:, 0] + x[:, 1] ** 2 - 7) ** 2) / 100.0


def c4(runs=int(os.environ.get("C4_RUNS", "4096")), steps=5):
    """BASELINE config C4: `runs` independent BLR-UCB runs (Himmelblau 2-D, ssp-rand ssp_dim=512 -> d=511, own space seed and
    posterior each) over a shared 100 x 100 grid, stepped in lock-step through BatchedRuns (grouped kernels)."""
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    agents = []
    t0 = time.perf_counter()
    for r in range(runs):
        sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=b2, length_scale=1.0, rng=r)
        X = np.random.default_rng(r).uniform(-5, 5, (10, 2))
        agents.append(pkg.SSPAgent(X, himmelblau(X).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
    t_build = time.perf_counter() - t0
    ax = np.linspace(-5, 5, 100)
    grid = np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2).astype(np.float32)
    br = BatchedRuns(agents, grid)
    br.step(himmelblau)
    torch.cuda.synchronize()
    t_sel = t_upd = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        ta = time.perf_counter()
        scores, idx, pts = br.select()
        tb = time.perf_counter()
        ys = himmelblau(pts)
        tc = time.perf_counter()
        br.update(pts, ys)
        torch.cuda.synchronize()
        t_sel += tb - ta; t_upd += time.perf_counter() - tc
    dt = (time.perf_counter() - t0) / steps
    eng = agents[0]._scoring_engine()
    print(f"C4: {runs} runs x {grid.shape[0]} candidates d={eng.d}: {dt * 1e3:.1f} ms per lock-step "
          f"({runs * grid.shape[0] / dt:.3e} candidate evaluations/s; selection {t_sel / steps * 1e3:.1f} ms, updates "
          f"{t_upd / steps * 1e3:.1f} ms; engine {eng.last_engine()}; set-up {t_build:.1f} s)")
    return {"workload": f"C4: {runs} independent BLR-UCB runs (Himmelblau 2-D, ssp-rand ssp_dim=512 -> d={eng.d}), shared 100x100 grid, "
                        "stepped in lock-step (BatchedRuns, grouped kernels)", "runs": runs, "candidates_per_run": int(grid.shape[0]),
            "steps": steps, "ms_per_lockstep": dt * 1e3, "selection_ms": t_sel / steps * 1e3, "updates_ms": t_upd / steps * 1e3,
            "candidate_evaluations_per_s": runs * grid.shape[0] / dt, "engine": eng.last_engine(),
            "posterior_rank": eng.lowrank_info()["rank"], "setup_s": t_build}


def c5(n=1 << 20, steps=3):
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    L, xd = 10, 2
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (12, L * xd))
    y = -np.sum(X ** 2, axis=1, keepdims=True) / 100.0
    agt = pkg.SSPTrajectoryAgent(X, y, x_dim=xd, traj_len=L, domain_bounds=b2, ssp_dim=2048, length_scale=1.0,
                                 time_length_scale=1.0, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0)
    eng = agt._scoring_engine()
    cand = eng.fill_uniform(0, 0, n, L * xd)
    cand.mul_(10.0).sub_(5.0)
    agt.best_candidate(cand)
    eng.best()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        agt.best_candidate(cand)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"C5: trajectory L={L} x_dim={xd} d={eng.d}: {n} candidates in {ms:.1f} ms = {n / ms * 1e3:.3e} candidates/s "
          f"(engine {eng.last_engine()}) -> 1e7 candidates: {1e7 / (n / ms * 1e3):.2f} s")
    return {"workload": f"C5: SSPTrajectoryAgent L={L} x_dim={xd} ssp_dim=2048 -> d={eng.d}, uniform trajectories, rank-12 posterior",
            "candidates": n, "steps": steps, "ms_per_pass": ms, "candidates_per_s": n / ms * 1e3, "engine": eng.last_engine(),
            "s_per_1e7_candidates": 1e7 / (n / ms * 1e3)}


def multi(n=1 << 20, steps=3):
    """Multi-agent trajectories: 3 agents x 10 timesteps, d = 1945, normalised scoring."""
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    n_agents, L, xd = 3, 10, 2
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (12, L * n_agents * xd))
    y = -np.sum(X ** 2, axis=1, keepdims=True) / 100.0
    agt = pkg.SSPMultiAgent(X, y, n_agents=n_agents, x_dim=xd, traj_len=L, domain_bounds=b2, ssp_dim=2048, length_scale=1.0,
                            time_length_scale=1.0, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0)
    eng = agt._scoring_engine()
    cand = eng.fill_uniform(0, 0, n, L * n_agents * xd)
    cand.mul_(10.0).sub_(5.0)
    agt.best_candidate(cand)
    eng.best()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        agt.best_candidate(cand)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    t0 = time.perf_counter()
    phi = agt.encode(cand[:4096].cpu().numpy().astype(np.float64))
    t_enc = time.perf_counter() - t0
    print(f"multi-agent {n_agents} x L={L} x_dim={xd} d={eng.d}: {n} candidates in {ms:.1f} ms = {n / ms * 1e3:.3e} candidates/s "
          f"(engine {eng.last_engine()}, stats {eng.score_stats(reset=False)}); float64 encode of 4096: {t_enc * 1e3:.1f} ms")
    return {"workload": f"multi-agent trajectories: SSPMultiAgent {n_agents} agents x L={L} x_dim={xd} ssp_dim=2048 -> d={eng.d}, "
                        "normalised scoring", "candidates": n, "steps": steps, "ms_per_pass": ms,
            "candidates_per_s": n / ms * 1e3, "engine": eng.last_engine(), "f64_encode_ms_per_4096": t_enc * 1e3}


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c5", "multi"]
    if "multi" in which:
        multi()
    if "c5" in which:
        c5()
    if "c4" in which:
        c4()
