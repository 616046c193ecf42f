"""Timeline of one C3 bench step (bench.py's bo_step): every device activity with start / duration, to see where the GPU idles
between the last scorer launch of one step and the first of the next."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import dist
from torch.profiler import profile, ProfilerActivity

X, y, orc = bench.problem_setup()
b6 = np.array([[0.0, 1.0]] * bench.N_DIM)
sp = pkg.RandomSSPSpace(bench.N_DIM, ssp_dim=bench.D_REQ, domain_bounds=b6, length_scale=bench.LS, rng=0)
agt = pkg.SSPAgent(X[:10], y[:10], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=bench.LAM, length_scale=bench.LS)
agt.update(X[10:], y[10:], 0.0)
eng = agt._scoring_engine()
n_local, chunk = bench.PER_GPU, 1 << 22
cand = eng.fill_uniform(0, 0, n_local)
saved, beta0 = eng.blr_export(), eng.blr_beta()


def step():
    first = True
    for c0 in range(0, n_local, chunk):
        agt.best_candidate(cand[c0:c0 + chunk], index0=c0, reset_best=first)
        first = False
    score, gidx = dist.reduce_best(eng)
    x_t = orc.counter_uniform(0, gidx, 1, bench.N_DIM).astype(np.float64)
    agt.update(x_t, np.atleast_2d(orc.hartmann6(x_t)), np.array([0.0]))
    eng.blr_import(saved, beta0)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
prev_end = None
scorers = [e for e in ev if "k_score" in e.name]
lo, hi = scorers[3].time_range.start, scorers[6].time_range.start          # second step: from its first scorer to the next step's
for e in ev:
    if lo <= e.time_range.start < hi:
        gap = 0.0 if prev_end is None else (e.time_range.start - prev_end)
        print(f"{(e.time_range.start - lo):9.1f} us  +{gap:7.1f} idle  {e.time_range.end - e.time_range.start:8.1f} us  {e.name[:90]}")
        prev_end = max(prev_end or 0, e.time_range.end)
print(f"step: {(hi - lo):.1f} us")
