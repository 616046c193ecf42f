"""Where does the end-to-end (host candidates) step of bench.py spend its time?  Config C3 at posterior rank ~100:
device-resident scoring against SSPAgent.best_candidate on a pinned host shard, for several piece sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from oracle import ssp_oracle as orc

b6 = np.array([[0.0, 1.0]] * 6)
sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
X = np.random.default_rng(1).uniform(0, 1, (100, 6))
agt = pkg.SSPAgent(X, orc.hartmann6(X).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
eng = agt._scoring_engine()
n = 12_500_000
cand = eng.fill_uniform(0, 0, n)
host = torch.empty((n, 6), dtype=torch.float32).pin_memory()
host.copy_(cand)


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def resident(chunk):
    def f():
        first = True
        for c0 in range(0, n, chunk):
            agt.best_candidate(cand[c0:c0 + chunk], index0=c0, reset_best=first)
            first = False
        eng.best()
    return f


def streamed(chunk, first=None, growth=2.0):
    def f():
        eng.score_host(host, agt.var_weight, 0.0, chunk=chunk, first_piece=first, growth=growth)
        eng.best()
    return f


for chunk in (1 << 22, 1 << 21, 1 << 20):
    print(f"resident, launches of {chunk:8d}: {timed(resident(chunk)):.2f} ms")
for chunk in (1 << 22, 1 << 21, 1 << 20, 1 << 19):
    print(f"streamed, pieces up to {chunk:8d}: {timed(streamed(chunk)):.2f} ms")
for chunk, first, growth in ((1 << 21, 1 << 19, 1.5), (1 << 22, 1 << 19, 1.5), (1 << 22, 1 << 20, 1.4), (3 << 20, 1 << 19, 1.6),
                             (1 << 21, 1 << 20, 1.4), (1 << 22, 1 << 18, 1.5)):
    print(f"streamed, pieces up to {chunk:8d}, first {first:8d}, growth {growth}: {timed(streamed(chunk, first, growth)):.2f} ms")
