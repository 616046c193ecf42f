"""Where does the dense-factor engine overtake the low-rank engine?  Kernel-only time of both on a candidate grid for a
small space (config C2: hex d=151) and a mid-size one (d=511) as the posterior rank grows.
Usage: python tools/rank_crossover.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib


def run(kind, ssp_dim, ranks, n=1 << 20):
    b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
    sp = (pkg.HexagonalSSPSpace(2, ssp_dim=ssp_dim, domain_bounds=b2, length_scale=1.0) if kind == "hex"
          else pkg.RandomSSPSpace(2, ssp_dim=ssp_dim, domain_bounds=b2, length_scale=1.0, rng=0))
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (max(ranks), 2))
    y = np.sin(X.sum(axis=1, keepdims=True))
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0)
    eng = agt._scoring_engine()
    cand = eng.fill_uniform(0, 0, n, 2)
    cand.mul_(10.0).sub_(5.0)
    done = 10
    for r in ranks:
        while done < r:
            agt.update(X[done:done + 1], y[done:done + 1], 0.0)
            done += 1
        row = [f"d={eng.d} rank {r:4d}"]
        for name, code in (("lowrank", _lib.ENGINE_UMMA_LR), ("dense", _lib.ENGINE_UMMA)):
            eng.score(cand, 10.0, 0.0, engine=code)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            for _ in range(5):
                eng.score(cand, 10.0, 0.0, engine=code)
            ev[1].record()
            torch.cuda.synchronize()
            row.append(f"{name} {ev[0].elapsed_time(ev[1]) / 5:.3f} ms")
        print("  ".join(row), flush=True)


if __name__ == "__main__":
    run("hex", 151, [40, 70, 76, 100, 130, 160, 200, 250])
    run("rand", 512, [100, 200, 250])
