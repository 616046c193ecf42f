#!/usr/bin/env python
"""Kernel-only timing of the fused scorer on config C3 for several posterior ranks (CUDA events, device-resident
candidates).  The SSPBO_DEBUG switches need a library built with `make -C ssp_bayesopt_b200/csrc EXPERIMENTS=1`.
Usage: python tools/kbench.py [--ranks 60,130,260,500] [--n 4194304] [--engines lr,full]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", default="60,130,260,500")
    ap.add_argument("--n", type=int, default=1 << 22)
    ap.add_argument("--engines", default="lr,full")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import ssp_bayesopt_b200 as pkg
    from ssp_bayesopt_b200 import _lib
    from oracle import ssp_oracle as orc
    b6 = np.array([[0.0, 1.0]] * 6)
    ranks = [int(r) for r in args.ranks.split(",")]
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    X = np.random.default_rng(1).uniform(0, 1, (max(ranks), 6))
    y = orc.hartmann6(X).reshape(-1, 1)
    model = pkg.BayesianLinearRegression(sp.ssp_dim, engine=sp.engine)
    eng = sp.engine
    cand = eng.fill_uniform(0, 0, args.n)
    phis = sp.engine.encode_device(X)
    done = 0
    for r in ranks:
        if done == 0:
            model.update(phis[:10].cpu().numpy(), y[:10]); done = 10
        while done < r:
            model.update(phis[done:done + 1].cpu().numpy(), y[done:done + 1]); done += 1
        for name in args.engines.split(","):
            code = {"lr": _lib.ENGINE_UMMA_LR, "full": _lib.ENGINE_UMMA, "auto": _lib.ENGINE_AUTO}[name]
            eng.score(cand, 10.0, 0.0, engine=code)
            torch.cuda.synchronize()
            eng.score_stats(reset=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                eng.score(cand, 10.0, 0.0, engine=code)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            st = eng.score_stats(reset=True)
            print(f"rank {r:4d} engine {name:5s} {ms:8.3f} ms  {args.n / ms * 1e-6:8.4f} e9 cand/s  flagged/call {st['flagged'] / args.reps:.0f}  "
                  f"best {eng.best()}", flush=True)


if __name__ == "__main__":
    main()
