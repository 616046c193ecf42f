"""One scoring launch of a chosen engine / operand format at a chosen posterior rank on config C3 (Random n=6, d=1023), for
`ncu -k regex:k_score -c 1 ...`.  Usage: python tools/prof_one.py --rank 255 --fmt 1 [--engine lr|f8|umma|auto] [--gen] [--reps 2]"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, default=255)
    ap.add_argument("--n", type=int, default=1 << 22)
    ap.add_argument("--fmt", type=int, default=-1)
    ap.add_argument("--engine", default="lr")
    ap.add_argument("--gen", action="store_true")
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    from oracle import ssp_oracle as orc
    b6 = np.array([[0.0, 1.0]] * 6)
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    X = np.random.default_rng(1).uniform(0, 1, (max(args.rank, 10), 6))
    y = orc.hartmann6(X).reshape(-1, 1)
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
    eng = agt._scoring_engine()
    for i in range(10, args.rank):
        agt.update(X[i:i + 1], y[i:i + 1], 0.0)
    code = {"lr": _lib.ENGINE_UMMA_LR, "f8": _lib.ENGINE_UMMA_F8, "umma": _lib.ENGINE_UMMA, "auto": _lib.ENGINE_AUTO}[args.engine]
    eng.set_operand_format(None if args.fmt < 0 else args.fmt)
    cand = None if args.gen else eng.fill_uniform(0, 0, args.n)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(args.reps):
        if rep == args.reps - 1:
            ev[0].record()
        if args.gen:
            eng.score_generated("uniform", args.n, 10.0, 0.0, seed=0, engine=code)
        else:
            eng.score(cand, 10.0, 0.0, engine=code)
    ev[1].record()
    torch.cuda.synchronize()
    print(f"rank {args.rank} engine {eng.last_engine()} fmt {args.fmt}: {ev[0].elapsed_time(ev[1]):.3f} ms per launch of {args.n}, "
          f"{args.n / ev[0].elapsed_time(ev[1]) * 1e-6:.3f}e9 cand/s, stats {eng.score_stats(reset=True)}")


if __name__ == "__main__":
    main()
