"""Kernel-only rate (CUDA events, candidates resident or generated) of every scoring engine on config C3 (Random n=6, d=1023)
as the posterior rank grows: split-fp16 low-rank kernel, fp16 + e4m3 low-rank kernel (1 and 2 chunks), factor form, the older
dense kernel.  Usage: python tools/rank_sweep.py [--n 4194304] [--ranks 10,60,...] [--gen]"""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib


def hartmann6(x):
    A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14.0]])
    P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381.0]])
    al = np.array([1.0, 1.2, 3.0, 3.2])
    return np.sum(al * np.exp(-np.sum(A * (x[:, None, :] - P) ** 2, axis=2)), axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 22)
    ap.add_argument("--ranks", default="10,60,89,100,127,128,160,200,255,256,320,400,511,600,1000")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--gen", action="store_true", help="candidates generated inside the kernel instead of resident rows")
    ap.add_argument("--dense-at", default="255,1000", help="ranks at which the older dense kernel is timed too")
    args = ap.parse_args()
    ranks = [int(r) for r in args.ranks.split(",")]
    dense_at = {int(r) for r in args.dense_at.split(",") if r}
    b6 = np.array([[0.0, 1.0]] * 6)
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    X = np.random.default_rng(1).uniform(0, 1, (max(ranks), 6))
    y = hartmann6(X).reshape(-1, 1)
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
    eng = agt._scoring_engine()
    cand = None if args.gen else eng.fill_uniform(0, 0, args.n)

    def timed(code, fmt):
        eng.set_operand_format(fmt)
        def go():
            if args.gen:
                eng.score_generated("uniform", args.n, 10.0, 0.0, seed=0, engine=code)
            else:
                eng.score(cand, 10.0, 0.0, engine=code)
        try:
            go()
        except _lib.SspboError as e:
            return None
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.reps):
            go()
        ev[1].record()
        torch.cuda.synchronize()
        st = eng.score_stats(reset=True)
        eng.set_policy(lowrank=True)
        return args.n * args.reps / (ev[0].elapsed_time(ev[1]) * 1e-3), st["flagged"] / max(st["scored"], 1)

    done = 10
    out = []
    for r in ranks:
        while done < r:
            agt.update(X[done:done + 1], y[done:done + 1], 0.0)
            done += 1
        row = {"rank": r}
        for name, code, fmt in (("lr_fp16x3", _lib.ENGINE_UMMA_LR, 0), ("lr_f8", _lib.ENGINE_UMMA_LR, 1),
                                ("factor_f8", _lib.ENGINE_UMMA_F8, None), ("auto", _lib.ENGINE_AUTO, None)):
            if name == "lr_fp16x3" and r > 255 or name.startswith("lr") and r > 511:
                continue
            res = timed(code, fmt)
            if res:
                row[name] = round(res[0] / 1e9, 4)
                if name != "factor_f8":
                    row[name + "_flagged"] = round(res[1], 5)
                if name == "auto":
                    row["auto_engine"] = eng.last_engine()
        if r in dense_at:
            res = timed(_lib.ENGINE_UMMA, None)
            if res:
                row["dense_umma"] = round(res[0] / 1e9, 4)
        eng.set_operand_format(None)
        print(json.dumps(row), flush=True)
        out.append(row)


if __name__ == "__main__":
    main()
