"""CTA-pair (cta_group::2) variant of the fp16 + e4m3 scorer against the single-CTA variant: same outputs, and the rate of
both at several ranks (config C3).  Usage: python tools/pair_check.py [--ranks 127,200,255,400,511] [--n 4194304]"""
import argparse, os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib
from oracle import ssp_oracle as orc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", default="60,127,200,255,400,511")
    ap.add_argument("--n", type=int, default=1 << 22)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--factor", action="store_true")
    args = ap.parse_args()
    ranks = [int(r) for r in args.ranks.split(",")]
    b6 = np.array([[0.0, 1.0]] * 6)
    sp = pkg.RandomSSPSpace(6, ssp_dim=1024, domain_bounds=b6, length_scale=0.25, rng=0)
    X = np.random.default_rng(1).uniform(0, 1, (max(ranks), 6))
    y = orc.hartmann6(X).reshape(-1, 1)
    agt = pkg.SSPAgent(X[:10], y[:10], sp, gamma_c=0.0, beta_ucb=10.0, length_scale=0.25)
    eng = agt._scoring_engine()
    cand = eng.fill_uniform(0, 0, args.n)
    small = cand[:20000 + 77]
    done = 10

    def timed(code, fmt):
        eng.set_operand_format(fmt)
        eng.score(cand, 10.0, 0.0, engine=code)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.reps):
            eng.score(cand, 10.0, 0.0, engine=code)
        ev[1].record()
        torch.cuda.synchronize()
        eng.score_stats(reset=True)
        eng.set_policy(lowrank=True)
        return round(args.n * args.reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9, 4)

    for r in ranks:
        while done < r:
            agt.update(X[done:done + 1], y[done:done + 1], 0.0)
            done += 1
        row = {"rank": r}
        for name, code in (("lr", _lib.ENGINE_UMMA_LR),) + ((("factor", _lib.ENGINE_UMMA_F8),) if args.factor else ()):
            outs = {}
            for fmt in (1, 2):
                eng.set_operand_format(fmt)
                _, o = eng.score(small, 10.0, 0.0, want_outputs=True, engine=code)
                outs[fmt] = [t.double().cpu().numpy() for t in o] + [eng.best()]
            for i, what in enumerate(("mu", "var", "acq")):
                a, b = outs[1][i], outs[2][i]
                row[f"{name}_{what}_maxrel"] = float(np.max(np.abs(a - b) / np.maximum(np.abs(a), 1e-30)))
            row[f"{name}_same_best"] = outs[1][3][1] == outs[2][3][1]
            row[f"{name}_single"] = timed(code, 1)
            row[f"{name}_pair"] = timed(code, 2)
            if name == "lr" and r <= 255:
                row["lr_fp16x3"] = timed(code, 0)
        eng.set_operand_format(None)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
