"""Time the phi-space acquisition ascent kernel (K6) alone: microseconds per fixed-point step for several ssp_dim,
on posteriors over raw features (a few observations).  Usage: python tools/ascent_bench.py [restarts]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssp_bayesopt_b200 as pkg  # noqa: E402


def main():
    import torch
    restarts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    rng = np.random.default_rng(0)
    for d in (151, 511, 1023, 1945, 2047):
        model = pkg.BayesianLinearRegression(d)
        X = rng.normal(size=(20, d)) / np.sqrt(d)
        model.update(X, X[:, :1] + 0.1 * rng.normal(size=(20, 1)))
        eng = model.engine
        p0 = eng.to_device(rng.normal(size=(restarts, d)), torch.float64)
        out = {}
        for steps in (20, 220):
            eng.acq_ascent(p0, 10.0, 1.0, max_iter=steps, tol=0.0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(5):
                eng.acq_ascent(p0, 10.0, 1.0, max_iter=steps, tol=0.0)
            ev[1].record()
            torch.cuda.synchronize()
            out[steps] = ev[0].elapsed_time(ev[1]) / 5 * 1e3
        per_step = (out[220] - out[20]) / 200
        print(f"d={d:5d} restarts={restarts}: {per_step:7.2f} us/step, launch+setup {out[20] - 20 * per_step:7.1f} us, "
              f"S re-read {8 * d * d / per_step / 1e3:7.1f} GB/s, f64 {2 * d * d * restarts / per_step / 1e3:7.1f} GFLOP/s")


if __name__ == "__main__":
    main()
