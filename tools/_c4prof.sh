C4_RUNS=256 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_c4_launches_b.csv python tools/configs_bench.py c4 > gpurun_out/r2_c4_ncu_b.log 2>&1
python profiles/launch_summary.py gpurun_out/r2_c4_launches_b.csv 1500
python - <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200.batched import BatchedRuns
from tools.configs_bench import himmelblau
b2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
agents = []
for r in range(256):
    sp = pkg.RandomSSPSpace(2, ssp_dim=512, domain_bounds=b2, length_scale=1.0, rng=r)
    X = np.random.default_rng(r).uniform(-5, 5, (10, 2))
    agents.append(pkg.SSPAgent(X, himmelblau(X).reshape(-1, 1), sp, gamma_c=0.0, beta_ucb=10.0, length_scale=1.0))
ax = np.linspace(-5, 5, 100)
grid = np.stack(np.meshgrid(ax, ax), axis=-1).reshape(-1, 2).astype(np.float32)
br = BatchedRuns(agents, grid)
br.step(himmelblau); torch.cuda.synchronize()
ts = {"select": 0.0, "obj": 0.0, "update": 0.0}
for _ in range(5):
    t0 = time.perf_counter(); scores, idx, pts = br.select(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ys = himmelblau(pts); t2 = time.perf_counter()
    br.update(pts, ys); torch.cuda.synchronize(); t3 = time.perf_counter()
    ts["select"] += t1 - t0; ts["obj"] += t2 - t1; ts["update"] += t3 - t2
print({k: round(v / 5 * 1e3, 3) for k, v in ts.items()}, "ms per lock-step of 256 runs")
PY
