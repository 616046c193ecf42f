"""Small end-to-end run of every scoring engine, the tensor-core encoder, the device decoder and the phi-space ascent (quick manual check)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib

b = np.array([[0.0, 1.0]] * 3)
sp = pkg.RandomSSPSpace(3, ssp_dim=200, domain_bounds=b, length_scale=0.4, rng=0)
rng = np.random.default_rng(0)
X = rng.uniform(0, 1, (20, 3))
agt = pkg.SSPAgent(X, np.sin(X.sum(axis=1, keepdims=True)), sp, gamma_c=0.0, beta_ucb=5.0, length_scale=0.4)
xc = rng.uniform(0, 1, (1500, 3)).astype(np.float32)
xc[:4] = X[:4]
eng = agt._scoring_engine()
for code in (_lib.ENGINE_UMMA_LR, _lib.ENGINE_UMMA, _lib.ENGINE_SIMT, _lib.ENGINE_EXACT):
    eng.score(xc if code != _lib.ENGINE_EXACT else xc[:40], 5.0, 0.0, engine=code, want_outputs=True)
    print(eng.last_engine(), eng.best())
agt.update(xc[:1].astype(np.float64), np.array([[0.1]]), np.array([0.0]))
print("f32 encode", sp.encode(xc[:300], dtype=np.float32).shape, "decode", sp.decode(sp.encode(xc[:3]), method="direct-optim"))
phis, vals = agt.maximize_acquisition(num_restarts=3, rng=0)
print("phi-space ascent", phis.shape, vals, "->", agt.decode(phis[int(np.argmax(vals))][None]))
