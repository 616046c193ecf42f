import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ssp_bayesopt_b200 as pkg
from ssp_bayesopt_b200 import _lib
from tests.conftest import himmelblau
g = np.load("tests/golden/agent.npz"); b = np.load("tests/golden/blr.npz")
B2 = np.array([[-5.0, 5.0], [-5.0, 5.0]])
sp = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=B2, length_scale=1.0, rng=0)
agt = pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0, var_decay=0.25, length_scale=1.0)
for i in range(6):
    x_t = g["xt"][i:i + 1]
    _, var_t, _ = agt.eval(x_t)
    agt.update(x_t, np.atleast_2d(himmelblau(x_t)), var_t, step_num=i)
eng = agt._scoring_engine()
print(eng.lowrank_info())
x = np.random.default_rng(3).uniform(-5, 5, (50001, 2)).astype(np.float32)
xd = eng.to_device(x)
for name, code in (("auto", _lib.ENGINE_AUTO), ("lr", _lib.ENGINE_UMMA_LR), ("umma", _lib.ENGINE_UMMA), ("simt", _lib.ENGINE_SIMT)):
    eng.score(xd, 10.0, 0.0, index0=7, engine=code)
    print(name, "resident", eng.best(), eng.last_engine(), eng.score_stats())
    res = []
    for c0 in range(0, 50001, 4096):
        eng.score(xd[c0:c0 + 4096], 10.0, 0.0, index0=7 + c0, engine=code if xd[c0:c0+4096].shape[0] >= 1024 or code != _lib.ENGINE_UMMA_LR else _lib.ENGINE_AUTO)
        res.append(eng.best())
    print(name, "chunks", [(round(s, 3), i) for s, i in res])
_, (mu, var, acq) = eng.score(xd[8192:12288], 10.0, 0.0, index0=0, engine=_lib.ENGINE_UMMA_LR, want_outputs=True)
_, (mu2, var2, acq2) = eng.score(xd[8192:12288], 10.0, 0.0, index0=0, engine=_lib.ENGINE_UMMA, want_outputs=True)
d = ((mu + acq) - (mu2 + acq2)).abs()
print("chunk2 lr-vs-umma max diff", d.max().item(), "at", d.argmax().item(), (mu+acq)[11906-8192].item(), (mu2+acq2)[11906-8192].item(), var[11906-8192].item(), var2[11906-8192].item())
eng.score_host(x, 10.0, 0.0, index0=7, chunk=4096)
print("score_host", eng.best())
print("---- second agent, test sequence")
sp = pkg.HexagonalSSPSpace(2, ssp_dim=51, domain_bounds=B2, length_scale=1.0, rng=0)
agt = pkg.SSPAgent(b["xs"], b["ys"], sp, decoder_method="from-set", gamma_c=0.0, beta_ucb=10.0, var_decay=0.25, length_scale=1.0)
for i in range(6):
    x_t = g["xt"][i:i + 1]
    _, var_t, _ = agt.eval(x_t)
    agt.update(x_t, np.atleast_2d(himmelblau(x_t)), var_t, step_num=i)
eng = agt._scoring_engine()
print(eng.lowrank_info())
eng.score(x, 10.0, 0.0, index0=7)
print("first", _lib.key_unpack(int(eng._best.item())), eng.last_engine(), eng.lowrank_info())
print(eng.score_stats(reset=False))
eng.score(x, 10.0, 0.0, index0=7)
print("second", _lib.key_unpack(int(eng._best.item())), eng.last_engine(), eng.lowrank_info())
_, (mu, var, acq) = eng.score(x, 10.0, 0.0, index0=7, want_outputs=True, engine=_lib.ENGINE_UMMA_LR)
_, (mu2, var2, acq2) = eng.score(x, 10.0, 0.0, index0=7, want_outputs=True, engine=_lib.ENGINE_UMMA)
print("lr 11906", mu[11906-7].item(), var[11906-7].item(), acq[11906-7].item(), "umma", mu2[11906-7].item(), var2[11906-7].item(), acq2[11906-7].item())
d = ((mu + acq) - (mu2 + acq2)).abs(); print("max diff", d.max().item(), d.argmax().item())
for rep in range(30):
    eng.score(x, 10.0, 0.0, index0=7, engine=_lib.ENGINE_UMMA)
    k = _lib.key_unpack(int(eng._best.item()))
    if k[1] != 33146: print("MISMATCH umma", rep, k)
    eng.score(x, 10.0, 0.0, index0=7, engine=_lib.ENGINE_UMMA_LR)
    k = _lib.key_unpack(int(eng._best.item()))
    if k[1] != 33146: print("MISMATCH lr", rep, k)
bad = torch.nonzero(d > 1e-3).flatten().cpu().numpy()
print("bad count", len(bad), bad[:40])
for i in bad[:6]:
    print(i, "lr", mu[i].item(), var[i].item(), acq[i].item(), "umma", mu2[i].item(), var2[i].item(), acq2[i].item())
print(eng.score_stats(reset=False))
