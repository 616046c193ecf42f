"""CPU emulation of the low-rank scorer's contraction formats (d=1023, C3): error of var = 1/beta + 1/alpha - |V psi|^2
against float64 for candidates at every distance from the observations.  Formats:
  3term : psi(hi+lo) x V(hi+lo), dropping lo*lo            (round-1 kernel)
  psi1  : psi rounded to fp16 (RN) x V(hi+lo)              (2 units)
  v1    : psi(hi+lo) x V_hi                                (2 units)
  one   : psi_hi x V_hi                                    (1 unit)
fp32 accumulation is emulated in float64 (the tensor core's accumulate error is measured separately on the GPU)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ssp_oracle as orc

def h16(a): return a.astype(np.float16).astype(np.float64)
def trunc13(a):
    f = a.astype(np.float32).view(np.uint32) & np.uint32(0xffffe000)
    return f.view(np.float32).astype(np.float64)

def main(rank=255, n=6, d_req=1024, ls=0.25, alpha=0.01, seed=1):
    A = orc.random_phase_matrix(n, d_req, rng=0)
    d = A.shape[0]; K = (d - 1) // 2
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, (rank, n)); y = orc.hartmann6(X)
    inv_ls = np.ones(n) / ls
    def psi_of(x):
        th = (x * inv_ls) @ A[1:K + 1].T
        return np.concatenate([np.ones((x.shape[0], 1)), np.cos(th), np.sin(th)], axis=1)
    Dg = np.concatenate([[1.0 / d], np.full(2 * K, 2.0 / d)])
    beta = 1.0 / np.var(y[:10])
    Sp = np.diag(Dg / alpha)
    V = np.zeros((rank, d))
    P = psi_of(X)
    for t in range(rank):
        yv = Sp @ P[t]
        c = beta / (1 + beta * P[t] @ yv)
        V[t] = yv * np.sqrt(c)
        Sp -= c * np.outer(yv, yv)
    # candidates: uniform + perturbations of observations at several radii
    cands = [rng.uniform(0, 1, (20000, n))]
    for rad in (1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2, 8e-2, 0.12):
        cands.append(np.clip(X[rng.integers(0, rank, 3000)] + rng.normal(0, rad, (3000, n)), 0, 1))
    xc = np.vstack(cands)
    psi = psi_of(xc)
    q_true = np.sum((psi @ V.T) ** 2, axis=1)
    var_true = 1 / beta + 1 / alpha - q_true
    rscale = 2.0 ** np.floor(np.log2(1024.0 / np.sqrt(2.0 / (d * alpha))))
    Vs = V * rscale
    Vh = h16(Vs); Vl = h16(Vs - Vh)
    ph_t = trunc13(psi); pl_t = h16(psi - ph_t)       # kernel's truncation split
    ph_r = h16(psi)                                   # RN single
    def q_of(Y): return np.sum(Y ** 2, axis=1) / rscale ** 2
    schemes = {
        "3term": ph_t @ Vh.T + ph_t @ Vl.T + pl_t @ Vh.T,
        "psi1(RN)xV(hi+lo)": ph_r @ (Vh + Vl).T,
        "psi(hi+lo)xVhi": (ph_t + pl_t) @ Vh.T,
        "one: psi1(RN)xVhi": ph_r @ Vh.T,
    }
    frac = var_true * alpha
    print(f"rank {rank} d {d}: var/prior quantiles", np.quantile(frac, [0, .01, .1, .5, .9, 1]).round(4))
    for name, Y in schemes.items():
        v = 1 / beta + 1 / alpha - q_of(Y)
        rel = np.abs(v - var_true) / var_true
        line = f"{name:22s}"
        for lo, hi in ((0.0, 0.1), (0.1, 0.3), (0.3, 0.5), (0.5, 0.7), (0.7, 0.9), (0.9, 2.0)):
            msk = (frac >= lo) & (frac < hi)
            line += f" [{lo:.1f},{hi:.1f}) n={msk.sum():5d} max={rel[msk].max() if msk.any() else 0:.1e} |"
        print(line)
    for thr in (0.3, 0.5, 0.7):
        print(f"  flagged at {thr}: uniform {np.mean(frac[:20000] < thr) * 100:.3f}%")

if __name__ == "__main__":
    for r in (60, 255):
        main(rank=r)
