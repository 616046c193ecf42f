"""CPU emulation of the fp16 + fp8 operand format of the scorer (d=1023, C3):
    Y = psi_h16 . V_h16  +  e4m3(psi) . e4m3(V_lo)  +  e4m3(psi_lo 2^s) . e4m3(V 2^-s)
against the round-1 three-term split and float64, for the low-rank form (var = 1/beta + 1/alpha - |V psi|^2) and the
factor form (var = 1/beta + |R psi|^2), candidates at every distance from the observations."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ssp_oracle as orc

def h16(a): return a.astype(np.float16).astype(np.float64)
def trunc13(a):
    f = a.astype(np.float32).view(np.uint32) & np.uint32(0xffffe000)
    return f.view(np.float32).astype(np.float64)
def e4m3(a):
    """round to nearest e4m3 (bias 7, 3 mantissa bits, subnormal spacing 2^-9, max 448, saturating)"""
    a = np.asarray(a, dtype=np.float64)
    s = np.sign(a); m = np.abs(a)
    e = np.floor(np.log2(np.maximum(m, 1e-300)))
    e = np.maximum(e, -6.0)                      # subnormal: fixed spacing 2^-9
    q = 2.0 ** (e - 3)
    r = np.round(m / q) * q                      # ties-to-even is close enough here
    return s * np.minimum(r, 448.0)

def posterior(rank, n=6, d_req=1024, ls=0.25, alpha=0.01, seed=1):
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
    cands = [rng.uniform(0, 1, (20000, n))]
    for rad in (1e-3, 3e-3, 1e-2, 2e-2, 3e-2, 5e-2, 8e-2, 0.12):
        cands.append(np.clip(X[rng.integers(0, rank, 2000)] + rng.normal(0, rad, (2000, n)), 0, 1))
    xc = np.vstack(cands)
    return dict(d=d, alpha=alpha, beta=beta, Sp=Sp, V=V, psi=psi_of(xc))

def contract(psi, W, s_lo=10, which="fp8lo"):
    """Y = psi W^T with W pre-scaled into fp16 range."""
    Wh = h16(W); Wl = W - Wh
    ph = trunc13(psi); pl = psi - ph
    if which == "3term":
        return ph @ Wh.T + ph @ h16(Wl).T + h16(pl) @ Wh.T
    if which == "fp8lo":
        return ph @ Wh.T + e4m3(psi) @ e4m3(Wl).T + e4m3(pl * 2.0 ** s_lo) @ e4m3(Wh * 2.0 ** -s_lo).T
    if which == "fp8lo-e":   # psi_hi by RN instead of truncation (lo symmetric about 0)
        ph = h16(psi); pl = psi - ph
        return ph @ Wh.T + e4m3(psi) @ e4m3(Wl).T + e4m3(pl * 2.0 ** s_lo) @ e4m3(Wh * 2.0 ** -s_lo).T
    raise ValueError(which)

def report(name, v, var_true, frac):
    rel = np.abs(v - var_true) / var_true
    line = f"{name:10s}"
    for lo, hi in ((0.0, 0.01), (0.01, 0.1), (0.1, 0.3), (0.3, 0.5), (0.5, 0.7), (0.7, 2.0)):
        msk = (frac >= lo) & (frac < hi)
        line += f" [{lo:.2f},{hi:.2f}) n={msk.sum():5d} max={rel[msk].max() if msk.any() else 0:.1e} |"
    print(line)

def main(rank):
    p = posterior(rank)
    d, alpha, beta, psi, V, Sp = p["d"], p["alpha"], p["beta"], p["psi"], p["V"], p["Sp"]
    rscale = 2.0 ** np.floor(np.log2(1024.0 / np.sqrt(2.0 / (d * alpha))))
    q_true = np.sum((psi @ V.T) ** 2, axis=1)
    var_true = 1 / beta + 1 / alpha - q_true
    frac = var_true * alpha
    print(f"== rank {rank}, low-rank form; max|V rscale| = {np.abs(V * rscale).max():.1f}")
    for which in ("3term", "fp8lo", "fp8lo-e"):
        Y = contract(psi, V * rscale, which=which)
        report(which, 1 / beta + 1 / alpha - np.sum(Y ** 2, axis=1) / rscale ** 2, var_true, frac)
    L = np.linalg.cholesky(Sp + 1e-300 * np.eye(d))          # psi^T Sp psi = |L^T psi|^2
    R = L.T
    qf_true = np.sum((psi @ R.T) ** 2, axis=1)
    varf_true = 1 / beta + qf_true
    print(f"== rank {rank}, factor form; |var_lr - var_f|/var max {np.max(np.abs(varf_true - var_true) / var_true):.1e}; max|R rscale| = {np.abs(R * rscale).max():.1f}")
    for which in ("3term", "fp8lo", "fp8lo-e"):
        Y = contract(psi, R * rscale, which=which)
        report(which, 1 / beta + np.sum(Y ** 2, axis=1) / rscale ** 2, varf_true, frac)

if __name__ == "__main__":
    for r in (60, 255, 511):
        main(r)
