"""numpy float64 restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy float64, the arithmetic the reference
(ctn-waterloo/ssp-bayesopt, pure Python/numpy) performs on the path
    candidates -> SSP encode -> BLR predict -> MI/UCB acquisition -> argmax
    -> from-set decode -> rank-one posterior update.
It is the *checker* for the CUDA path; it is never imported by the product
package and is never the thing being measured, except as the CPU baseline leg
of bench.py.

Parity status: PINNED.  The reference ships no golden vectors for this path
(SURVEY.md section 8c), so the oracle is pinned against outputs of the live
reference imported in the build container: `oracle/gen_golden.py` writes
`tests/golden/*.npz`, and `tests/test_oracle_golden.py` checks every function
here against them (float64, rtol 1e-12 or bit-exact where stated).

Third-party arithmetic on the reference path lives in numpy (pocketfft `ifft`,
complex128 `exp`, OpenBLAS `@`), pinned by the reference at numpy>=1.22
(`pyproject.toml:21-26`); numpy 2.3.5 is installed here.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------
# phase matrices (host-side setup; must reproduce the reference's RNG order)
# --------------------------------------------------------------------------

def conjugate_symmetric_rows(half: np.ndarray) -> np.ndarray:
    """Stack [0; half; -flip(half)] -> (2K+1, n).  Ref: sspspace.py:821-827
    (`conjsym`; the `even` flag is ignored there, so the result is always odd)."""
    k, n = half.shape
    out = np.zeros((2 * k + 1, n))
    out[1:k + 1] = half
    out[k + 1:] = -half[::-1]
    return out


def random_phase_matrix(domain_dim, ssp_dim, scale_min=1e-3, scale_max=np.pi,
                        rng=None) -> np.ndarray:
    """Uniform-sampler phase matrix.  Ref: sspspace.py:618-622, :647.
    Call order on the Generator is `uniform` then `choice` - kept."""
    gen = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
    k = (ssp_dim - 1) // 2
    a = gen.uniform(0, 1, (k, domain_dim))
    sign = gen.choice((-1, +1), a.shape)
    half = sign * scale_max * (scale_min + a * (1 - 2 * scale_min))
    return conjugate_symmetric_rows(half)


def hex_phase_matrix(domain_dim, ssp_dim=151, n_scales=-1, n_rotates=-1,
                     scale_min=0.1, scale_max=np.pi) -> np.ndarray:
    """Hexagonal (simplex x scales x rotations) phase matrix for domain_dim<=2 with
    'lin' scale sampling.  Ref: sspspace.py:729-785 (domain_dim>2 draws random
    rotations from scipy and is covered by the golden fixtures only)."""
    n = domain_dim
    if n_rotates <= 0 or n_scales <= 0:
        n_rotates = int(np.sqrt((ssp_dim - 1) / (2 * (n + 1))))
        n_scales = n_rotates
    simplex = np.hstack([
        np.sqrt(1 + 1 / n) * np.identity(n) - (n ** (-3 / 2)) * (np.sqrt(n + 1) + 1),
        (n ** (-1 / 2)) * np.ones((n, 1))]).T                      # (n+1, n)
    if n == 1:
        n_scales = n_scales * n_rotates
    scales = np.linspace(scale_min, scale_max, n_scales)
    scaled = np.vstack([simplex * s for s in scales])
    if n_rotates == 1 or n == 1:
        rows = scaled
    elif n == 2:
        ang = np.linspace(0, 2 * np.pi / 3, n_rotates, endpoint=False)
        rot = np.stack([np.stack([np.cos(ang), -np.sin(ang)], axis=1),
                        np.stack([np.sin(ang), np.cos(ang)], axis=1)], axis=1)
        rows = (rot @ scaled.T).transpose(0, 2, 1).reshape(-1, n)
    else:
        raise NotImplementedError("oracle hex matrix: domain_dim<=2 only")
    return conjugate_symmetric_rows(rows)


# --------------------------------------------------------------------------
# encode / bind / decode
# --------------------------------------------------------------------------

def encode(phase_matrix: np.ndarray, length_scale, x) -> np.ndarray:
    """phi(x) = Re IFFT_axis0( exp(i A (x/l)^T) )^T -> (N, d) float64.
    Ref: sspspace.py:270-276."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    inv_ls = 1.0 / np.asarray(length_scale, dtype=np.float64).reshape(-1)
    if inv_ls.size == 1:
        inv_ls = np.repeat(inv_ls, phase_matrix.shape[1])
    scaled = x @ np.diag(inv_ls)
    spectrum = np.exp(1j * (phase_matrix @ scaled.T))
    return np.fft.ifft(spectrum, axis=0).real.T


def encode_and_deriv(phase_matrix: np.ndarray, length_scale, x):
    """(phi (N, d), grad (N, d, n)): grad[s, j, a] = d phi_j / d x_a = Re ifft_k( i (A_ka / l_a) exp(i A_k . x_s / l) )_j.
    Ref: sspspace.py:278-306 (`encode_and_deriv`).  The reference's own expression at :303-304,
    `(phase_matrix @ ls_mat) @ exp(...)`, multiplies (d, n) by (d, N) and raises for n != d; this restates the derivative its
    docstring describes (the derivative of `encode`, :275), checked against central differences of `encode` in the tests."""
    x = np.atleast_2d(x)
    ls = np.asarray(length_scale, dtype=np.float64).reshape(-1) * np.ones(phase_matrix.shape[1])
    z = np.exp(1.j * phase_matrix @ (x / ls).T)                                  # (d, N)
    phi = np.fft.ifft(z, axis=0).real.T
    w = phase_matrix / ls                                                        # (d, n)
    grad = np.stack([np.fft.ifft(1.j * w[:, a:a + 1] * z, axis=0).real.T for a in range(phase_matrix.shape[1])], axis=2)
    return phi, grad


def encode_closed_form(phase_matrix: np.ndarray, length_scale, x) -> np.ndarray:
    """Independent cross-check of `encode`: with A+ the K
    positive-frequency rows, phi_j = (1 + 2 sum_k cos(theta_k + 2 pi j k / d)) / d.
    Not a reference restatement; used to validate the folded psi-space algebra."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    d = phase_matrix.shape[0]
    k = (d - 1) // 2
    inv_ls = 1.0 / np.asarray(length_scale, dtype=np.float64).reshape(-1)
    if inv_ls.size == 1:
        inv_ls = np.repeat(inv_ls, phase_matrix.shape[1])
    theta = (x * inv_ls) @ phase_matrix[1:k + 1].T                 # (N, K)
    j = np.arange(d)[:, None] * np.arange(1, k + 1)[None, :]       # (d, K)
    w = 2 * np.pi * j / d
    return (1.0 + 2.0 * (np.cos(theta) @ np.cos(w).T - np.sin(theta) @ np.sin(w).T)) / d


def bind(*arrays) -> np.ndarray:
    """Circular convolution through the FFT.  Ref: sspspace.py:551-557."""
    arrays = [np.atleast_2d(a) for a in arrays]
    acc = np.fft.fft(arrays[0], axis=-1)
    for a in arrays[1:]:
        acc = acc * np.fft.fft(a, axis=-1)
    return np.fft.ifft(acc, axis=-1).real


def invert(a) -> np.ndarray:
    """Index reversal a[:, -arange(d)].  Ref: sspspace.py:558-560."""
    a = np.atleast_2d(a)
    return a[:, -np.arange(a.shape[1])]


def grid_sample_points(bounds: np.ndarray, per_dim) -> np.ndarray:
    """meshgrid('xy') of per-axis linspaces, flattened C-order -> (M, n).
    Ref: sspspace.py:488-495."""
    n = bounds.shape[0]
    if np.isscalar(per_dim):
        per_dim = [int(per_dim)] * n
    axes = [np.linspace(bounds[i, 0], bounds[i, 1], per_dim[i]) for i in range(n)]
    mesh = np.meshgrid(*axes)
    return np.array([m.reshape(-1) for m in mesh]).T


def decode_from_set(ssp, sample_ssps, sample_points) -> np.ndarray:
    """unit = ssp / max(||ssp||, 1e-6); argmax over the set of sample_ssps @ unit^T.
    Ref: sspspace.py:352-356."""
    ssp = np.atleast_2d(ssp)
    unit = ssp / np.maximum(np.linalg.norm(ssp, axis=-1, keepdims=True), 1e-6)
    sims = sample_ssps @ unit.T
    return sample_points[np.argmax(sims, axis=0), :]


# --------------------------------------------------------------------------
# Bayesian linear regression
# --------------------------------------------------------------------------

class BLR:
    """Posterior N(m, S) with rank-one updates.  Ref: blr.py:27-97."""

    def __init__(self, size_in, alpha=0.01, beta=None):
        self.d = size_in
        self.S_inv = alpha * np.eye(size_in)              # blr.py:30
        self.m = np.zeros((size_in, 1))                   # blr.py:31
        self.beta = beta
        self.S = np.linalg.pinv(self.S_inv)               # blr.py:33

    def update(self, phis, ts):
        """Ref: blr.py:54-77.  Note blr.py:67 evaluates `S @ phi.T @ phi @ S`
        left-to-right; the association below matches it so the float64 rounding
        is the same."""
        phis = np.asarray(phis, dtype=np.float64)
        ts = np.asarray(ts, dtype=np.float64)
        assert phis.shape[1] == self.d and ts.ndim == 2 and ts.shape[1] == 1
        if self.beta is None:                             # blr.py:54-59
            var_ts = np.var(ts) if len(ts) > 1 else np.abs(np.copy(ts[0]))
            self.beta = 1.0 / var_ts
        S_inv = self.S_inv + self.beta * (phis.T @ phis)  # blr.py:60
        S = self.S.copy()
        for i in range(phis.shape[0]):                    # blr.py:63-67
            phi = phis[i:i + 1] * np.sqrt(self.beta)
            scale = 1.0 + phi @ S @ phi.T
            S = S - (((S @ phi.T) @ phi) @ S) / scale
        x = self.beta * (phis.T @ ts)                     # blr.py:70
        self.m = S @ (self.S_inv @ self.m + x)            # blr.py:75
        self.S_inv = S_inv
        self.S = S

    def predict(self, phi):
        """mu (1,N), var (N,).  Ref: blr.py:96-97."""
        var = (1.0 / self.beta) + np.einsum('ij,ij->i', phi, phi @ self.S.T)
        return self.m.T @ phi.T, var


# --------------------------------------------------------------------------
# agent-level composition
# --------------------------------------------------------------------------

def decode_direct_optim(phase_matrix, length_scale, ssp, x0, bounds=None) -> np.ndarray:
    """x* = argmin_x -phi(x) . unit(ssp) by scipy L-BFGS-B from x0, per query.  Ref: sspspace.py:352, 358-375
    (`min_func`, `minimize(..., method='L-BFGS-B', bounds=self.domain_bounds)`, numerical gradients)."""
    from scipy.optimize import minimize
    ssp = np.atleast_2d(ssp)
    unit = ssp / np.maximum(np.linalg.norm(ssp, axis=-1, keepdims=True), 1e-6)
    out = np.zeros((ssp.shape[0], phase_matrix.shape[1]))
    for i, u in enumerate(unit):
        res = minimize(lambda x: -np.inner(encode(phase_matrix, length_scale, np.atleast_2d(x)), np.atleast_2d(u)).flatten(),
                       np.asarray(x0[i], dtype=np.float64).flatten(), method='L-BFGS-B', bounds=bounds)
        out[i] = res.x
    return out


def acquisition_func(m, S, beta, var_weight, gamma_t, margin=4):
    """(min_func, gradient) over phi-space.  Ref: ssp_agent.py:156-185 -- phi is pulled back to norm `margin` when it is
    longer; min_func = -(phi.m + var_weight sqrt(gamma_t + 1/beta + phi^T S phi) - sqrt(gamma_t));
    gradient = -(m + var_weight S phi / sqrt(phi^T S phi + gamma_t + 1/beta)) at the clamped point."""
    m = np.asarray(m, dtype=np.float64).reshape(-1, 1)
    beta_inv = 1.0 / float(np.ravel(beta)[0])

    def _clamp(phi):
        nrm = np.linalg.norm(phi)
        return margin * phi / nrm if nrm > margin else phi

    def min_func(phi):
        phi = _clamp(phi)
        val = phi.T @ m
        mi = var_weight * np.sqrt(gamma_t + beta_inv + phi.T @ S @ phi) - np.sqrt(gamma_t)
        return -(val + mi).flatten()

    def gradient(phi):
        phi = _clamp(phi)
        s_phi = S @ phi
        scale = np.sqrt(phi.T @ s_phi + gamma_t + beta_inv)
        return -(m.flatten() + var_weight * s_phi / scale)

    return min_func, gradient


def maximize_acquisition_lbfgs(min_func, gradient, phi_inits):
    """One scipy L-BFGS-B run per starting point.  Ref: bayesian_optimization.py:270-287 (`minimize(optim_func,
    phi_init.flatten(), jac=jac_func, method='L-BFGS-B')`, vals = -soln.fun).  Returns (phis (R, d), vals (R,), nit (R,))."""
    from scipy.optimize import minimize
    xs, vals, nit = [], [], []
    for p0 in np.atleast_2d(phi_inits):
        soln = minimize(min_func, np.asarray(p0, dtype=np.float64).flatten(), jac=gradient, method='L-BFGS-B')
        xs.append(soln.x); vals.append(-float(np.ravel(soln.fun)[0])); nit.append(soln.nit)
    return np.stack(xs), np.array(vals), np.array(nit)


def agent_eval(phis, blr: BLR, var_weight, gamma_t):
    """(mu, var, acq) with acq = var_weight (sqrt(var+gamma_t) - sqrt(gamma_t)).
    Ref: ssp_agent.py:133-137."""
    mu, var = blr.predict(phis)
    acq = var_weight * (np.sqrt(var + gamma_t) - np.sqrt(gamma_t))
    return mu, var, acq


def score_and_argmax(phis, blr: BLR, var_weight, gamma_t):
    """score = mu + acq ; i* = first index of the max (np.argmax semantics).
    Composition used by the reference's own demos: experiments/demo_blr_gif.py:129-136."""
    mu, var, acq = agent_eval(phis, blr, var_weight, gamma_t)
    score = mu.ravel() + np.ravel(acq)
    return score, int(np.argmax(score))


def agent_update_scalars(var_weight, gamma_t, var_decay, gamma_c, sigma_t):
    """lambda += var_decay ; gamma_t += gamma_c * sigma_t.  Ref: ssp_agent.py:208-218."""
    return var_weight + var_decay, gamma_t + gamma_c * sigma_t


def traj_timestep_ssps(t_phase_matrix, t_length_scale, traj_len):
    """Timestep SSPs at linspace(1, L+1, L).  Ref: ssp_traj_agent.py:81-83."""
    ts = np.linspace(1, traj_len + 1, traj_len).reshape(-1, 1)
    return encode(t_phase_matrix, t_length_scale, ts)


def traj_encode(x_phase_matrix, x_length_scale, timestep_ssps, x, traj_len, x_dim):
    """phi(traj) = sum_j bind(T_j, phi_x(x_j)).  Ref: ssp_traj_agent.py:155-163."""
    x = np.atleast_2d(x)
    pts = x.reshape(-1, traj_len, x_dim)
    acc = np.zeros((x.shape[0], x_phase_matrix.shape[0]))
    for j in range(traj_len):
        acc = acc + bind(timestep_ssps[j, :], encode(x_phase_matrix, x_length_scale, pts[:, j, :]))
    return acc


def traj_decode_from_set(ssp, timestep_ssps, sample_ssps, sample_points):
    """Unbind every timestep, then from-set decode each query.
    Ref: ssp_traj_agent.py:172-176."""
    queries = bind(invert(timestep_ssps), ssp)
    return decode_from_set(queries, sample_ssps, sample_points).reshape(-1)


def kfold_indices(n, n_splits):
    """Contiguous, unshuffled folds of sklearn.model_selection.KFold(n_splits): the first n % n_splits folds get one
    extra sample.  Yields (train_idx, test_idx).  Ref: ssp_traj_agent.py:121, 127."""
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[:n % n_splits] += 1
    idx = np.arange(n)
    start = 0
    for sz in sizes:
        test = idx[start:start + sz]
        yield np.concatenate([idx[:start], idx[start + sz:]]), test
        start += sz


def traj_lengthscale_cv_loss(x_phase_matrix, t_phase_matrix, length_scale, xs, ys, traj_len, x_dim):
    """Objective of the trajectory agent's length-scale search at length_scale = (ls_x, ls_t): K-fold (min(n, 50) folds)
    negative predictive log-likelihood of a fresh BLR per fold, timestep SSPs at linspace(0, L, L).
    Ref: ssp_traj_agent.py:117-137."""
    ls_x, ls_t = float(length_scale[0]), float(length_scale[1])
    T = encode(t_phase_matrix, np.array([ls_t]), np.linspace(0, traj_len, traj_len).reshape(-1, 1))
    phis = traj_encode(x_phase_matrix, ls_x * np.ones(x_dim), T, xs, traj_len, x_dim)
    errors = []
    for train, test in kfold_indices(xs.shape[0], min(xs.shape[0], 50)):
        b = BLR(phis.shape[1])
        b.update(phis[train], ys[train])
        mu, var = b.predict(phis[test])
        diff = ys[test].flatten() - mu.flatten()
        loss = -0.5 * np.log(var) - np.divide(np.power(diff, 2), var)
        errors.append(np.sum(-loss))
    return np.sum(errors)


def sp_vectors(domain_size, dim, rng) -> np.ndarray:
    """Unitary tag vectors of SPSpace.  Ref: sspspace.py:55-67, 153-159 (normal draws, column-normalised, then every
    Fourier coefficient divided by its magnitude)."""
    if domain_size == 1:
        v = np.zeros((1, dim))
        v[:, 0] = 1
        return v
    rng = np.random.default_rng(rng) if isinstance(rng, (int, np.integer)) else rng
    v = rng.normal(size=(domain_size, dim))
    v /= np.linalg.norm(v, axis=0, keepdims=True)
    fv = np.fft.fft(v, axis=-1)
    fv = fv / np.sqrt(fv.real ** 2 + fv.imag ** 2)
    return np.fft.ifft(fv, axis=-1).real


def _per_agent(obj, n_agents):
    """One entry per agent: a list / stacked array of per-agent items, or one shared item."""
    if isinstance(obj, (list, tuple)):
        assert len(obj) == n_agents
        return list(obj)
    return [obj] * n_agents


def multi_agent_encode(x_phase_matrix, x_length_scale, timestep_ssps, agent_sps, x, traj_len, n_agents, x_dim):
    """phi = unit(sum_i bind(a_i, sum_j bind(T_j, phi_x_i(x_ji)))), input layout (traj_len, n_agents, x_dim).
    x_phase_matrix / x_length_scale: one shared space, or lists with one entry per agent (same_agt_x_space=False).
    Ref: ssp_multi_agent.py:166-186."""
    x = np.atleast_2d(x)
    pts = x.reshape(-1, traj_len, n_agents, x_dim)
    As, lss = _per_agent(x_phase_matrix, n_agents), _per_agent(x_length_scale, n_agents)
    total = np.zeros((x.shape[0], As[0].shape[0]))
    for i in range(n_agents):
        si = np.zeros_like(total)
        for j in range(traj_len):
            si = si + bind(timestep_ssps[j, :], encode(As[i], lss[i], pts[:, j, i, :]))
        total = total + bind(agent_sps[i, :], si)
    return total / np.linalg.norm(total, axis=-1, keepdims=True)


def multi_agent_decode_from_set(ssp, timestep_ssps, agent_sps, sample_ssps, sample_points):
    """Unbind tag and timestep, from-set decode each query against the agent's own sample set (lists with one entry per
    agent, or one shared set).  Ref: ssp_multi_agent.py:195-208."""
    ssp = np.atleast_2d(ssp)
    ssp = ssp / np.linalg.norm(ssp, axis=-1, keepdims=True)
    L, n_agents = timestep_ssps.shape[0], agent_sps.shape[0]
    sss, sps = _per_agent(sample_ssps, n_agents), _per_agent(sample_points, n_agents)
    out = np.zeros((L, n_agents, sps[0].shape[1]))
    for i in range(n_agents):
        sspi = bind(invert(agent_sps[i:i + 1]), ssp)
        queries = bind(invert(timestep_ssps), sspi)
        out[:, i, :] = decode_from_set(queries, sss[i], sps[i])
    return out.reshape(-1)


def multi_agent_lengthscale_cv_loss(x_phase_matrix, t_phase_matrix, agent_sps, length_scale, xs, ys, traj_len, n_agents, x_dim):
    """Objective of the multi-agent length-scale search at length_scale = (ls_x, ls_t) -- the reference reads entries 0 and 1
    only and gives every agent's x space ls_x: K-fold (min(n, 50) folds) negative predictive log-likelihood of a fresh BLR
    per fold on the L2-normalised encodings, timestep SSPs at linspace(0, L, L).  Ref: ssp_multi_agent.py:131-157."""
    ls_x, ls_t = float(length_scale[0]), float(length_scale[1])
    T = encode(t_phase_matrix, np.array([ls_t]), np.linspace(0, traj_len, traj_len).reshape(-1, 1))
    As = _per_agent(x_phase_matrix, n_agents)
    phis = multi_agent_encode(As, [ls_x * np.ones(x_dim)] * n_agents, T, agent_sps, xs, traj_len, n_agents, x_dim)
    errors = []
    for train, test in kfold_indices(xs.shape[0], min(xs.shape[0], 50)):
        b = BLR(phis.shape[1])
        b.update(phis[train], ys[train])
        mu, var = b.predict(phis[test])
        diff = ys[test].flatten() - mu.flatten()
        loss = -0.5 * np.log(var) - np.divide(np.power(diff, 2), var)
        errors.append(np.sum(-loss))
    return np.sum(errors)


# --------------------------------------------------------------------------
# synthetic candidate generator shared by oracle, tests and the CUDA path
# --------------------------------------------------------------------------

def rff_encode(W, B, scale, x) -> np.ndarray:
    """Random-Fourier-feature map: scale * cos(W x^T + B)^T, (N, m).  Ref: agents/rff_agent.py:153-156 (`encode`);
    scale = sqrt(2 alpha / m) (:64-65), W (m, n) and B (m, 1) drawn at :66-73."""
    x = np.atleast_2d(x)
    return scale * np.cos(np.dot(W, x.T) + np.reshape(B, (-1, 1))).T


def counter_uniform(seed: int, start: int, count: int, n: int) -> np.ndarray:
    """Counter-based uniform(0,1) float32 candidates: element (i, a) of the
    stream is a function of (seed, i*n + a) only, so any index range can be
    regenerated anywhere (device kernel `sspbo_fill_uniform` implements the
    same integer hash; bit-exact).  Not part of the reference: BASELINE config
    C3 asks for synthetic candidates and SURVEY.md section 8d specifies this
    generator."""
    idx = (np.arange(start * n, (start + count) * n, dtype=np.uint64)
           + (np.uint64(seed) << np.uint64(40)))
    z = idx + np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over='ignore'):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    return u.reshape(count, n)


def hartmann6(x):
    """Negated Hartmann-6 (Surjanovic & Bingham constants).  Not in the reference
    (SURVEY.md section 2 row 19); used as the objective of BASELINE config C3."""
    a = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14],
                  [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    x = np.atleast_2d(x)
    inner = np.einsum('ij,nij->ni', A, (x[:, None, :] - P[None]) ** 2)
    return (a[None] * np.exp(-inner)).sum(axis=1)
