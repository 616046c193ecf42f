"""Trajectory agent on the GPU.  Mirrors `ssp_bayes_opt/agents/ssp_traj_agent.py`.

phi(traj) = sum_j T_j (*) phi_x(x_j) (ssp_traj_agent.py:145-163).  In the Fourier domain the binding is a
product of unit phasors, so the spectrum of phi(traj) is sum_j exp(i(A x_j / l_x + A_t t_j / l_t)) and
the device encoder / scorer just sums L phasors per frequency (`sspbo_space_set_traj`); no FFT and no
explicit bind is executed.  Decoding unbinds each timestep and runs the GPU from-set search
(ssp_traj_agent.py:165-176).

Omitting `length_scale` runs the reference's K-fold length-scale search (:113-143): scipy's L-BFGS-B
on the host drives it as in the reference; every objective evaluation encodes the initial design on the device at the
trial length scales, forms its Gram matrix there (`sspbo_gram`), and solves each fold's BLR in dual form (n x n).

Out of scope (raise): the sklearn MLP 'regression' decoder (:85-102).
"""
from __future__ import annotations

import numpy as np

from .. import sspspace
from ..engine import DeviceEngine
from .ssp_agent import SSPAgent


def cv_loss_from_gram(G, ys, alpha=0.01, max_splits=50):
    """K-fold negative predictive log-likelihood of ssp_traj_agent.py:121-137 from the Gram matrix G = Phi Phi^T of the
    encoded design: min(n, 50) contiguous folds (KFold without shuffling), per fold a fresh BLR with prior precision
    alpha and beta = 1 / var(train targets) (blr.py:54-59), written in its dual form
        mu = k^T (G_tr + (alpha/beta) I)^-1 t ,   var = 1/beta + (k** - k^T (G_tr + (alpha/beta) I)^-1 k) / alpha."""
    ys = np.asarray(ys, dtype=np.float64).reshape(-1)
    n = ys.shape[0]
    n_splits = min(n, max_splits)
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[:n % n_splits] += 1
    total, start, idx = 0.0, 0, np.arange(n)
    for sz in sizes:
        te = idx[start:start + sz]
        tr = np.concatenate([idx[:start], idx[start + sz:]])
        start += sz
        t = ys[tr]
        beta = 1.0 / np.var(t) if len(t) > 1 else 1.0 / np.abs(t[0])
        A = G[np.ix_(tr, tr)] + (alpha / beta) * np.eye(len(tr))
        k = G[np.ix_(te, tr)]
        sol = np.linalg.solve(A, np.column_stack([t, k.T]))
        mu = k @ sol[:, 0]
        var = 1.0 / beta + (G[te, te] - np.sum(k * sol[:, 1:].T, axis=1)) / alpha
        diff = ys[te] - mu
        total += np.sum(0.5 * np.log(var) + diff ** 2 / var)
    return total


class SSPTrajectoryAgent(SSPAgent):
    def __init__(self, init_xs, init_ys, **kwargs):
        super().__init__(init_xs, init_ys, None, **kwargs)
        if self.init_pos is not None:
            constraint_val = self.ssp_x_space.bind(self.timestep_ssps[0, :], self.ssp_x_space.encode(self.init_pos))
            self.constraint_ssp += constraint_val.T

    def _set_ssp_space(self, x_dim, traj_len, domain_bounds, ssp_x_space=None, ssp_t_space=None, init_pos=None,
                       seed=0, **kwargs):
        self.data_dim = x_dim * traj_len
        self.x_dim = x_dim
        self.init_pos = None if init_pos is None else np.atleast_2d(init_pos)
        self.traj_len = traj_len
        if domain_bounds is not None:
            domain_bounds = np.array([np.min(domain_bounds[:, 0]) * np.ones(x_dim),
                                      np.max(domain_bounds[:, 1]) * np.ones(x_dim)]).T
        self.domain_bounds = domain_bounds
        if ssp_x_space is None:
            ssp_x_space = sspspace.HexagonalSSPSpace(x_dim, ssp_dim=kwargs.get('ssp_dim', 100),
                                                     domain_bounds=domain_bounds, length_scale=1, rng=seed)
        if ssp_t_space is None:
            ssp_t_space = sspspace.RandomSSPSpace(1, ssp_dim=ssp_x_space.ssp_dim,
                                                  domain_bounds=np.array([[0, self.traj_len]]), length_scale=1, rng=seed)
        assert ssp_t_space.ssp_dim == ssp_x_space.ssp_dim
        self.ssp_dim = ssp_x_space.ssp_dim
        self.ssp_x_space = ssp_x_space
        self.ssp_t_space = ssp_t_space
        self.ssp_space = ssp_x_space
        # device context that sees whole trajectories: spatial phases + per-timestep phase offsets
        self._traj_engine = DeviceEngine(self.ssp_x_space._device)
        if 'length_scale' not in kwargs or np.any(np.array(kwargs.get('length_scale')) < 0):
            optres = self._optimize_lengthscale(self.init_xs, self.init_ys)
            self.ssp_x_space.update_lengthscale(optres[0])
            self.ssp_t_space.update_lengthscale(optres[1])
        else:
            self.ssp_x_space.update_lengthscale(kwargs.get('length_scale'))
            self.ssp_t_space.update_lengthscale(kwargs.get('time_length_scale', 1))

        timesteps = np.linspace(1, self.traj_len + 1, self.traj_len).reshape(-1, 1)
        self.timestep_ssps = self.ssp_t_space.encode(timesteps)
        self.timestep_inv_ssps = self.ssp_t_space.encode(-timesteps)
        self._point_engine(timesteps)

    def _point_engine(self, timesteps):
        """(Re)load the trajectory context with the spaces' current length scales and the given timestep values."""
        K = (self.ssp_dim - 1) // 2
        a_t = self.ssp_t_space.phase_matrix[1:K + 1, 0]
        ls_t = float(np.ravel(self.ssp_t_space.length_scale)[0])
        tau = np.asarray(timesteps, dtype=np.float64).reshape(-1, 1) * (a_t / ls_t)[None, :]     # (L, K) radians
        self._traj_engine.set_space(self.ssp_x_space.phase_matrix, self.ssp_x_space.length_scale)
        if self.domain_bounds is not None:               # declared |x| bound: lets the scorer use its fast phase arithmetic
            self._traj_engine.set_xbound(float(np.max(np.abs(np.asarray(self.domain_bounds, dtype=np.float64)))))
        self._traj_engine.set_traj(tau, self.traj_len)

    def _cv_loss(self, length_scale, xs, ys):
        """Objective of the length-scale search at (ls_x, ls_t) (ssp_traj_agent.py:117-137): negative predictive
        log-likelihood summed over min(n, 50) contiguous folds, each with a fresh BLR (alpha = 0.01, beta = 1 / var(train
        targets)), timestep SSPs at linspace(0, L, L).  Encode + Gram matrix on the device, folds in dual form."""
        ls = np.ravel(length_scale)
        self.ssp_x_space.update_lengthscale(ls[0])
        self.ssp_t_space.update_lengthscale(ls[1])
        self._point_engine(np.linspace(0, self.traj_len, self.traj_len))
        eng = self._traj_engine
        G = eng.gram(eng.encode_device(xs)).cpu().numpy()
        return cv_loss_from_gram(G, ys)

    def _optimize_lengthscale(self, init_trajs, init_ys):
        from scipy.optimize import minimize
        xs = np.ascontiguousarray(np.atleast_2d(init_trajs), dtype=np.float64)
        low = 1 / np.sqrt(xs.shape[0])
        retval = minimize(lambda ls: self._cv_loss(ls, xs, init_ys), x0=np.array([4., 10.]), method='L-BFGS-B',
                          bounds=[(low, None), (low, None)])
        return np.abs(retval.x)

    def _scoring_engine(self):
        return self._traj_engine

    def _set_decoder(self):
        if self.decoder_method in ('regression', 'network', 'network-optim'):
            raise NotImplementedError(f"decoder_method={self.decoder_method!r} (learned decoders) is out of scope; "
                                      "use 'from-set' or 'direct-optim'")
        self._init_samples_of = lambda: self.get_init_samples(self.ssp_x_space)

    def length_scale(self):
        return np.array([self.ssp_x_space.length_scale, self.ssp_t_space.length_scale])

    def encode(self, x):
        """(n, traj_len * x_dim) -> (n, ssp_dim)."""
        x = np.atleast_2d(x)
        return self._traj_engine.encode(x.reshape(x.shape[0], self.traj_len * self.x_dim))

    def _random_inputs(self, count, rng):
        b = np.tile(np.asarray(self.domain_bounds, dtype=np.float64), (self.traj_len, 1))
        return rng.uniform(b[:, 0], b[:, 1], size=(count, b.shape[0]))

    def decode(self, ssp):
        queries = self.ssp_x_space.bind(self.ssp_t_space.invert(self.timestep_ssps), ssp)
        decoded = self.ssp_x_space.decode(queries, method=self.decoder_method, samples=self.init_samples)
        return decoded.reshape(-1)
