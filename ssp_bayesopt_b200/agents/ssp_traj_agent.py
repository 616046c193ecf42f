"""Trajectory agent on the GPU.  Mirrors `ssp_bayes_opt/agents/ssp_traj_agent.py`.

phi(traj) = sum_j T_j (*) phi_x(x_j) (ssp_traj_agent.py:145-163).  In the Fourier domain the binding is a
product of unit phasors, so the spectrum of phi(traj) is sum_j exp(i(A x_j / l_x + A_t t_j / l_t)) and
the device encoder / scorer just sums L phasors per frequency (`sspbo_space_set_traj`); no FFT and no
explicit bind is executed.  Decoding unbinds each timestep and runs the GPU from-set search
(ssp_traj_agent.py:165-176).

Out of scope (raise): K-fold length-scale search (:113-143) and the sklearn MLP 'regression' decoder (:85-102).
"""
from __future__ import annotations

import numpy as np

from .. import sspspace
from ..engine import DeviceEngine
from .ssp_agent import SSPAgent


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
        if 'length_scale' not in kwargs or np.any(np.array(kwargs.get('length_scale')) < 0):
            raise NotImplementedError(
                "length_scale / time_length_scale must be given: the K-fold search of the reference "
                "(ssp_traj_agent.py:113-143) is outside the GPU hot-path scope")
        self.ssp_x_space.update_lengthscale(kwargs.get('length_scale'))
        self.ssp_t_space.update_lengthscale(kwargs.get('time_length_scale', 1))

        timesteps = np.linspace(1, self.traj_len + 1, self.traj_len).reshape(-1, 1)
        self.timestep_ssps = self.ssp_t_space.encode(timesteps)
        self.timestep_inv_ssps = self.ssp_t_space.encode(-timesteps)

        # device context that sees whole trajectories: spatial phases + per-timestep phase offsets
        K = (self.ssp_dim - 1) // 2
        a_t = self.ssp_t_space.phase_matrix[1:K + 1, 0]
        ls_t = float(np.ravel(self.ssp_t_space.length_scale)[0])
        tau = timesteps * (a_t / ls_t)[None, :]                       # (L, K) radians
        self._traj_engine = DeviceEngine(self.ssp_x_space._device)
        self._traj_engine.set_space(self.ssp_x_space.phase_matrix, self.ssp_x_space.length_scale)
        self._traj_engine.set_traj(tau, self.traj_len)

    def _scoring_engine(self):
        return self._traj_engine

    def _set_decoder(self):
        if self.decoder_method in ('regression', 'network', 'network-optim'):
            raise NotImplementedError(f"decoder_method={self.decoder_method!r} (learned decoders) is out of scope; "
                                      "use 'from-set' or 'direct-optim'")
        self.init_samples = self.get_init_samples(self.ssp_x_space)

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
