"""Multi-agent trajectory agent on the GPU.  Mirrors `ssp_bayes_opt/agents/ssp_multi_agent.py`.

phi = unit( sum_i a_i (*) sum_j T_j (*) phi_x(x_ji) )   (ssp_multi_agent.py:166-186).  The tags a_i and the timestep SSPs
T_j are unitary, so the spectrum of the un-normalised sum is a sum of unit phasors over the traj_len * n_agents
waypoints, exp(i(A x_ji / l + phase(T_j) + phase(a_i))) per frequency: the device encoder / scorer of the single-agent
trajectory path does the whole thing with one phase-offset row per (timestep, agent) (`sspbo_space_set_waypoints`), and
divides by |phi| at the end; no FFT and no explicit bind is executed.  Decoding unbinds tag and timestep and runs the GPU
from-set / direct-optim search per waypoint (ssp_multi_agent.py:188-208).

Agents may share one x space (`same_agt_x_space=True`, the reference default) or have one each (`same_agt_x_space=False`
or a list of distinct `ssp_x_spaces`, :58-66): the device context then holds one frequency table per agent and waypoint
(j, i) reads table i (`sspbo_space_set_classes`).  Without `length_scale` the K-fold search of :131-164 runs with the
encodings and the Gram matrix computed on the device.  Out of scope (raise): learned decoders.
"""
from __future__ import annotations

import numpy as np

from .. import sspspace
from ..engine import DeviceEngine
from .ssp_agent import SSPAgent


def waypoint_phase_table(timesteps, t_phase_matrix, t_length_scale, agent_sps):
    """Phase offsets of the waypoints (timestep j, agent i), in the candidate layout (traj_len, n_agents, x_dim):
    tau[(j, i), k] = A_t[k] t_j / l_t + angle(FFT(a_i))[k] for the K positive frequencies (radians), and the +-1 DC
    coefficient of each waypoint's tag (a unitary real vector has a DC coefficient of +1 or -1).  With them
    sum_i a_i (*) sum_j T_j (*) phi_x(x_ji) has the spectrum sum_(j,i) exp(i (A x_ji / l + tau[(j, i)])) per frequency and
    sum_(j,i) dc[(j, i)] at DC."""
    timesteps = np.asarray(timesteps, dtype=np.float64).reshape(-1, 1)
    agent_sps = np.atleast_2d(agent_sps)
    K = (agent_sps.shape[1] - 1) // 2
    a_t = np.asarray(t_phase_matrix)[1:K + 1, 0]
    ls_t = float(np.ravel(t_length_scale)[0])
    tau_t = timesteps * (a_t / ls_t)[None, :]                                   # (L, K)
    spec = np.fft.fft(agent_sps, axis=-1)
    tag_phase = np.angle(spec[:, 1:K + 1])                                      # (n_agents, K)
    tag_dc = np.where(spec[:, 0].real < 0, -1, 1).astype(np.int32)
    tau = (tau_t[:, None, :] + tag_phase[None, :, :]).reshape(timesteps.shape[0] * agent_sps.shape[0], K)
    return tau, np.tile(tag_dc, timesteps.shape[0])


class SSPMultiAgent(SSPAgent):
    def __init__(self, init_xs, init_ys, **kwargs):
        super().__init__(init_xs, init_ys, None, **kwargs)

    def _set_ssp_space(self, n_agents, x_dim, traj_len, domain_bounds, ssp_x_spaces=None, ssp_t_space=None, seed=0,
                       init_pos=None, same_agt_x_space=True, **kwargs):
        assert n_agents > 0
        self.data_dim = x_dim * traj_len * n_agents
        self.n_agents = n_agents
        self.x_dim = x_dim
        self.init_pos = None if init_pos is None else np.atleast_2d(init_pos)
        self.traj_len = traj_len
        self.same_agt_x_space = same_agt_x_space
        self.domain_bounds = domain_bounds
        if ssp_x_spaces is None:
            if same_agt_x_space:
                ssp_x_space = sspspace.HexagonalSSPSpace(x_dim, ssp_dim=kwargs.get('ssp_dim', 100),
                                                         domain_bounds=domain_bounds[:x_dim, :], length_scale=1, rng=seed)
                ssp_x_spaces = [ssp_x_space] * n_agents
            else:                                        # one random space per agent on its own rows of the bounds (:58-66)
                ssp_x_spaces = [sspspace.RandomSSPSpace(x_dim, ssp_dim=kwargs.get('ssp_dim', 100),
                                                        domain_bounds=domain_bounds[i * x_dim:(i + 1) * x_dim, :],
                                                        length_scale=1, rng=seed + i) for i in range(n_agents)]
        ssp_x_spaces = list(ssp_x_spaces)
        assert len(ssp_x_spaces) == n_agents
        if any(sp.ssp_dim != ssp_x_spaces[0].ssp_dim or sp.domain_dim != x_dim for sp in ssp_x_spaces):
            raise ValueError("the agents' x spaces must share ssp_dim and x_dim")
        self._shared_x_space = all(sp is ssp_x_spaces[0] for sp in ssp_x_spaces)
        if ssp_t_space is None:
            ssp_t_space = sspspace.RandomSSPSpace(1, ssp_dim=ssp_x_spaces[0].ssp_dim,
                                                  domain_bounds=np.array([[0, self.traj_len + 1]]), length_scale=1,
                                                  rng=seed + n_agents)
        assert ssp_t_space.ssp_dim == ssp_x_spaces[0].ssp_dim
        self.ssp_dim = ssp_x_spaces[0].ssp_dim
        self.ssp_x_spaces = ssp_x_spaces
        self.ssp_t_space = ssp_t_space
        self.ssp_space = ssp_x_spaces[0]

        agent_space = sspspace.SPSpace(n_agents, self.ssp_dim, rng=seed)
        self.agent_sps = agent_space.vectors
        self.agent_inv_sps = agent_space.inverse_vectors

        self._traj_engine = DeviceEngine(self.ssp_space._device)
        if 'length_scale' not in kwargs or np.any(np.array(kwargs.get('length_scale')) < 0):
            optres = self._optimize_lengthscale(self.init_xs, self.init_ys)
            for sp in self._distinct_x_spaces():
                sp.update_lengthscale(optres[0])
            self.ssp_t_space.update_lengthscale(optres[-1])
        else:
            for sp in self._distinct_x_spaces():
                sp.update_lengthscale(kwargs.get('length_scale', 4))
            self.ssp_t_space.update_lengthscale(kwargs.get('time_length_scale', 1))

        timesteps = np.linspace(1, self.traj_len + 1, self.traj_len).reshape(-1, 1)
        self.timestep_ssps = self.ssp_t_space.encode(timesteps)
        self.timestep_inv_ssps = self.ssp_t_space.encode(-timesteps)
        self._point_engine(timesteps)

    def _distinct_x_spaces(self):
        return self.ssp_x_spaces[:1] if self._shared_x_space else self.ssp_x_spaces

    def _point_engine(self, timesteps):
        """(Re)load the device context that sees whole multi-agent trajectories with the spaces' current length scales: one
        phase-offset row per (timestep, agent) waypoint in the candidate layout (traj_len, n_agents, x_dim), and one
        frequency table per agent when the agents have their own x spaces."""
        tau, dc = waypoint_phase_table(timesteps, self.ssp_t_space.phase_matrix, self.ssp_t_space.length_scale, self.agent_sps)
        eng = self._traj_engine
        if self._shared_x_space:
            eng.set_space(self.ssp_space.phase_matrix, self.ssp_space.length_scale)
        else:
            eng.set_space_classes([sp.phase_matrix for sp in self.ssp_x_spaces], [sp.length_scale for sp in self.ssp_x_spaces])
        if self.domain_bounds is not None:               # declared |x| bound: lets the scorer use its fast phase arithmetic
            eng.set_xbound(float(np.max(np.abs(np.asarray(self.domain_bounds, dtype=np.float64)))))
        eng.set_waypoints(tau, dc_sign=dc, normalize=True)

    def _cv_loss(self, length_scale, xs, ys):
        """Objective of the length-scale search (ssp_multi_agent.py:135-157): entries 0 and 1 of the trial vector are the
        length scale of every agent's x space and of the time space; negative predictive log-likelihood over min(n, 50)
        contiguous folds of fresh BLRs on the normalised encodings, timestep SSPs at linspace(0, L, L).  Encode + Gram
        matrix on the device, folds in dual form."""
        from .ssp_traj_agent import cv_loss_from_gram
        ls = np.ravel(length_scale)
        for sp in self._distinct_x_spaces():
            sp.update_lengthscale(ls[0])
        self.ssp_t_space.update_lengthscale(ls[1])
        self._point_engine(np.linspace(0, self.traj_len, self.traj_len))
        eng = self._traj_engine
        G = eng.gram(eng.encode_device(xs)).cpu().numpy()
        return cv_loss_from_gram(G, ys)

    def _optimize_lengthscale(self, init_trajs, init_ys):
        """ssp_multi_agent.py:131-164.  The reference starts from 4 * ones(n_agents + 1) against two bounds, which scipy only
        accepts for one agent; the objective reads entries 0 and 1 alone, so the search here runs over those two -- (x, time)
        from (4, 4) -- for any number of agents.  Returns them."""
        from scipy.optimize import minimize
        xs = np.ascontiguousarray(np.atleast_2d(init_trajs), dtype=np.float64)
        low = 1 / np.sqrt(xs.shape[0])
        retval = minimize(lambda ls: self._cv_loss(ls, xs, init_ys), x0=np.array([4., 4.]), method='L-BFGS-B',
                          bounds=[(low, None), (low, None)])
        return np.abs(retval.x)

    def _scoring_engine(self):
        return self._traj_engine

    def _set_decoder(self):
        if self.decoder_method in ('regression', 'network', 'network-optim'):
            raise NotImplementedError(f"decoder_method={self.decoder_method!r} (learned decoders) is out of scope; "
                                      "use 'from-set' or 'direct-optim'")
        if self._shared_x_space:
            self._init_samples_of = lambda: [self.get_init_samples(self.ssp_x_spaces[0])] * self.n_agents
        else:
            self._init_samples_of = lambda: [self.get_init_samples(sp) for sp in self.ssp_x_spaces]

    def length_scale(self):
        return np.array([np.ravel(sp.length_scale)[0] for sp in self.ssp_x_spaces] + [np.ravel(self.ssp_t_space.length_scale)[0]])

    def encode(self, x, timestep_ssps=None):
        """(n, traj_len * n_agents * x_dim) -> (n, ssp_dim), L2-normalised."""
        if timestep_ssps is not None:
            raise NotImplementedError("overriding timestep_ssps belongs to the length-scale search (out of scope)")
        x = np.atleast_2d(x)
        return self._traj_engine.encode(x.reshape(x.shape[0], self.data_dim))

    def _random_inputs(self, count, rng):
        b = np.asarray(self.domain_bounds, dtype=np.float64)
        if not self._shared_x_space and b.shape[0] >= self.n_agents * self.x_dim:     # each agent on its own rows of the bounds
            b = np.tile(b[:self.n_agents * self.x_dim], (self.traj_len, 1))
        else:
            b = np.tile(b[:self.x_dim], (self.traj_len * self.n_agents, 1))
        return rng.uniform(b[:, 0], b[:, 1], size=(count, b.shape[0]))

    def decode(self, ssp, timestep_ssps=None):
        """SSP -> flattened (traj_len, n_agents, x_dim) trajectory (ssp_multi_agent.py:188-208)."""
        inv_t = self.timestep_inv_ssps if timestep_ssps is None else self.ssp_t_space.invert(timestep_ssps)
        ssp = np.atleast_2d(ssp)
        ssp = ssp / np.linalg.norm(ssp, axis=-1, keepdims=True)
        out = np.zeros((self.traj_len, self.n_agents, self.x_dim))
        for i in range(self.n_agents):
            sp = self.ssp_x_spaces[i]
            sspi = sp.bind(self.agent_inv_sps[i, :], ssp)
            queries = sp.bind(inv_t, sspi)
            out[:, i, :] = sp.decode(queries, method=self.decoder_method, samples=self.init_samples[i])
        return out.reshape(-1)
