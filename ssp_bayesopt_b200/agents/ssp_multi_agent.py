"""Multi-agent trajectory agent on the GPU.  Mirrors `ssp_bayes_opt/agents/ssp_multi_agent.py`.

phi = unit( sum_i a_i (*) sum_j T_j (*) phi_x(x_ji) )   (ssp_multi_agent.py:166-186).  The tags a_i and the timestep SSPs
T_j are unitary, so the spectrum of the un-normalised sum is a sum of unit phasors over the traj_len * n_agents
waypoints, exp(i(A x_ji / l + phase(T_j) + phase(a_i))) per frequency: the device encoder / scorer of the single-agent
trajectory path does the whole thing with one phase-offset row per (timestep, agent) (`sspbo_space_set_waypoints`), and
divides by |phi| at the end; no FFT and no explicit bind is executed.  Decoding unbinds tag and timestep and runs the GPU
from-set / direct-optim search per waypoint (ssp_multi_agent.py:188-208).

Scope: agents share one x space (`same_agt_x_space=True`, the reference default).  Out of scope (raise): per-agent phase
matrices, the K-fold length-scale search (:131-164), learned decoders.
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
            if not same_agt_x_space:
                raise NotImplementedError("per-agent x spaces (same_agt_x_space=False) are outside the device encoder: "
                                          "all agents must share one phase matrix")
            ssp_x_space = sspspace.HexagonalSSPSpace(x_dim, ssp_dim=kwargs.get('ssp_dim', 100),
                                                     domain_bounds=domain_bounds[:x_dim, :], length_scale=1, rng=seed)
            ssp_x_spaces = [ssp_x_space] * n_agents
        elif any(sp is not ssp_x_spaces[0] for sp in ssp_x_spaces):
            raise NotImplementedError("per-agent x spaces are outside the device encoder: pass one shared SSPSpace")
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

        if 'length_scale' not in kwargs or np.any(np.array(kwargs.get('length_scale')) < 0):
            raise NotImplementedError(
                "length_scale / time_length_scale must be given: the K-fold search of the reference "
                "(ssp_multi_agent.py:131-164) is outside the GPU hot-path scope")
        self.ssp_x_spaces[0].update_lengthscale(kwargs.get('length_scale', 4))
        self.ssp_t_space.update_lengthscale(kwargs.get('time_length_scale', 1))

        timesteps = np.linspace(1, self.traj_len + 1, self.traj_len).reshape(-1, 1)
        self.timestep_ssps = self.ssp_t_space.encode(timesteps)
        self.timestep_inv_ssps = self.ssp_t_space.encode(-timesteps)

        # device context that sees whole multi-agent trajectories: one phase-offset row per (timestep, agent) waypoint,
        # in the candidate layout (traj_len, n_agents, x_dim)
        tau, dc = waypoint_phase_table(timesteps, self.ssp_t_space.phase_matrix, self.ssp_t_space.length_scale, self.agent_sps)
        self._traj_engine = DeviceEngine(self.ssp_space._device)
        self._traj_engine.set_space(self.ssp_space.phase_matrix, self.ssp_space.length_scale)
        if domain_bounds is not None:                    # declared |x| bound: lets the scorer use its fast phase arithmetic
            self._traj_engine.set_xbound(float(np.max(np.abs(np.asarray(domain_bounds, dtype=np.float64)))))
        self._traj_engine.set_waypoints(tau, dc_sign=dc, normalize=True)

    def _scoring_engine(self):
        return self._traj_engine

    def _set_decoder(self):
        if self.decoder_method in ('regression', 'network', 'network-optim'):
            raise NotImplementedError(f"decoder_method={self.decoder_method!r} (learned decoders) is out of scope; "
                                      "use 'from-set' or 'direct-optim'")
        self._init_samples_of = lambda: [self.get_init_samples(self.ssp_x_spaces[0])] * self.n_agents

    def length_scale(self):
        return np.array([sp.length_scale for sp in self.ssp_x_spaces] + [self.ssp_t_space.length_scale])

    def encode(self, x, timestep_ssps=None):
        """(n, traj_len * n_agents * x_dim) -> (n, ssp_dim), L2-normalised."""
        if timestep_ssps is not None:
            raise NotImplementedError("overriding timestep_ssps belongs to the length-scale search (out of scope)")
        x = np.atleast_2d(x)
        return self._traj_engine.encode(x.reshape(x.shape[0], self.data_dim))

    def _random_inputs(self, count, rng):
        b = np.asarray(self.domain_bounds, dtype=np.float64)[:self.x_dim]
        b = np.tile(b, (self.traj_len * self.n_agents, 1))
        return rng.uniform(b[:, 0], b[:, 1], size=(count, b.shape[0]))

    def decode(self, ssp, timestep_ssps=None):
        """SSP -> flattened (traj_len, n_agents, x_dim) trajectory (ssp_multi_agent.py:188-208)."""
        inv_t = self.timestep_inv_ssps if timestep_ssps is None else self.ssp_t_space.invert(timestep_ssps)
        ssp = np.atleast_2d(ssp)
        ssp = ssp / np.linalg.norm(ssp, axis=-1, keepdims=True)
        out = np.zeros((self.traj_len, self.n_agents, self.x_dim))
        sp = self.ssp_x_spaces[0]
        for i in range(self.n_agents):
            sspi = sp.bind(self.agent_inv_sps[i, :], ssp)
            queries = sp.bind(inv_t, sspi)
            out[:, i, :] = sp.decode(queries, method=self.decoder_method, samples=self.init_samples[i])
        return out.reshape(-1)
