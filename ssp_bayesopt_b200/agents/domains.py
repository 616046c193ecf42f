"""Initial-design samplers (ssp_bayes_opt/agents/domains.py).  Host-side, unseeded Sobol exactly like the
reference; pass `init_points=(xs, ys)` to `maximize` for reproducible runs."""
import numpy as np
from scipy.stats import qmc


class Domain:
    def sample(self, num_points: int = 10, method: str = 'sobol') -> np.ndarray:
        raise NotImplementedError


class BoundedDomain(Domain):
    """domains.py:24-38"""

    def __init__(self, bounds: np.ndarray):
        self.bounds = bounds
        self.data_dim = bounds.shape[0]
        self.sampler = qmc.Sobol(d=self.data_dim)

    def sample(self, num_points: int = 10, method: str = 'sobol') -> np.ndarray:
        if method != 'sobol':
            raise NotImplementedError(f'{method} not implemented')
        return qmc.scale(self.sampler.random(num_points), self.bounds[:, 0], self.bounds[:, 1])


class TrajectoryDomain(Domain):
    """Flattened (traj_len * spatial_dim) trajectories inside a square box (domains.py:84-106)."""

    def __init__(self, trajectory_length: int, spatial_dim: int, bounds: np.ndarray):
        self.traj_len = trajectory_length
        self.spatial_dim = spatial_dim
        self.domain_bounds = np.array([bounds[:, 0].min() * np.ones(spatial_dim),
                                       bounds[:, 1].max() * np.ones(spatial_dim)]).T
        self.sampler = qmc.Sobol(d=spatial_dim)

    def sample(self, num_points: int = 10, method: str = 'sobol') -> np.ndarray:
        if method != 'sobol':
            raise NotImplementedError(f'{method} not implemented')
        u = self.sampler.random(num_points * self.traj_len)
        pts = qmc.scale(u, self.domain_bounds[:, 0], self.domain_bounds[:, 1])
        return pts.reshape(num_points, self.traj_len * self.spatial_dim)


class MultiTrajectoryDomain(Domain):
    """Flattened (traj_len * spatial_dim * n_agents) multi-agent trajectories (domains.py:41-81), including the
    reference's (num_points, traj_len, spatial_dim, n_agents) view when the last waypoints are pinned near goals."""

    def __init__(self, n_agents: int, trajectory_length: int, spatial_dim: int, bounds: np.ndarray, goals=None):
        self.n_agents = n_agents
        self.traj_len = trajectory_length
        self.spatial_dim = spatial_dim
        self.domain_bounds = np.array([bounds[:, 0].min() * np.ones(spatial_dim),
                                       bounds[:, 1].max() * np.ones(spatial_dim)]).T
        self.goals = list(goals) if goals is not None else None
        self.sampler = qmc.Sobol(d=spatial_dim)

    def sample(self, num_points: int = 10, method: str = 'sobol') -> np.ndarray:
        if method != 'sobol':
            raise NotImplementedError(f'{method} not implemented')
        u = self.sampler.random(num_points * self.traj_len * self.n_agents)
        pts = qmc.scale(u, self.domain_bounds[:, 0], self.domain_bounds[:, 1])
        if self.goals is not None:
            pts = pts.reshape((num_points, self.traj_len, self.spatial_dim, self.n_agents))
            for i in range(self.n_agents):
                noise = np.random.uniform(-1, 1, size=(num_points, self.spatial_dim))
                goal = np.array(self.goals[(2 * i + 1) % len(self.goals)])
                pts[:, -1, :, i] = goal + 0.1 * noise
        return pts.reshape(num_points, self.traj_len * self.spatial_dim * self.n_agents)
