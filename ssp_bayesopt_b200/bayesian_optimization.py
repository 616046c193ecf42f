"""BO driver with the reference's surface, acquisition maximised by fused candidate scoring on the GPU.

Mirrors `ssp_bayes_opt/bayesian_optimization.py:21-333`: same constructor, `initialize_agent`,
`maximize(init_points, n_iter, num_restarts, agent_type, **kwargs)`, attributes `xs ys times full_times
memory total_time agt`, properties `res`, `max`, the `log_and_plot_f` hook, and the rule that `f` is
called as `f(x_2d)` when it takes one positional argument, else `f(x_2d, info_str)` (:51-54).

What changed, on purpose: the reference maximises the acquisition with multi-restart
L-BFGS in phi-space and decodes phi* (:249-282).  Here each step scores a whole candidate set with the fused
kernel (encode -> BLR-UCB/MI -> argmax) and takes the arg-max candidate; `num_restarts` and `use_jac`
are accepted and ignored.  Candidate sets (extra kwargs, all optional):
    acq_candidates      'grid' (default) | 'uniform' | ndarray (N, data_dim)
    candidates_per_dim  grid resolution per axis (default: the agent's decoding-grid rule)
    n_candidates        size of the 'uniform' set (counter-based uniform in bounds, float32)
    candidate_seed      seed of the 'uniform' set
    candidate_chunk     candidates per kernel launch (default 1<<22)
    acq_optimizer       'candidates' (default) | 'phi-ascent': the reference's own route (:249-287) on the device --
                        `num_restarts` phi-space climbs of `acquisition_func`'s objective per step
                        (`SSPAgent.maximize_acquisition`), the best one decoded with the agent's decoder.  Starting
                        points come from `np.random.default_rng(random_state)`; every rank computes the same step.
    ascent_iters, ascent_tol   step cap (default 200) and stopping tolerance (default 1e-5: far below the resolution of the decoding grid) of 'phi-ascent'
Supported agent types: 'ssp-hex', 'ssp-rand', 'ssp-custom' (with ssp_space=...), 'ssp-traj', 'ssp-multi'.  Under
torch.distributed (one process per GPU) the candidate index range is sharded across ranks.
"""
from __future__ import annotations

import logging
import time
from typing import Callable

import numpy as np

from . import agents, dist, sspspace

_BO_KW = ('acq_candidates', 'candidates_per_dim', 'n_candidates', 'candidate_seed', 'candidate_chunk',
          'acq_optimizer', 'ascent_iters', 'ascent_tol')


class BayesianOptimization:
    def __init__(self, f: Callable[..., float] = None, bounds: np.ndarray = None,
                 log_and_plot_f: Callable[..., float] = None, random_state: int = None,
                 verbose: bool = False, sampling_seed: int = None):
        assert f is not None, 'Must specify a callable target function'
        assert bounds is not None, 'Must specify input bounds'
        assert bounds.shape[1] == 2, 'bounds must have shape (n_dims, 2)'
        if random_state is not None:
            np.random.seed(random_state)
        self.random_state = random_state
        if hasattr(f, '__code__') and f.__code__.co_argcount == 1:
            self.target = lambda x, info=None: f(x)
        else:
            self.target = f
        self.sampling_seed = sampling_seed
        self.bounds = bounds
        self.log_and_plot_f = log_and_plot_f
        self.verbose = verbose
        self.data_dim = self.bounds.shape[0]
        self.num_decoding = 10000
        self.xs = None
        self.ys = None
        self.logger = logging.getLogger(__name__)

    # ------------------------------------------------------------------ agent construction
    def initialize_agent(self, init_points=10, agent_type='ssp-hex', **kwargs):
        """Sample / adopt the initial design, evaluate the target, build the agent (:67-185)."""
        num_perimeter_samples = kwargs.pop('num_perimeter_samples', 0)
        if 'traj' in agent_type:
            domain = agents.domains.TrajectoryDomain(kwargs['traj_len'], kwargs['x_dim'], self.bounds)
        elif 'multi' in agent_type:
            domain = agents.domains.MultiTrajectoryDomain(kwargs['n_agents'], kwargs['traj_len'], kwargs['x_dim'],
                                                          self.bounds, kwargs.pop('goals', None))
        else:
            domain = agents.domains.BoundedDomain(self.bounds)
        if isinstance(init_points, (int, np.integer)):
            init_xs = domain.sample(int(init_points))
        else:
            init_xs = np.copy(init_points[0])
        n_init = init_xs.shape[0]
        if isinstance(init_points, tuple) and init_points[1] is not None:
            init_ys = init_points[1]
        else:
            init_ys = np.array([self.target(np.atleast_2d(x), str(i)) for i, x in enumerate(init_xs)]
                               ).reshape((n_init, -1))
        if dist.is_distributed():
            # the posterior is replicated, never reconciled: every rank must start from rank 0's initial design (the
            # domains' Sobol samplers are unseeded, agents/domains.py:30) and, below, from its perimeter draws
            import torch.distributed as td
            box = [init_xs, init_ys]
            td.broadcast_object_list(box, src=0)
            init_xs, init_ys = box
            n_init = init_xs.shape[0]

        if agent_type == 'ssp-hex':
            space = sspspace.HexagonalSSPSpace(self.data_dim, **kwargs)
            agt = agents.SSPAgent(init_xs, init_ys, space, **kwargs)
        elif agent_type == 'ssp-rand':
            space = sspspace.RandomSSPSpace(self.data_dim, **kwargs)
            agt = agents.SSPAgent(init_xs, init_ys, space, **kwargs)
        elif agent_type == 'ssp-custom':
            assert 'ssp_space' in kwargs
            kw = {k: v for k, v in kwargs.items() if k != 'ssp_space'}
            kw.setdefault('length_scale', kwargs['ssp_space'].length_scale)
            agt = agents.SSPAgent(init_xs, init_ys, kwargs['ssp_space'], **kw)
        elif 'rff' in agent_type:                                    # random-Fourier-feature baseline (:152-153), features + BLR on the GPU
            agt = agents.RFFAgent(init_xs, init_ys, **{k: v for k, v in kwargs.items() if k != 'domain_bounds'})
        elif agent_type == 'ssp-traj':
            agt = agents.SSPTrajectoryAgent(init_xs, init_ys, **kwargs)
            init_xs, init_ys = agt.init_xs, agt.init_ys
        elif agent_type == 'ssp-multi':
            agt = agents.SSPMultiAgent(init_xs, init_ys, **kwargs)
            init_xs, init_ys = agt.init_xs, agt.init_ys
        else:
            raise NotImplementedError(
                f'{agent_type} agent is not part of the GPU hot path (supported: ssp-hex, ssp-rand, ssp-custom, ssp-traj, '
                'ssp-multi, rff-*)')
        if num_perimeter_samples > 0:                                # corners of the box as zero-valued prior samples (:173-182)
            perimeter_xs = []
            for _ in range(num_perimeter_samples):
                idxs = np.random.choice(self.bounds.shape[1], size=(self.bounds.shape[0],))
                perimeter_xs.append([self.bounds[r, c] for r, c in enumerate(idxs)])
            perimeter_xs = np.atleast_2d(np.squeeze(perimeter_xs))
            if dist.is_distributed():
                import torch.distributed as td
                box = [perimeter_xs]
                td.broadcast_object_list(box, src=0)
                perimeter_xs = box[0]
            agt.update(perimeter_xs, np.zeros((perimeter_xs.shape[0], 1)), 0, 0)
        self.logger.info(f'{type(agt).__name__} agent created')
        return agt, init_xs, init_ys

    # ------------------------------------------------------------------ candidate sets
    def _build_candidates(self, agt, acq_candidates, candidates_per_dim, n_candidates, candidate_seed):
        """Returns (kind, payload, total).  Each rank materialises only its shard on its GPU."""
        eng = agt._scoring_engine()
        rank, world = dist.rank_world()
        lo, hi = self.bounds[:, 0], self.bounds[:, 1]
        if isinstance(acq_candidates, np.ndarray):
            total = acq_candidates.shape[0]
            s, e = dist.shard_range(total, rank, world)
            host = np.asarray(acq_candidates)
            return dict(total=total, start=s, x=eng.to_device(acq_candidates[s:e]),
                        lookup=lambda i: np.asarray(host[i], dtype=np.float64))
        if isinstance(agt, (agents.SSPTrajectoryAgent, agents.SSPMultiAgent)):
            if acq_candidates == 'grid':
                acq_candidates = 'uniform'                           # a grid over L*x_dim axes is not meaningful
            waypoints = agt.traj_len * getattr(agt, 'n_agents', 1)
            box = np.asarray(agt.domain_bounds, dtype=np.float64)
            per_row = getattr(agt, 'n_agents', 1) * agt.x_dim
            if not getattr(agt, '_shared_x_space', True) and box.shape[0] >= per_row:   # one box per agent (rows i*x_dim ...)
                lo, hi = np.tile(box[:per_row, 0], agt.traj_len), np.tile(box[:per_row, 1], agt.traj_len)
            else:
                box = box[:agt.x_dim]
                lo, hi = np.tile(box[:, 0], waypoints), np.tile(box[:, 1], waypoints)
        if acq_candidates == 'grid':
            if candidates_per_dim is None:
                candidates_per_dim = int(np.min([100, int(np.ceil((1e7 / agt.ssp_dim) ** (1 / self.data_dim)))]))
            axes = [np.linspace(self.bounds[i, 0], self.bounds[i, 1], int(candidates_per_dim))
                    for i in range(self.data_dim)]
            total = int(candidates_per_dim) ** self.data_dim
            s, e = dist.shard_range(total, rank, world)
            counts = [len(a) for a in axes]

            def grid_row(i):                                         # the reference's meshgrid('xy') order, on the host
                out = np.empty(len(axes))
                p = int(i)
                for a in range(len(axes) - 1, 1, -1):
                    out[a] = axes[a][p % counts[a]]
                    p //= counts[a]
                out[0] = axes[0][p % counts[0]]
                p //= counts[0]
                if len(axes) >= 2:
                    out[1] = axes[1][p]
                return out
            return dict(total=total, start=s, x=eng.grid_points(axes, s, e - s), lookup=grid_row)
        if acq_candidates == 'uniform':
            total = int(n_candidates or 100000)
            s, e = dist.shard_range(total, rank, world)
            torch = eng.torch
            u = eng.fill_uniform(candidate_seed, s, e - s, len(lo))
            span = torch.as_tensor((hi - lo).astype(np.float32), device=u.device)
            off = torch.as_tensor(lo.astype(np.float32), device=u.device)
            u.mul_(span).add_(off)                                   # float32 candidates inside the bounds
            return dict(total=total, start=s, x=u, lookup=None)
        raise ValueError(f'unknown acq_candidates={acq_candidates!r}')

    def _select(self, agt, cand, chunk):
        """One acquisition step: fused scoring of this rank's shard, key all-reduce, x* lookup."""
        eng = agt._scoring_engine()
        x, start = cand['x'], cand['start']
        n_local = x.shape[0]
        key = None
        first = True
        for c0 in range(0, max(n_local, 1), chunk):
            if n_local == 0:
                break
            key = agt.best_candidate(x[c0:c0 + chunk], index0=start + c0, reset_best=first)
            first = False
        if key is None:
            eng._best.zero_()
            key = eng._best
        score, gidx = dist.reduce_best(eng)            # one process: eng.best(); one per GPU: key exchange over peer memory
        if cand.get('lookup') is not None:             # every rank can name the winning row without touching the device
            return score, gidx, np.asarray(cand['lookup'](gidx), dtype=np.float64).reshape(1, -1)
        # owner of the winning row publishes it (one small broadcast; single-GPU: a D2H copy of one row)
        rank, world = dist.rank_world()
        owner_has = start <= gidx < start + n_local
        torch = eng.torch
        row = x[gidx - start].to(torch.float64) if owner_has else torch.zeros(x.shape[1], dtype=torch.float64,
                                                                               device=x.device)
        if world > 1:
            import torch.distributed as td
            td.all_reduce(row, op=td.ReduceOp.SUM)                   # exactly one rank contributes non-zeros
        return score, gidx, row.cpu().numpy().reshape(1, -1)

    # ------------------------------------------------------------------ BO loop
    def maximize(self, init_points=10, n_iter: int = 100, num_restarts: int = 5, agent_type='ssp-hex',
                 save_memory=True, use_jac=True, **kwargs):
        bo_kw = {k: kwargs.pop(k) for k in _BO_KW if k in kwargs}
        acq_candidates = bo_kw.get('acq_candidates', 'grid')
        chunk = int(bo_kw.get('candidate_chunk', 1 << 22))
        kwargs.setdefault('decoder_method', 'from-set')

        agt, init_xs, init_ys = self.initialize_agent(init_points, agent_type, domain_bounds=self.bounds, **kwargs)
        self.agt = agt
        optimizer = bo_kw.get('acq_optimizer', 'candidates')
        if isinstance(agt, agents.RFFAgent):
            optimizer = 'x-lbfgs'                                    # the reference's route for this agent (:252-262)
        if optimizer not in ('candidates', 'phi-ascent', 'x-lbfgs'):
            raise ValueError(f'unknown acq_optimizer={optimizer!r}')
        cand = None
        if optimizer == 'candidates':
            cand = self._build_candidates(agt, acq_candidates, bo_kw.get('candidates_per_dim'),
                                          bo_kw.get('n_candidates'), int(bo_kw.get('candidate_seed', 0)))
        ascent_rng = np.random.default_rng(self.random_state)

        n0 = init_xs.shape[0]
        self.times = np.zeros((n_iter,))
        self.full_times = np.zeros((n_iter,))
        self.memory = np.zeros((n_iter, 1))
        self.xs = np.zeros((n_iter + n0, init_xs.shape[1]))
        self.ys = np.zeros((n_iter + n0,))
        self.xs[:n0] = init_xs
        self.ys[:n0] = np.ravel(init_ys)
        self.acq_values = np.zeros((n_iter,))
        self.acq_indices = np.zeros((n_iter,), dtype=np.int64)
        rank, _ = dist.rank_world()

        full_start = time.perf_counter_ns()
        for t in range(n_iter):
            start = time.perf_counter_ns()
            if optimizer == 'candidates':
                score, gidx, x_t = self._select(agt, cand, chunk)    # synchronises on the 8-byte key
            elif optimizer == 'x-lbfgs':
                # multi-restart L-BFGS-B in x on the agent's own objective / gradient (host closures over copies of m, S), as
                # the reference does for its GP / RFF agents; eval and update of the chosen point run on the device
                from scipy.optimize import minimize
                optim_func, jac_func = agt.acquisition_func()
                vals, solns = np.zeros(num_restarts), np.zeros((num_restarts, init_xs.shape[1]))
                for r in range(num_restarts):
                    x_init = np.random.uniform(low=self.bounds[:, 0], high=self.bounds[:, 1], size=(self.bounds.shape[0],))
                    soln = minimize(optim_func, x_init, jac=jac_func if use_jac else None, method='L-BFGS-B', bounds=self.bounds)
                    vals[r], solns[r] = -float(np.ravel(soln.fun)[0]), soln.x
                gidx = int(np.argmax(vals))
                score, x_t = vals[gidx], np.atleast_2d(solns[gidx])
            else:
                phis, vals = agt.maximize_acquisition(num_restarts=num_restarts, rng=ascent_rng,
                                                      max_iter=int(bo_kw.get('ascent_iters', 200)),
                                                      tol=float(bo_kw.get('ascent_tol', 1e-5)))
                gidx = int(np.argmax(vals))                          # all restarts (not the reference's vals[:(t+1)] slice)
                score = vals[gidx]
                x_t = np.asarray(agt.decode(phis[gidx:gidx + 1]), dtype=np.float64).reshape(1, -1)
            self.times[t] = time.perf_counter_ns() - start
            self.acq_values[t], self.acq_indices[t] = score, gidx

            status = f'{t + n0}'
            y_t = np.atleast_2d(self.target(x_t, status))
            if dist.is_distributed():                                 # keep replicas identical even if f is noisy
                import torch
                import torch.distributed as td
                yb = torch.as_tensor(y_t, dtype=torch.float64, device=agt._scoring_engine()._dev())
                td.broadcast(yb, src=0)
                y_t = yb.cpu().numpy()

            update_start = time.perf_counter_ns()
            # eval(x_t) under the current posterior, then update(x_t, y_t, var_t) (:299-303): one device round trip
            pure_ucb = isinstance(getattr(agt, 'gamma_c', None), (int, float)) and agt.gamma_c == 0
            if pure_ucb and not self.verbose and hasattr(agt, 'observe'):
                # gamma_t += 0 * var_t: nothing of eval(x_t) is used, so the update is only enqueued -- no read-back, no second
                # synchronisation in the step; the next selection queues behind it on the stream
                agt.update(x_t, y_t, 0.0, step_num=t + n0)
                mu_t = var_t = phi_t = None
            elif hasattr(agt, 'observe'):
                mu_t, var_t, phi_t = agt.observe(x_t, y_t, step_num=t + n0)
            else:
                mu_t, var_t, phi_t = agt.eval(x_t)
                agt.update(x_t, y_t, var_t, step_num=t + n0)
            if self.verbose and rank == 0:
                print(f'| step {t + n0}\t | {y_t}, {np.sqrt(var_t)}, {phi_t}\t ')
            self.times[t] += time.perf_counter_ns() - update_start
            self.full_times[t] = time.perf_counter_ns() - start

            t_now = t + n0
            self.xs[t_now] = np.copy(x_t)
            self.ys[t_now] = np.ravel(y_t)[0]
            if self.log_and_plot_f is not None:
                self.log_and_plot_f(np.vstack(self.xs[:t_now + 1]), np.vstack(self.ys[:t_now + 1]),
                                    times=self.times, trial=t_now, memory=self.memory)
        self.total_time = time.perf_counter_ns() - full_start

    @property
    def res(self):
        return [{'target': t, 'params': p} for t, p in zip(self.ys, self.xs)]

    @property
    def max(self):
        i = np.argmax(self.ys)
        return {'target': self.ys[i], 'params': self.xs[i]}
