"""Lock-step stepping of many independent BLR-UCB runs on one GPU (BASELINE config C4).

Each run is an ordinary `SSPAgent` (own SSP space / posterior / scoring factor in its own device context); all
runs share one candidate set resident on the device.  A lock-step = for every run: fused scoring of the shared
candidates, arg-max, then (after the host evaluated the objectives) one float64 rank-one posterior update.
Scoring launches of different runs are issued round-robin on a small pool of CUDA streams so the many small
kernels overlap; the winners are gathered with ONE device-to-host copy per lock-step.

Runs are independent, so multi-GPU use is plain partitioning of the run list (`dist.shard_range` over run ids),
with no collective.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .agents import SSPAgent


class BatchedRuns:
    def __init__(self, agents: list[SSPAgent], candidates, n_streams: int = 8):
        assert len(agents) > 0
        self.agents = agents
        eng0 = agents[0]._scoring_engine()
        self.torch = eng0.torch
        self.device = eng0._dev()
        self.candidates = eng0.to_device(candidates)            # shared, resident
        self.streams = [self.torch.cuda.Stream(device=self.device) for _ in range(max(1, min(n_streams, len(agents))))]
        self._keys = self.torch.zeros(len(agents), dtype=self.torch.int64, device=self.device)

    def select(self):
        """Arg-max candidate of every run.  Returns (scores (R,), indices (R,), points (R, n))."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for r, agt in enumerate(self.agents):
            st = self.streams[r % len(self.streams)]
            with torch.cuda.stream(st):
                key = agt.best_candidate(self.candidates)
                self._keys[r:r + 1].copy_(key)
        for s in self.streams:
            cur.wait_stream(s)
        keys = self._keys.cpu().numpy()                          # one D2H per lock-step
        unpacked = [_lib.key_unpack(int(k)) for k in keys]
        scores = np.array([u[0] for u in unpacked])
        idx = np.array([u[1] for u in unpacked], dtype=np.int64)
        pts = self.candidates[torch.as_tensor(idx, device=self.device)].double().cpu().numpy()
        return scores, idx, pts

    def update(self, xs, ys):
        """xs (R, n), ys (R,): one observation per run."""
        for r, agt in enumerate(self.agents):
            x_t = np.atleast_2d(xs[r])
            _, var_t, _ = agt.eval(x_t)
            agt.update(x_t, np.atleast_2d(ys[r]), var_t)

    def step(self, objective):
        """One lock-step with a vectorised objective f(points (R, n)) -> (R,)."""
        scores, idx, pts = self.select()
        ys = np.asarray(objective(pts)).reshape(-1)
        self.update(pts, ys)
        return pts, ys, scores
