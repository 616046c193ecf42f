"""Lock-step stepping of many independent BLR-UCB runs on one GPU (BASELINE config C4).

Each run is an ordinary `SSPAgent` (own SSP space / posterior / scoring factor in its own device context); all
runs share one candidate set resident on the device.  A lock-step = for every run: fused scoring of the shared
candidates, arg-max, then (after the host evaluated the objectives) one float64 rank-one posterior update.
The scoring launches of all runs are enqueued by ONE library call (`sspbo_score_batch`: a C loop over the runs' contexts,
round-robin on a small pool of CUDA streams so the many small kernels overlap), likewise the encode + posterior updates
(`sspbo_update_batch`); the winners are gathered with ONE device-to-host copy per lock-step.

Runs are independent, so multi-GPU use is plain partitioning of the run list (`dist.shard_range` over run ids),
with no collective.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .agents import SSPAgent


def _unpack_keys(keys):
    """numpy form of sspbo_key_unpack for an int64 array of packed keys: (scores float64, indices int64).
    key = orderable(score) << 32 | (0xFFFFFFFF - index); orderable(u) = ~u for negative floats, u | 0x80000000 otherwise."""
    k = np.asarray(keys).astype(np.int64).view(np.uint64)
    ob = (k >> np.uint64(32)).astype(np.uint32)
    bits = np.where(ob & np.uint32(0x80000000), ob & np.uint32(0x7FFFFFFF), ~ob).astype(np.uint32)
    scores = bits.view(np.float32).astype(np.float64)
    scores[ob == np.uint32(0xFFFFFFFF)] = np.nan                 # NaN counts as the maximum
    scores[k == 0] = -np.inf                                     # no candidate scored
    idx = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    return scores, idx


class BatchedRuns:
    def __init__(self, agents: list[SSPAgent], candidates, n_streams: int = 8):
        assert len(agents) > 0
        self.agents = agents
        eng0 = agents[0]._scoring_engine()
        self.torch = eng0.torch
        self.device = eng0._dev()
        self.candidates = eng0.to_device(candidates)            # shared, resident
        self.streams = [self.torch.cuda.Stream(device=self.device) for _ in range(max(1, min(n_streams, len(agents))))]
        self._keys = self.torch.zeros(2 * len(agents), dtype=self.torch.int64, device=self.device)   # R keys, R status words
        self._step_num = 0
        # per-run contexts and streams as C arrays: one library call enqueues a whole lock-step
        R = len(agents)
        self._engines = [a._scoring_engine() for a in agents]
        self._lib = self._engines[0].lib
        self._ctxs = (C.c_void_p * R)(*[e._ctx.value for e in self._engines])
        self._cstreams = (C.c_void_p * len(self.streams))(*[st.cuda_stream for st in self.streams])
        self._width = self._engines[0].n * self._engines[0].L
        assert all(e.n * e.L == self._width for e in self._engines), "runs must share the candidate row width"
        assert self.candidates.shape[1] == self._width
        self._read_scalars()

    def _read_scalars(self):
        """Acquisition scalars of every run as arrays (lambda, gamma_t): read from the agents once, then stepped here
        (vectorised) and written back, so a lock-step of thousands of runs does not pay thousands of method calls."""
        self._lam = np.array([a.var_weight for a in self.agents], dtype=np.float64)
        self._gam = np.array([a._gamma() for a in self.agents], dtype=np.float64)
        # pure UCB with a constant weight step: gamma_t ignores the predictive variance and lambda += var_decay
        self._pure_ucb = all(isinstance(a.gamma_c, (int, float)) and a.gamma_c == 0 and isinstance(a.var_decay, (int, float))
                             for a in self.agents)
        self._decay = np.array([a.var_decay if isinstance(a.var_decay, (int, float)) else 0.0 for a in self.agents], dtype=np.float64)

    def select(self):
        """Arg-max candidate of every run.  Returns (scores (R,), indices (R,), points (R, n))."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        R = len(self.agents)
        lam, gam = self._lam, self._gam
        self._keys.zero_()
        for s in self.streams:
            s.wait_stream(cur)                                   # ... and the zeroed keys / status words
        failed = C.c_int(-1)
        rc = self._lib.sspbo_score_batch(self._ctxs, R, C.c_void_p(self.candidates.data_ptr()),
                                         self._engines[0]._dtype_code(self.candidates), self.candidates.shape[0],
                                         lam.ctypes.data_as(_lib._pd), gam.ctypes.data_as(_lib._pd),
                                         C.c_void_p(self._keys.data_ptr()), self._cstreams, len(self.streams), C.byref(failed))
        if rc:
            _lib.check(self._lib, self._engines[max(failed.value, 0)]._ctx, rc, f"sspbo_score_batch (run {failed.value})")
        # the 'rerun' protocol of DeviceEngine.best(): per-run overflow / out-of-bound flags, gathered next to the keys
        rc = self._lib.sspbo_score_batch_status(self._ctxs, R, C.c_void_p(self._keys.data_ptr() + 8 * R), self._cstreams,
                                                len(self.streams), C.byref(failed))
        if rc:
            _lib.check(self._lib, self._engines[max(failed.value, 0)]._ctx, rc, f"sspbo_score_batch_status (run {failed.value})")
        for s in self.streams:
            cur.wait_stream(s)
        both = self._keys.cpu().numpy()                          # one D2H per lock-step: keys and status words
        keys, status = both[:R].copy(), both[R:]
        for r in np.nonzero(status)[0]:                          # rare: recompute that run with the engines that never flag
            eng = self._engines[r]
            eng.set_policy(lowrank=False if status[r] & 1 else None, phase32=False if status[r] & 2 else None)
            eng.score(self.candidates, float(lam[r]), float(gam[r]))
            score, idx_r = eng.best()
            keys[r] = np.array(_lib.key_pack(score, idx_r), dtype=np.uint64).astype(np.int64)
        scores, idx = _unpack_keys(keys)                         # vectorised sspbo_key_unpack (thousands of runs per lock-step)
        pts = self.candidates[torch.as_tensor(idx, device=self.device)].double().cpu().numpy()
        return scores, idx, pts

    def update(self, xs, ys):
        """xs (R, n), ys (R,): one observation per run."""
        torch = self.torch
        if self._pure_ucb:
            # pure UCB: gamma_t ignores the predictive variance, so nothing is read back: one library call enqueues every
            # run's encode + posterior update
            xs = np.ascontiguousarray(xs, dtype=np.float64).reshape(len(self.agents), self._width)
            ys = np.ascontiguousarray(ys, dtype=np.float64).reshape(-1)
            cur = torch.cuda.current_stream(self.device)
            for s in self.streams:
                s.wait_stream(cur)
            failed = C.c_int(-1)
            rc = self._lib.sspbo_update_batch(self._ctxs, len(self.agents), xs.ctypes.data_as(_lib._pd), ys.ctypes.data_as(_lib._pd),
                                              self._cstreams, len(self.streams), C.byref(failed))
            if rc:
                _lib.check(self._lib, self._engines[max(failed.value, 0)]._ctx, rc, f"sspbo_update_batch (run {failed.value})")
            for s in self.streams:
                cur.wait_stream(s)
            self._lam = self._lam + self._decay                  # ssp_agent.py:208-218 with sigma_t unused (gamma_c = 0)
            for a, lam_r in zip(self.agents, self._lam.tolist()):
                a.var_weight = lam_r
            self._step_num += 1
            return
        for r, agt in enumerate(self.agents):
            agt.observe(np.atleast_2d(xs[r]), np.atleast_2d(ys[r]), step_num=self._step_num)   # eval(x_t) + update in one device round trip
        self._step_num += 1
        self._read_scalars()

    def step(self, objective):
        """One lock-step with a vectorised objective f(points (R, n)) -> (R,)."""
        scores, idx, pts = self.select()
        ys = np.asarray(objective(pts)).reshape(-1)
        self.update(pts, ys)
        return pts, ys, scores
