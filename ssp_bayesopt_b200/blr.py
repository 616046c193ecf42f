"""Bayesian linear regression with the posterior resident on the GPU in float64.

Mirrors `ssp_bayes_opt/blr.py:6-116` of the reference: same constructor, `update(phis, ts)`,
`predict(phi) -> (mu (1,n), var (n,))`, attributes `m (d,1)`, `S (d,d)`, `S_inv`, `beta`, `input_dim`.
The Sherman-Morrison recursion (blr.py:63-67), the information-form accumulator (blr.py:60) and
m = S (S_inv_old m_old + beta Phi^T t) (blr.py:75, kept as m = S b with b = beta sum phi_i t_i) run in
the float64 kernels of libsspbo.so (`sspbo_blr_update`); attributes are device->host copies made on read.

Differences, on purpose: the unbounded posterior history `ms`/`Ss` (blr.py:34-37, never read by the
reference) is not kept; `sample()` (blr.py:99-116, an O(d^3) host SVD used only to seed L-BFGS) draws on
the host from the copied posterior.
"""
from __future__ import annotations

import weakref

import numpy as np

from . import _lib
from .engine import DeviceEngine


class BayesianLinearRegression:
    def __init__(self, size_in, size_out=1, alpha=0.01, beta=None, engine: DeviceEngine | None = None, device=None):
        self.input_dim = size_in
        self.output_dim = size_out
        assert size_out == 1, "only size_out=1 is supported (as in the reference)"
        self.alpha = alpha
        self.rank_one_updates = True
        if engine is None:                      # stand-alone BLR over raw features
            engine = DeviceEngine(device)
            engine.blr_init(alpha, dim=size_in)
        else:                                   # BLR over an SSP space: also maintains the scoring factor
            assert engine.d == size_in, f"engine encodes {engine.d}-d SSPs, BLR asked for {size_in}"
            owner = engine._blr_owner() if engine._blr_owner is not None else None
            if owner is not None and owner is not self:
                # a device context holds ONE posterior: initialising a second BLR on it would silently wipe the first one's
                # m / S / V and leave both models updating the same state (in the reference every BLR is independent)
                raise _lib.SspboError("this DeviceEngine already hosts the posterior of a live BayesianLinearRegression; "
                                      "give every BLR its own engine (SSPAgent does so for agents that share an SSPSpace)")
            engine.blr_init(alpha)
        engine._blr_owner = weakref.ref(self)
        self.engine = engine
        self.beta = beta
        if beta is not None:
            engine.blr_set_beta(float(np.ravel(beta)[0]))

    # -- posterior attributes (device -> host copies) --
    @property
    def m(self):
        return self.engine.blr_get(m=True, S=False)[0].cpu().numpy().reshape(-1, 1)

    @property
    def S(self):
        return self.engine.blr_get(m=False, S=True)[1].cpu().numpy()

    @property
    def S_inv(self):
        return self.engine.blr_get(m=False, S=False, Sinv=True)[2].cpu().numpy()

    def update(self, phis: np.ndarray, ts: np.ndarray):
        """phis (n, input_dim), ts (n, 1); shape errors as blr.py:47-52."""
        phis_shape = tuple(phis.shape)
        ts = np.asarray(ts, dtype=np.float64)
        assert phis_shape[1] == self.input_dim, (
            f'Expected input shape ({ts.shape[0]}, {self.input_dim}), got {phis_shape}')
        assert len(ts.shape) > 1 and ts.shape[1] == self.output_dim, (
            f'Expected output shape ({phis_shape[0]}, 1), got {ts.shape}')
        if self.beta is None:                   # blr.py:54-59: noise precision from the first batch only
            var_ts = np.var(ts) if len(ts) > 1 else np.abs(np.copy(ts[0]))
            self.beta = 1. / var_ts
            self.engine.blr_set_beta(float(np.ravel(self.beta)[0]))
        self.engine.blr_update(phis, ts)

    def update_points(self, xs: np.ndarray, ts: np.ndarray):
        """update(encode(xs), ts) for host points of the engine's SSP space without materialising phi on the host (the
        device encodes them): SSPAgent.update's path.  xs (n, data_dim) float64, ts (n, 1)."""
        ts = np.asarray(ts, dtype=np.float64)
        assert len(ts.shape) > 1 and ts.shape[1] == self.output_dim, (
            f'Expected output shape ({xs.shape[0]}, 1), got {ts.shape}')
        if self.beta is None:                   # blr.py:54-59
            var_ts = np.var(ts) if len(ts) > 1 else np.abs(np.copy(ts[0]))
            self.beta = 1. / var_ts
            self.engine.blr_set_beta(float(np.ravel(self.beta)[0]))
        self.engine.blr_update_points(xs, ts)

    def predict(self, phi: np.ndarray):
        """mu (1, n), var (n,) for explicit features phi (n, input_dim) (blr.py:84-97)."""
        if self.beta is None:
            raise TypeError("predict() before the first update(): beta is None (blr.py:96 divides by it)")
        mu, var = self.engine.blr_predict(phi)
        return mu.cpu().numpy().reshape(1, -1), var.cpu().numpy()

    def sample(self) -> np.ndarray:
        m, S = self.m, self.S
        try:
            return np.atleast_2d(np.random.multivariate_normal(m.flatten(), S).reshape(-1, 1))
        except np.linalg.LinAlgError as e:      # pragma: no cover - mirrors blr.py:111-113
            print(e)
            return (-self.S_inv @ m).reshape(-1, 1)
