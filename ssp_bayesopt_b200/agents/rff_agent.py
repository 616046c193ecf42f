"""Random-Fourier-feature baseline agent (`ssp_bayes_opt/agents/rff_agent.py`) with its hot path on the
GPU: the feature map sqrt(2 alpha / m) cos(W x + b) (`rff_agent.py:153-156`, `sspbo_rff_encode`) and the float64 BLR over it
(`sspbo_blr_update` / `sspbo_blr_predict` on a raw-feature context).  Set-up is the reference's, on the host and in its order:
scikit-learn GP fit of a scaled RBF / Matern kernel on the initial design (a handful of points), then W ~ N(0, 1/l) or
Student-t and b ~ U(0, 2 pi) from numpy's global generator (`rff_agent.py:43-73`), so seeded runs draw the same features.
The fused tensor-core scorer is not used: these features are cosines only (no conjugate-symmetric spectrum), so there is no
psi-space in which the prior is diagonal; candidate sets are evaluated as encode -> predict, all on the device."""
from __future__ import annotations

import numpy as np

from .. import blr
from .agent import Agent


class RFFAgent(Agent):
    def __init__(self, init_xs, init_ys, ssp_dim, kernel_type='rbf', gamma_c=1.0, beta_ucb=np.log(2 / 1e-6), var_decay=0.,
                 device=None, features=None, **kwargs):
        super().__init__()
        from sklearn.gaussian_process import GaussianProcessRegressor
        from sklearn.gaussian_process.kernels import Matern, ConstantKernel, RBF
        import scipy.stats
        (num_pts, data_dim) = init_xs.shape
        self.data_dim = data_dim
        self.init_xs = init_xs
        self.init_ys = init_ys
        self.dim = ssp_dim
        if features is not None:                # extension: adopt a given feature map (W (m, n), b (m, 1), scale), no GP fit
            self.W, self.B = np.asarray(features[0], dtype=np.float64), np.asarray(features[1], dtype=np.float64).reshape(-1, 1)
            self.sqrt_2_alpha_over_m = float(features[2])
            self.lengthscales, self.scaling = None, None
            assert self.W.shape == (self.dim, data_dim) and self.B.shape == (self.dim, 1)
            self._finish_init(init_xs, init_ys, gamma_c, beta_ucb, var_decay, device)
            return
        k_scaling = ConstantKernel(1, (0.1, 10))
        if kernel_type == 'matern':
            k_cov = Matern(nu=2.5, length_scale_bounds=(1 / np.sqrt(init_xs.shape[0] + 1), 1e5))
            kernel_nu = 2.5
        elif kernel_type == 'rbf':
            k_cov = RBF(1, length_scale_bounds=(1 / np.sqrt(init_xs.shape[0] + 1), 1e5))
            kernel_nu = np.inf
        else:
            raise NotImplementedError(f'kernel_type {kernel_type!r} not supported')
        fit_gp = GaussianProcessRegressor(kernel=k_scaling * k_cov, alpha=1e-6, normalize_y=True, n_restarts_optimizer=20,
                                          random_state=0)
        fit_gp.fit(self.init_xs, self.init_ys)
        lengthscales = fit_gp.kernel_.k2.length_scale
        scaling = fit_gp.kernel_.k1.constant_value
        self.lengthscales, self.scaling = lengthscales, scaling
        alpha = scaling ** 2
        self.sqrt_2_alpha_over_m = np.sqrt(2 * alpha / self.dim)
        if np.isinf(kernel_nu):
            self.W = np.random.normal(loc=0, scale=1 / lengthscales, size=(self.dim, data_dim))
        else:
            self.W = scipy.stats.t.rvs(loc=0, scale=1 / lengthscales, df=kernel_nu, size=(self.dim, data_dim))
        self.B = np.random.uniform(0, 2 * np.pi, size=(self.dim, 1))
        self._finish_init(init_xs, init_ys, gamma_c, beta_ucb, var_decay, device)

    def _finish_init(self, init_xs, init_ys, gamma_c, beta_ucb, var_decay, device):
        self.blr = blr.BayesianLinearRegression(self.dim, device=device)       # raw-feature context (no SSP space)
        eng = self.blr.engine
        self._W_dev = eng.to_device(np.ascontiguousarray(self.W), eng.torch.float64)
        self._B_dev = eng.to_device(np.ascontiguousarray(self.B.reshape(-1)), eng.torch.float64)
        self.blr.update(self._encode_device(init_xs), np.array(init_ys))
        self.constraint_ssp = np.zeros((self.dim, 1))
        self.gamma_t = 0
        self.gamma_c = gamma_c
        self.var_weight = beta_ucb
        self.var_decay = var_decay

    # ------------------------------------------------------------------ hot path
    def _scoring_engine(self):
        return self.blr.engine

    def _encode_device(self, x):
        return self.blr.engine.rff_encode_device(self._W_dev, self._B_dev, self.sqrt_2_alpha_over_m, np.atleast_2d(x))

    def encode(self, x):
        """(n, data_dim) -> (n, dim) float64: sqrt(2 alpha / m) cos(W x + b) (rff_agent.py:153-156)."""
        return self._encode_device(np.asarray(x, dtype=np.float64)).cpu().numpy()

    def eval(self, xs):
        """mu (1, n), var (n,), acq (n,) (rff_agent.py:85-89); features and predictive moments stay on the device."""
        mu, var = self.blr.engine.blr_predict(self._encode_device(np.asarray(xs, dtype=np.float64)))
        mu, var = mu.cpu().numpy().reshape(1, -1), var.cpu().numpy()
        phi = self.var_weight * (np.sqrt(var + self.gamma_t) - np.sqrt(self.gamma_t))
        return mu, var, phi

    def initial_guess(self):
        return self.blr.sample()

    def acquisition_func(self):
        """(min_func, gradient) in x space on host copies of the posterior (rff_agent.py:95-125)."""
        m, sigma, beta_inv, gamma = self.blr.m, self.blr.S, 1 / self.blr.beta, self.gamma_t
        W, B, c, lam = self.W, self.B, self.sqrt_2_alpha_over_m, self.var_weight

        def host_encode(x):
            return c * np.cos(np.dot(W, np.atleast_2d(x).T) + B).T

        def min_func(x, m=m, sigma=sigma, gamma=gamma, beta_inv=beta_inv, norm_margin=4):
            phi = host_encode(x).T
            val = phi.T @ m
            mi = lam * np.sqrt(gamma + beta_inv + phi.T @ sigma @ phi) - np.sqrt(gamma)
            return -(val + mi).flatten()

        def gradient(x, m=m, sigma=sigma, gamma=gamma, beta_inv=beta_inv):
            x = np.atleast_2d(x)
            Wx_plus_B = np.dot(W, x.T) + B
            phi_x = c * np.cos(Wx_plus_B)
            sin_Wx_plus_B = np.sin(Wx_plus_B)
            sigma_phi = sigma @ phi_x
            sigma_x = np.sqrt(beta_inv + phi_x.T @ sigma_phi)
            grad_mu = -c * ((m * sin_Wx_plus_B).T @ W)
            grad_sigma = -(c / sigma_x) * ((sigma_phi * sin_Wx_plus_B).T @ W)
            return -(grad_mu + lam * grad_sigma).flatten()

        return min_func, gradient

    def update(self, x_t: np.ndarray, y_t: np.ndarray, sigma_t: float, step_num=0):
        x_val, y_val = x_t, y_t
        if len(x_t.shape) < 2:
            x_val = x_t.reshape(1, x_t.shape[0])
            y_val = y_t.reshape(1, y_t.shape[0])
        self.blr.update(self._encode_device(np.asarray(x_val, dtype=np.float64)), np.asarray(y_val, dtype=np.float64))
        if isinstance(self.var_decay, (int, float)):
            self.var_weight = self.var_weight + self.var_decay
        elif callable(self.var_decay):
            self.var_weight = self.var_weight + self.var_decay(step_num)
        else:
            raise RuntimeError(f'unable to use {self.var_weight}, expected number or callable')
        if isinstance(self.gamma_c, (int, float)):
            self.gamma_t = self.gamma_t + self.gamma_c * sigma_t
        elif callable(self.gamma_c):
            self.gamma_t = self.gamma_t + self.gamma_c(step_num) * sigma_t
        else:
            raise RuntimeError(f'unable to use {self.gamma_c}, expected number or callable')

    def untrusted(self, x, badness=-1):
        self.constraint_ssp += badness * self.encode(x).reshape(-1, 1)

    def length_scale(self):
        return self.lengthscales
