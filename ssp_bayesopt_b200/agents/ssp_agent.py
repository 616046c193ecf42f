"""SSP + BLR agent on the GPU.  Mirrors `ssp_bayes_opt/agents/ssp_agent.py` of the reference.

Hot path: `eval` (ssp_agent.py:125-137) and `best_candidate` call the fused kernel
(`sspbo_score`: encode -> BLR predict -> MI/UCB bonus -> argmax) and never materialise phi;
`update` (ssp_agent.py:187-224) encodes x_t and applies the float64 rank-one posterior update on the
device.  Deviations, documented in DESIGN.md: `decoder_method` defaults to 'from-set' (the reference's
default 'network-optim' needs TensorFlow).  Omitting `length_scale` fits the sinc-kernel GP of the reference on the
host (ssp_agent.py:96-109; `agents/kernels/sinc_kernel.py`).
"""
from __future__ import annotations

import numpy as np

from .. import blr, sspspace
from .._lib import ENGINE_AUTO
from .agent import Agent


class SSPAgent(Agent):
    def __init__(self, init_xs, init_ys, ssp_space, decoder_method='from-set', gamma_c=1.0,
                 beta_ucb=np.log(2 / 1e-6), var_decay=0., **kwargs):
        super().__init__()
        self._init_samples, self._init_samples_of = None, None
        (num_pts, data_dim) = init_xs.shape
        self.data_dim = data_dim
        self.init_xs = init_xs
        self.init_ys = init_ys
        self.decoder_method = decoder_method
        self.score_engine = kwargs.pop('score_engine', ENGINE_AUTO)

        self._set_ssp_space(ssp_space=ssp_space, **kwargs)
        self._set_decoder()

        init_phis = self.encode(init_xs)
        self.blr = blr.BayesianLinearRegression(self.ssp_dim, engine=self._scoring_engine())
        self.blr.update(init_phis, np.array(init_ys))
        self.constraint_ssp = np.zeros((self.ssp_dim, 1))

        self.gamma_t = 0
        self.gamma_c = gamma_c
        self.var_weight = beta_ucb
        self.var_decay = var_decay

    # ------------------------------------------------------------------ construction
    def _scoring_engine(self):
        """Device context whose `score` sees candidates in this agent's input layout and which holds this agent's posterior.
        The first agent over an SSPSpace uses the space's own context; an agent built over a space whose context already
        hosts another live agent's posterior gets a private context programmed from the same phase matrix (in the reference
        every agent owns an independent BLR and sharing a space is legal: 'ssp-custom', lists of agents)."""
        space_eng = self.ssp_space.engine
        if getattr(self, '_own_engine', None) is None:
            owner = space_eng._blr_owner() if space_eng._blr_owner is not None else None
            mine = getattr(self, 'blr', None)
            if owner is None or owner is mine:
                return space_eng
            from ..engine import DeviceEngine
            self._own_engine = DeviceEngine(self.ssp_space._device)
            self._own_engine_ls = None
        ls = np.asarray(self.ssp_space.length_scale, dtype=np.float64).reshape(-1)
        if self._own_engine_ls is None or not np.array_equal(ls, self._own_engine_ls):
            self._own_engine.set_space(self.ssp_space.phase_matrix, ls)
            if self.ssp_space.domain_bounds is not None:
                self._own_engine.set_xbound(float(np.max(np.abs(np.asarray(self.ssp_space.domain_bounds, dtype=np.float64)))))
            self._own_engine_ls = ls.copy()
        return self._own_engine

    def _set_ssp_space(self, ssp_space, seed=0, **kwargs):
        if ssp_space is None:
            ssp_space = sspspace.HexagonalSSPSpace(self.data_dim, ssp_dim=kwargs.get('ssp_dim', 100),
                                                   domain_bounds=kwargs.get('domain_bounds', None),
                                                   length_scale=1, rng=seed)
        self.ssp_space = ssp_space
        self.ssp_dim = ssp_space.ssp_dim
        if 'length_scale' not in kwargs or np.any(np.asarray(kwargs.get('length_scale')) < 0):
            self.ssp_space.update_lengthscale(self._optimize_lengthscale(self.init_xs, self.init_ys))
        else:
            self.ssp_space.update_lengthscale(kwargs.get('length_scale'))

    def _set_decoder(self):
        if self.decoder_method in ('network', 'network-optim', 'regression'):
            raise NotImplementedError(f"decoder_method={self.decoder_method!r} (learned decoders) is out of scope; "
                                      "use 'from-set' or 'direct-optim'")
        self._init_samples_of = lambda: self.get_init_samples(self.ssp_space)   # built on first use (see init_samples)

    @property
    def init_samples(self):
        """Decoding set (sample SSPs, sample points) of ssp_agent.py:86-91.  The reference builds it in the constructor;
        here it is built on first use, because it is M x ssp_dim float64 (41 MB at M = 1e4, d = 511) and runs that pick
        their points straight from a candidate set never decode (thousands of agents coexist in config C4)."""
        if self._init_samples is None and self._init_samples_of is not None:
            self._init_samples = self._init_samples_of()
        return self._init_samples

    @init_samples.setter
    def init_samples(self, value):
        self._init_samples = value

    def length_scale(self):
        return self.ssp_space.length_scale

    def _optimize_lengthscale(self, init_xs, init_ys):
        """Length scale when the caller gives none (ssp_agent.py:96-109): scikit-learn GP with the sinc kernel on the
        initial design, 20 optimiser restarts, random_state 0.  Host-side set-up on a handful of points."""
        from sklearn.gaussian_process import GaussianProcessRegressor
        from .kernels import SincKernel
        fit_gp = GaussianProcessRegressor(
            kernel=SincKernel(length_scale=1., length_scale_bounds=(1 / np.sqrt(init_xs.shape[0] + 1), 1e5)),
            alpha=1e-6, normalize_y=True, n_restarts_optimizer=20, random_state=0)
        fit_gp.fit(init_xs, init_ys)
        return np.exp(fit_gp.kernel_.theta)

    def get_init_samples(self, ssp_space):
        """Decoding set: grid size rule of ssp_agent.py:111-123."""
        ls = np.asarray(ssp_space.length_scale, dtype=np.float64).reshape(-1)
        n_ls = [2 * int(np.ceil((b[1] - b[0]) / ls[i])) for i, b in enumerate(ssp_space.domain_bounds)]
        if (np.prod(n_ls) * ssp_space.ssp_dim > 1e7) or ('optim' not in self.decoder_method):
            per_dim = np.min([100, int(np.ceil((1e7 / ssp_space.ssp_dim) ** (1 / ssp_space.domain_dim)))])
            return ssp_space.get_sample_pts_and_ssps(per_dim, 'grid')
        return ssp_space.get_sample_pts_and_ssps(1, 'length-scale')

    # ------------------------------------------------------------------ hot path
    def _gamma(self):
        return float(np.ravel(self.gamma_t)[0])

    def eval(self, xs):
        """mu (1, n), var (n,), acq (n,) with acq = var_weight (sqrt(var + gamma_t) - sqrt(gamma_t))."""
        mu, var, acq = self._scoring_engine().eval_host(xs, self.var_weight, self._gamma(), engine=self.score_engine)
        return mu.reshape(1, -1), var, acq

    def best_candidate(self, xs, index0=0, reset_best=True):
        """argmax_i mu_i + acq_i over a candidate block (host array or device tensor).  Returns the device
        key tensor (see sspbo_score); `self._scoring_engine().best()` unpacks it."""
        eng = self._scoring_engine()
        on_host = isinstance(xs, np.ndarray) or (hasattr(xs, "is_cuda") and not xs.is_cuda)
        if on_host and len(xs) >= (1 << 21):                  # large host sets: overlap the H2D copy with the scoring
            return eng.score_host(xs, self.var_weight, self._gamma(), index0=index0, engine=self.score_engine,
                                  reset_best=reset_best)
        key, _ = eng.score(xs, self.var_weight, self._gamma(), index0=index0, want_outputs=False,
                           engine=self.score_engine, reset_best=reset_best)
        return key

    def initial_guess(self):
        return self.blr.sample()

    def untrusted(self, x, badness=-1):
        self.constraint_ssp += badness * self.encode(x).reshape(-1, 1)

    def acquisition_func(self):
        """phi-space objective/gradient pair of ssp_agent.py:156-185, on host copies of (m, S).  This is the
        reference's L-BFGS route ('next' row f.2 of SURVEY.md section 8); `maximize` here uses `best_candidate`."""
        margin = 4
        m, sigma = self.blr.m, self.blr.S
        gamma, beta_inv, lam = self._gamma(), 1 / float(np.ravel(self.blr.beta)[0]), self.var_weight

        def _clamp(phi):
            nrm = np.linalg.norm(phi)
            return margin * phi / nrm if nrm > margin else phi

        def min_func(phi):
            phi = _clamp(phi)
            mi = lam * np.sqrt(gamma + beta_inv + phi.T @ sigma @ phi) - np.sqrt(gamma)
            return -(phi.T @ m + mi).flatten()

        def gradient(phi):
            phi = _clamp(phi)
            sig_phi = sigma @ phi
            return -(m.flatten() + lam * sig_phi / np.sqrt(phi.T @ sig_phi + gamma + beta_inv))

        return min_func, gradient

    def maximize_acquisition(self, phi_init=None, num_restarts=5, rng=None, max_iter=200, tol=1e-7, margin=4.0):
        """Device version of the reference's multi-restart phi-space search (bayesian_optimization.py:249-287 over
        `acquisition_func`): every restart climbs the same objective with the monotone fixed-point step of
        sspbo_acq_ascent.  phi_init (R, ssp_dim) are the starting points; by default `num_restarts` points drawn
        uniformly from the domain and encoded on the device (the reference draws from `blr.sample()`, an O(d^3) SVD on
        the host).  Returns (phi (R, ssp_dim), values (R,)) as numpy float64; values are -min_func(phi)."""
        eng = self._scoring_engine()
        if phi_init is None:
            rng = np.random.default_rng(rng)
            phi_init = self._encode_device(self._random_inputs(int(num_restarts), rng))
        phi, val, _ = eng.acq_ascent(phi_init, self.var_weight, self._gamma(), margin=margin, max_iter=max_iter, tol=tol)
        return phi.cpu().numpy(), val.cpu().numpy()

    def _random_inputs(self, count, rng):
        b = np.asarray(self.ssp_space.domain_bounds, dtype=np.float64)
        return rng.uniform(b[:, 0], b[:, 1], size=(count, b.shape[0]))

    def observe(self, x_t: np.ndarray, y_t: np.ndarray, step_num=0):
        """`eval(x_t)` followed by `update(x_t, y_t, var_t, step_num)` for ONE point, as `maximize` calls them back to back
        (bayesian_optimization.py:299-303): the predictive mean / variance come out of the update launch itself, so the
        pair costs one device round trip.  Returns what eval returns: mu (1, 1), var (1,), acq (1,)."""
        x_val = np.asarray(x_t, dtype=np.float64).reshape(1, -1)
        t = float(np.ravel(y_t)[0])
        lam, gamma = self.var_weight, self._gamma()
        mu, var = self._scoring_engine().observe(x_val, t)
        var_t = np.array([var])
        acq = lam * (np.sqrt(var_t + gamma) - np.sqrt(gamma))                 # ssp_agent.py:136, with the pre-update lambda, gamma
        self._step_scalars(var_t, step_num)
        return np.array([[mu]]), var_t, acq

    def update(self, x_t: np.ndarray, y_t: np.ndarray, sigma_t: float, step_num=0):
        x_val, y_val = x_t, y_t
        if len(x_t.shape) < 2:
            x_val = x_t.reshape(1, x_t.shape[0])
            y_val = y_t.reshape(1, y_t.shape[0])
        x_val = np.asarray(x_val, dtype=np.float64)
        self.blr.update_points(x_val.reshape(x_val.shape[0], -1), y_val)   # encoded on the device: phi never visits the host
        self._step_scalars(sigma_t, step_num)

    def _step_scalars(self, sigma_t, step_num):
        """lambda += var_decay ; gamma_t += gamma_c * sigma_t (ssp_agent.py:208-218)."""
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

    def encode(self, x):
        return self.ssp_space.encode(x)

    def _encode_device(self, x):
        """(k, n) -> device tensor (k, ssp_dim) float64 in this agent's input layout."""
        return self._scoring_engine().encode_device(np.atleast_2d(x))

    def decode(self, ssp):
        return self.ssp_space.decode(ssp, method=self.decoder_method, samples=self.init_samples)
