"""SSP spaces with the reference's public surface, backed by the B200 kernels.

Mirrors `ssp_bayes_opt/sspspace.py` of ctn-waterloo/ssp-bayesopt:
  SSPSpace            sspspace.py:192-590   (encode :254-276, decode :317-397,
                                             get_sample_points :452-514, bind :551-557)
  RandomSSPSpace      sspspace.py:593-654
  HexagonalSSPSpace   sspspace.py:680-785
Phase matrices are built on the host with numpy in the reference's RNG call
order, so seeded spaces are bit-identical.  `encode`, grid generation and
`decode(method='from-set')` run on the GPU through libsspbo.so; there is no
numpy implementation of them in this package.

SPSpace (sspspace.py:14-186) is re-hosted as host numpy only (tag vectors of the multi-agent encoder); its vectors
enter the device encoder as per-frequency phases.
Not re-hosted (out of the hot-path scope): train_decoder_net / 'network' decoders
(TensorFlow), Nengo encoder samplers, plotting.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np
from scipy.special import gammainc
from scipy.stats import qmc, special_ortho_group

from .engine import DeviceEngine


def _get_rng(rng=None):
    if rng is None:
        return np.random.default_rng()
    if isinstance(rng, (int, np.integer)):
        return np.random.default_rng(int(rng))
    return rng


def conjsym(K: np.ndarray, even: bool = False) -> np.ndarray:
    """[0; K; -flip(K)]: (2k+1, n) real phase matrix of a conjugate-symmetric spectrum.
    The `even` argument is accepted and ignored, as in the reference (sspspace.py:821-827)."""
    k, n = K.shape
    full = np.zeros((2 * k + 1, n), dtype=np.float64)
    full[1:k + 1] = K
    full[k + 1:] = -K[::-1]
    return full


def _rd_sequence(count: int, dim: int, seed: float = 0.5) -> np.ndarray:
    """R_d low-discrepancy sequence (sspspace.py:830-845)."""
    g = 2.0
    for _ in range(10):
        g = (1 + g) ** (1.0 / (dim + 1))
    alpha = np.mod((1.0 / g) ** np.arange(1, dim + 1), 1.0)
    return np.mod(seed + np.outer(np.arange(1, count + 1), alpha), 1.0)


class SPSpace:
    """Random unitary tag vectors for discrete symbols (sspspace.py:14-186), host numpy in the reference's RNG call order.
    Only the pieces the multi-agent trajectory agent uses: `vectors`, `inverse_vectors`, `make_unitary`, `invert`, `bind`,
    `identity`."""

    def __init__(self, domain_size: int, dim: int, vectors: Optional[np.ndarray] = None,
                 rng: Optional[Union[int, np.random.Generator]] = None, **kwargs):
        self.domain_size = int(domain_size)
        self.dim = int(dim)
        self.rng = _get_rng(rng)
        if self.domain_size == 1:                                   # a single symbol is the identity
            self.vectors = np.zeros((self.domain_size, self.dim))
            self.vectors[:, 0] = 1
        elif vectors is not None:
            assert vectors.shape[0] == self.domain_size
            assert vectors.shape[1] == self.dim
            self.vectors = vectors
        else:
            v = self.rng.normal(size=(self.domain_size, self.dim))
            v /= np.linalg.norm(v, axis=0, keepdims=True)
            self.vectors = self.make_unitary(v)
        self.inverse_vectors = self.invert(self.vectors)

    def make_unitary(self, v):
        fv = np.fft.fft(v, axis=-1)
        fv = fv / np.sqrt(fv.real ** 2 + fv.imag ** 2)
        return np.fft.ifft(fv, axis=-1).real

    def identity(self):
        s = np.zeros((1, self.dim))
        s[:, 0] = 1
        return s

    def bind(self, *arrays):
        arrays = [np.atleast_2d(a) for a in arrays]
        f = np.fft.fft(arrays[0], axis=-1)
        for a in arrays[1:]:
            f = f * np.fft.fft(a, axis=-1)
        return np.fft.ifft(f, axis=-1).real

    def invert(self, a):
        a = np.atleast_2d(a)
        return a[:, -np.arange(self.dim)]


class SSPSpace:
    """Spatial Semantic Pointer space over a `domain_dim`-dimensional domain.

    Same constructor and attributes as the reference (sspspace.py:216-241).
    """

    def __init__(self, domain_dim: int, ssp_dim: int, phase_matrix: np.ndarray,
                 domain_bounds: Optional[np.ndarray] = None,
                 length_scale: Optional[Union[float, list, np.ndarray]] = 1,
                 rng=None, device=None):
        self.domain_dim = domain_dim
        self.ssp_dim = ssp_dim
        if isinstance(length_scale, (np.ndarray, list)):
            length_scale = np.array(length_scale).reshape(self.domain_dim, 1)
        self.length_scale = np.array(length_scale) * np.ones((self.domain_dim, 1))
        self.rng = _get_rng(rng)
        if domain_bounds is not None:
            assert domain_bounds.shape[0] == domain_dim
        self.domain_bounds = domain_bounds
        self.decoder_model = None
        assert phase_matrix.shape[0] == ssp_dim
        assert phase_matrix.shape[1] == domain_dim
        self.phase_matrix = phase_matrix
        self._device = device
        self._engine: Optional[DeviceEngine] = None
        self._engine_ls = None

    # ------------------------------------------------------------------ device plumbing
    @property
    def engine(self) -> DeviceEngine:
        """Device context of this space; (re)programmed whenever the length scale changed."""
        ls = np.asarray(self.length_scale, dtype=np.float64).reshape(-1)
        if self._engine is None:
            self._engine = DeviceEngine(self._device)
            self._engine_ls = None
        if self._engine_ls is None or not np.array_equal(ls, self._engine_ls):
            self._engine.set_space(self.phase_matrix, ls)
            if self.domain_bounds is not None:           # lets the scorer use its 32-bit fixed-point phase path
                self._engine.set_xbound(float(np.max(np.abs(np.asarray(self.domain_bounds, dtype=np.float64)))))
            self._engine_ls = ls.copy()
        return self._engine

    def update_lengthscale(self, scale):
        """sspspace.py:243-252: scalar or per-axis array; stored with shape (domain_dim,)."""
        if not isinstance(scale, np.ndarray) or scale.size == 1:
            self.length_scale = scale * np.ones((self.domain_dim,))
        else:
            assert scale.size == self.domain_dim
            self.length_scale = scale
        assert self.length_scale.size == self.domain_dim

    # ------------------------------------------------------------------ encode
    def encode(self, x, dtype=None):
        """(num_samples, domain_dim) -> (num_samples, ssp_dim) float64 (sspspace.py:254-276), on the GPU.
        `dtype=np.float32` (extension) uses the tensor-core encoder: ~1e-6 accurate, ~50x faster for large batches."""
        x = np.atleast_2d(x)
        assert x.shape[1] == self.domain_dim, \
            f'Expected data with {self.domain_dim} columns, got {x.shape}'
        if dtype is not None and np.dtype(dtype) == np.float32:
            return self.engine.encode_f32_device(x).cpu().numpy()
        return self.engine.encode(x)


    def encode_and_deriv(self, x):
        """(num_samples, domain_dim) -> (ssp (num_samples, ssp_dim), grad (num_samples, ssp_dim, domain_dim)) float64
        (sspspace.py:278-306), on the GPU.  grad[s, j, a] = d phi_j(x_s) / d x_a = Re IFFT(i (A[:, a] / l_a) exp(i A x_s / l)).
        The reference's expression (:303-304) multiplies a (ssp_dim, domain_dim) by a (ssp_dim, num_samples) matrix and
        raises unless they happen to agree; this is the derivative its docstring describes."""
        x = np.atleast_2d(x)
        assert x.shape[1] == self.domain_dim, f'Expected {self.domain_dim}-d data, got {x.shape}'
        phi, grad = self.engine.encode_and_deriv_device(np.ascontiguousarray(x, dtype=np.float64))
        return phi.cpu().numpy(), grad.contiguous().cpu().numpy()
    def encode_fourier(self, x):
        """exp(i A x / l) (sspspace.py:308-315): host numpy; not on the device path (diagnostic helper)."""
        x = np.atleast_2d(x)
        scaled = x / np.asarray(self.length_scale, dtype=np.float64).reshape(1, -1)
        return np.exp(1.j * self.phase_matrix @ scaled.T).T

    # ------------------------------------------------------------------ decode
    def decode(self, ssp, method='from-set', sampling_method='grid', num_samples=300, samples=None):
        """sspspace.py:317-397.  'from-set' runs entirely on the GPU; 'direct-optim' seeds scipy's
        L-BFGS-B from the GPU from-set answer and evaluates the objective with the GPU encoder."""
        ssp = np.atleast_2d(ssp)
        if ssp.shape[0] == 0 and method in ('direct-optim', 'from-set'):
            return np.zeros((0, self.domain_dim))
        if method in ('direct-optim', 'from-set'):
            if samples is None:
                # the reference's direct-optim seeds itself from a 'length-scale' decoding set (sspspace.py:366-369)
                sm = 'length-scale' if method == 'direct-optim' else sampling_method
                sample_ssps, sample_points = self.get_sample_pts_and_ssps(method=sm, num_points_per_dim=num_samples)
            else:
                sample_ssps, sample_points = samples
                assert sample_ssps.shape[1] == ssp.shape[1], f'Expected {sample_ssps.shape} dim, got {ssp.shape}'
        if method == 'from-set':
            idx = self.engine.from_set_argmax(sample_ssps, ssp).cpu().numpy()
            return np.asarray(sample_points)[idx, :]
        if method == 'direct-optim':
            # sspspace.py:358-375: from-set start, then maximise phi(x) . unit(ssp) inside the domain bounds.  Both steps
            # run on the device (sspbo_from_set_argmax, sspbo_decode_optim); the optimiser is an analytic projected Newton
            # iteration instead of scipy's L-BFGS-B with numerical gradients, on the same objective from the same start.
            idx = self.engine.from_set_argmax(sample_ssps, ssp).cpu().numpy()
            x0 = np.asarray(sample_points, dtype=np.float64)[idx, :]
            return self.engine.decode_optim(ssp, x0, self.domain_bounds).cpu().numpy()
        if method in ('network', 'network-optim'):
            raise NotImplementedError(
                f"decoding method {method!r} needs the TensorFlow decoder net, which is outside the hot-path scope")
        raise NotImplementedError(f'Unrecognized decoding method: {method}')

    def clean_up(self, ssp, method='from-set', sampling_method='grid', num_samples=300):
        return self.encode(self.decode(ssp, method, sampling_method, num_samples))

    # ------------------------------------------------------------------ candidate / decoding sets
    def _bounds(self):
        if self.domain_bounds is None:
            return np.vstack([-10 * np.ones(self.domain_dim), 10 * np.ones(self.domain_dim)]).T
        return self.domain_bounds

    def grid_axes(self, samples_per_dim=100, method='grid'):
        """Per-axis np.linspace arrays of the 'grid' / 'length-scale' samplers (sspspace.py:480-491)."""
        bounds = self._bounds()
        if method == 'grid':
            counts = [samples_per_dim for _ in range(bounds.shape[0])]
        elif method == 'length-scale':
            ls = np.asarray(self.length_scale, dtype=np.float64).reshape(-1)
            counts = [2 * int(np.ceil((b[1] - b[0]) / ls[i])) for i, b in enumerate(bounds)]
        else:
            raise NotImplementedError(method)
        return [np.linspace(bounds[i, 0], bounds[i, 1], int(counts[i])) for i in range(self.domain_dim)]

    def get_sample_points(self, samples_per_dim=100, method='length-scale'):
        """(num_samples, domain_dim) candidate points (sspspace.py:452-514).  Grids are generated on the GPU
        in the reference's meshgrid('xy') order from host linspaces, so they are bit-identical."""
        bounds = self._bounds()
        if method in ('grid', 'length-scale'):
            axes = self.grid_axes(samples_per_dim, method)
            pts = self.engine.grid_points(axes).cpu().numpy()
            assert pts.shape[1] == self.domain_dim
            return pts
        if method == 'sobol':
            num_points = int(np.prod(samples_per_dim))
            u = qmc.Sobol(d=self.domain_dim, seed=self.rng).random(num_points)
        elif method == 'Rd':
            u = _rd_sequence(int(np.prod(samples_per_dim)), self.domain_dim)
        else:
            raise NotImplementedError(f'Sampling method {method} is not implemented')
        return qmc.scale(u, bounds[:, 0], bounds[:, 1])

    def get_sample_ssps(self, num_points, **kwargs):
        return self.encode(self.get_sample_points(num_points, **kwargs))

    def get_sample_pts_and_ssps(self, num_points_per_dim=100, method='grid'):
        pts = self.get_sample_points(method=method, samples_per_dim=num_points_per_dim)
        if method == 'grid':
            expected = int(num_points_per_dim ** self.domain_dim)
            assert pts.shape[0] == expected, f'Expected {expected} samples, got {pts.shape[0]}.'
        return self.encode(pts), pts

    # ------------------------------------------------------------------ algebra on SSP vectors (host, O(d log d))
    def normalize(self, ssp):
        return ssp / np.maximum(np.linalg.norm(ssp, axis=-1, keepdims=True), 1e-8)

    def make_unitary(self, ssp):
        f = np.fft.fft(ssp, axis=-1)
        f = f / np.maximum(np.abs(f), 1e-8)
        return np.fft.ifft(f, axis=-1).real

    def identity(self):
        s = np.zeros((1, self.ssp_dim))
        s[:, 0] = 1
        return s

    def bind(self, *arrays):
        """Circular convolution (sspspace.py:551-557), on the GPU (`sspbo_bind`, direct float64 convolution; the
        reference's fft/ifft route agrees to ~1e-15).  The trajectory encoder never calls this (it sums phasors in
        the Fourier domain); SSPTrajectoryAgent.decode uses it to unbind the L timesteps."""
        eng = self.engine
        acc = eng.to_device(np.atleast_2d(arrays[0]), eng.torch.float64)
        for a in arrays[1:]:
            acc = eng.bind(acc, np.atleast_2d(a))
        return acc.cpu().numpy()

    def invert(self, a):
        a = np.atleast_2d(a)
        return a[:, -np.arange(self.ssp_dim)]


class RandomSSPSpace(SSPSpace):
    """Random phase matrix; same signature and RNG call order as sspspace.py:608-654."""

    def __init__(self, domain_dim: int, ssp_dim: int, domain_bounds: Optional[np.ndarray] = None,
                 scale_min: Optional[float] = 1e-3, scale_max: Optional[float] = np.pi,
                 length_scale: Optional[Union[float, list, np.ndarray]] = 1,
                 sampler: Optional[str] = 'unif', norm_scale: Optional[float] = None,
                 rng=None, device=None, **kwargs):
        self.rng = _get_rng(rng)
        half = (ssp_dim - 1) // 2
        if sampler == 'unif':
            a = self.rng.uniform(0, 1, (half, domain_dim))
            sign = self.rng.choice((-1, +1), a.shape)
            phases = sign * scale_max * (scale_min + a * (1 - 2 * scale_min))
        elif sampler == 'norm':
            if norm_scale is None:
                norm_scale = np.sqrt(np.pi / 2) * ((scale_max - scale_min) / 2 + scale_min)
            phases = self.rng.normal(loc=0., scale=norm_scale, size=(half, domain_dim))
        elif sampler == 'log':
            draws = self.rng.normal(size=(half, domain_dim))
            ssq = np.sum(draws ** 2, axis=1)
            u = gammainc(domain_dim / 2, ssq / 2)
            log_r = np.log(scale_min) + (np.log(scale_max) - np.log(scale_min)) * u
            phases = draws * (np.exp(log_r) / np.sqrt(ssq))[:, None]
        else:
            raise NotImplementedError()
        phase_matrix = conjsym(phases, ssp_dim % 2 == 0)
        super().__init__(domain_dim, phase_matrix.shape[0], phase_matrix=phase_matrix,
                         domain_bounds=domain_bounds, length_scale=length_scale, rng=self.rng, device=device)


class HexagonalSSPSpace(SSPSpace):
    """Simplex x scales x rotations phase matrix; same signature as sspspace.py:713-785."""

    def __init__(self, domain_dim: int, ssp_dim: Optional[int] = 151, domain_bounds: Optional[np.ndarray] = None,
                 n_scales: Optional[int] = -1, n_rotates: Optional[int] = -1,
                 scale_min: Optional[float] = 0.1, scale_max: Optional[float] = np.pi,
                 length_scale: Optional[Union[float, list, np.ndarray]] = 1,
                 scale_sampling: Optional[str] = 'lin', rng=None, rand_pad: Optional[bool] = False,
                 device=None, **kwargs):
        assert ssp_dim >= 2 * (domain_dim + 1) + 1
        self.rng = _get_rng(rng)
        n = domain_dim
        if (n_rotates <= 0) or (n_scales <= 0):
            n_rotates = int(np.sqrt((ssp_dim - 1) / (2 * (n + 1))))
            n_scales = n_rotates
        # n+1 vertices of the regular n-simplex, as rows
        simplex = np.hstack([np.sqrt(1 + 1 / n) * np.identity(n) - (n ** (-3 / 2)) * (np.sqrt(n + 1) + 1),
                             (n ** (-1 / 2)) * np.ones((n, 1))]).T
        self.grid_basis_dim = n + 1
        self.num_grids = n_rotates * n_scales
        self.scale_min, self.scale_max = scale_min, scale_max
        self.n_scales, self.n_rotates = n_scales, n_rotates

        golden = (1 + np.sqrt(5)) / 2
        if n == 1:
            n_scales = n_scales * n_rotates
        if scale_sampling == 'lin':
            if scale_min is None:
                scale_min = scale_max / (n_scales * (golden - 1) + 1)
            scales = np.linspace(scale_min, scale_max, n_scales)
        elif scale_sampling == 'log':
            if scale_min is None:
                scale_min = scale_max / (golden ** (n_scales - 1))
            scales = np.geomspace(scale_min, scale_max, n_scales)
        elif scale_sampling == 'rand':
            if scale_min is None:
                scale_min = 0
            scales = self.rng.uniform(scale_min, scale_max, n_scales)
        else:
            raise NotImplementedError(scale_sampling)
        scaled = np.vstack([simplex * s for s in scales])

        if (n_rotates == 1) or (n == 1):
            rows = scaled
        else:
            if n == 2:
                ang = np.linspace(0, 2 * np.pi / 3, n_rotates, endpoint=False)
                rot = np.stack([np.stack([np.cos(ang), -np.sin(ang)], axis=1),
                                np.stack([np.sin(ang), np.cos(ang)], axis=1)], axis=1)
            else:
                rot = special_ortho_group.rvs(n, size=n_rotates, random_state=self.rng)
            rows = (rot @ scaled.T).transpose(0, 2, 1).reshape(-1, n)

        actual = 2 * rows.shape[0] + 1
        if rand_pad and (actual < ssp_dim):
            pad = RandomSSPSpace(n, ssp_dim - actual).phase_matrix[1:1 + (ssp_dim - actual) // 2]
            rows = np.vstack([rows, pad])
        phase_matrix = conjsym(rows)
        super().__init__(n, phase_matrix.shape[0], phase_matrix=phase_matrix, domain_bounds=domain_bounds,
                         length_scale=length_scale, rng=self.rng, device=device)
