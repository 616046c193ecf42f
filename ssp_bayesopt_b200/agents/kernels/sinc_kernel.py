"""Product-of-sincs covariance for scikit-learn's GaussianProcessRegressor: k(x, y) = prod_a sinc((x_a - y_a) / l_a),
the kernel an SSP space with uniformly distributed phases induces.  Mirrors `ssp_bayes_opt/agents/kernels/sinc_kernel.py`
of the reference; host-side set-up only (a GP on the handful of initial points picks the length scale when the caller
gives none, agents/ssp_agent.py:96-109) -- nothing here touches the device.

The hyper-parameter gradient is the reference's: d k / d l_a scaled as there (sinc'(u_a) * (-u_a / l_a^2) times the other
factors, u = (x - y) / l), because scikit-learn's optimiser path -- and therefore the selected length scale -- depends on
it; the fixture tests/golden/lengthscale_gp.npz pins both the kernel values and that gradient.
"""
import numpy as np
from sklearn.gaussian_process.kernels import (Hyperparameter, Kernel, NormalizedKernelMixin, StationaryKernelMixin,
                                              _check_length_scale)


def _sinc_prime(u):
    """d/du of np.sinc(u) = sin(pi u) / (pi u); the removable point u = 0 is nudged off zero as the reference does."""
    u = np.where(np.asanyarray(u) == 0, 1.0e-20, u)
    return np.cos(np.pi * u) / u - np.sin(np.pi * u) / (np.pi * u ** 2)


class SincKernel(StationaryKernelMixin, NormalizedKernelMixin, Kernel):
    def __init__(self, length_scale=1.0, length_scale_bounds=(1e-5, 1e5)):
        self.length_scale = length_scale
        self.length_scale_bounds = length_scale_bounds

    @property
    def anisotropic(self):
        return np.iterable(self.length_scale) and len(self.length_scale) > 1

    @property
    def hyperparameter_length_scale(self):
        if self.anisotropic:
            return Hyperparameter('length_scale', 'numeric', self.length_scale_bounds, len(self.length_scale))
        return Hyperparameter('length_scale', 'numeric', self.length_scale_bounds)

    def __call__(self, X, Y=None, eval_gradient=False):
        X = np.atleast_2d(X)
        ls = _check_length_scale(X, self.length_scale)
        if Y is not None and eval_gradient:
            raise ValueError('Gradient can only be evaluated when Y is None')
        other = X if Y is None else np.atleast_2d(Y)
        u = (X[:, None, :] - other[None, :, :]) / ls                       # (n, m, features)
        factors = np.sinc(u)
        K = np.prod(factors, axis=-1)
        if not eval_gradient:
            return K
        if self.hyperparameter_length_scale.fixed:
            return K, np.empty((X.shape[0], X.shape[0], 0))
        n_feat = X.shape[1]
        ls_vec = np.broadcast_to(np.asarray(ls, dtype=np.float64), (n_feat,))
        # term a: derivative of factor a times the product of the other factors
        terms = np.empty(u.shape)
        for a in range(n_feat):
            others = np.prod(np.delete(factors, a, axis=-1), axis=-1)
            terms[:, :, a] = others * _sinc_prime(u[:, :, a]) * (-u[:, :, a] / ls_vec[a] ** 2)
        if self.anisotropic:
            return K, terms
        return K, terms.sum(axis=-1, keepdims=True)

    def __repr__(self):
        if self.anisotropic:
            return '{0}(length_scale=[{1}])'.format(self.__class__.__name__, ', '.join(map('{0:.3g}'.format, self.length_scale)))
        return '{0}(length_scale={1:.3g})'.format(self.__class__.__name__, np.ravel(self.length_scale)[0])
