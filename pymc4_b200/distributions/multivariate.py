"""Multivariate distributions on the sampler's hot path.

``MvNormal`` mirrors /root/reference/pymc4/distributions/multivariate.py:159-199 (loc, covariance_matrix ->
``tfd.MultivariateNormalFullCovariance``).  On the device it is built for the marginal GP regression of
BASELINE configs[3]: an *observed* vector whose covariance is ``ExpQuad(l, a)(X, X) + noise * I`` with free
scalar hyper-parameters (pymc4_b200/ir.py: GP; csrc/pm4_gp.cu).  ``MvNormalCholesky`` mirrors :331-375 (loc, scale_tril ->
``tfd.MultivariateNormalTriL``).  With a CONSTANT covariance / Cholesky factor both also run as latent
(sampled) variables or as observations with a free ``loc``: the quadratic form over the constant precision
matrix is lowered to an element-wise potential over the k (k + 1) / 2 index pairs (pymc4_b200/ir.py:
_mvn_constant_term).  Host-side ``log_prob`` / ``sample`` work for any concrete covariance.
"""
import math

import numpy as np

from .distribution import ContinuousDistribution

__all__ = ["MvNormal", "MvNormalCholesky"]


class MvNormal(ContinuousDistribution):
    """Multivariate normal parameterised by ``loc`` and a full ``covariance_matrix``."""

    _kind = "mvn"
    _param_names = ("loc", "covariance_matrix")

    def __init__(self, name, loc, covariance_matrix, **kwargs):
        super().__init__(name, loc=loc, covariance_matrix=covariance_matrix, **kwargs)
        cov = self.params["covariance_matrix"]
        if cov.ndim != 2 or cov.shape[0] != cov.shape[1]:
            raise ValueError("covariance_matrix must be square [k, k], got shape {}".format(cov.shape))
        if self.params["loc"].ndim > 1 or (self.params["loc"].ndim == 1 and self.params["loc"].shape[0] != cov.shape[0]):
            raise ValueError("loc must be a scalar or a vector of length {}".format(cov.shape[0]))

    @property
    def _param_batch_shape(self):
        return ()

    @property
    def event_shape(self):
        return (self.params["covariance_matrix"].shape[0],)

    @property
    def batch_shape(self):
        from .distribution import _as_shape

        return _as_shape(self.batch_stack)

    @property
    def dist_shape(self):
        return self.batch_shape + self.event_shape

    def _log_prob_numpy(self, value, loc, covariance_matrix):
        k = covariance_matrix.shape[0]
        L = np.linalg.cholesky(covariance_matrix)
        r = np.asarray(value, dtype=np.float64) - loc
        z = np.linalg.solve(L, r.reshape(-1, k).T)  # [k, batch]
        lp = -0.5 * np.sum(np.square(z), axis=0) - np.sum(np.log(np.diag(L))) - 0.5 * k * math.log(2.0 * math.pi)
        return lp.reshape(np.shape(value)[:-1])

    def _sample_numpy(self, rng, shape, loc, covariance_matrix):
        L = np.linalg.cholesky(covariance_matrix)
        z = rng.standard_normal(shape)
        return loc + z @ L.T


class MvNormalCholesky(MvNormal):
    """Multivariate normal parameterised by ``loc`` and the lower-triangular Cholesky factor ``scale_tril`` of its
    covariance (reference multivariate.py:331-375)."""

    _kind = "mvn_chol"
    _param_names = ("loc", "scale_tril")

    def __init__(self, name, loc, scale_tril, **kwargs):
        ContinuousDistribution.__init__(self, name, loc=loc, scale_tril=scale_tril, **kwargs)
        L = self.params["scale_tril"]
        if L.ndim != 2 or L.shape[0] != L.shape[1]:
            raise ValueError("scale_tril must be square [k, k], got shape {}".format(L.shape))
        if self.params["loc"].ndim > 1 or (self.params["loc"].ndim == 1 and self.params["loc"].shape[0] != L.shape[0]):
            raise ValueError("loc must be a scalar or a vector of length {}".format(L.shape[0]))

    @property
    def event_shape(self):
        return (self.params["scale_tril"].shape[0],)

    def _log_prob_numpy(self, value, loc, scale_tril):
        L = np.tril(scale_tril)
        return MvNormal._log_prob_numpy(self, value, loc, L @ L.T)

    def _sample_numpy(self, rng, shape, loc, scale_tril):
        z = rng.standard_normal(shape)
        return loc + z @ np.tril(scale_tril).T
