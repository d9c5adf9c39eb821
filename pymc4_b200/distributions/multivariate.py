"""Multivariate distributions on the sampler's hot path.

``MvNormal`` mirrors /root/reference/pymc4/distributions/multivariate.py:159-199 (loc, covariance_matrix ->
``tfd.MultivariateNormalFullCovariance``).  On the device it is built for the marginal GP regression of
BASELINE configs[3]: an *observed* vector whose covariance is ``ExpQuad(l, a)(X, X) + noise * I`` with free
scalar hyper-parameters (pymc4_b200/ir.py: GP; csrc/pm4_gp.cu).  Host-side ``log_prob`` / ``sample`` work for
any concrete covariance.
"""
import math

import numpy as np

from .distribution import ContinuousDistribution

__all__ = ["MvNormal"]


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
