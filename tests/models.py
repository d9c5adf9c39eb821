"""Models shared by the tests, bench.py and __graft_entry__ (no reference files are read at run time)."""
import os

import numpy as np

import pymc4_b200 as pm
from pymc4_b200.benchmodels import (  # noqa: F401  (re-exported: the bench workloads live in the package)
    gp_data, gp_model, hierarchical_model, logistic_glm_data, logistic_glm_model, radon_large, radon_synthetic)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# data of /root/reference/tests/test_8schools.py:7-9
J = 8
Y8 = np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype=np.float32)
SIGMA8 = np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype=np.float32)


@pm.model
def schools_pm4():
    """Eight schools, non-centred (/root/reference/tests/test_8schools.py:12-21)."""
    eta = yield pm.Normal("eta", 0, 1, batch_stack=J)
    mu = yield pm.Normal("mu", 1, 10)
    tau = yield pm.HalfNormal("tau", 1 * 2.0)
    theta = mu + tau * eta
    obs = yield pm.Normal("obs", theta, SIGMA8, observed=Y8)
    return obs


def radon_real():
    d = np.load(os.path.join(GOLDEN, "radon_919.npz"))
    return d["idx"], d["x"], d["y"], 85


def radon_model(real=False, **kw):
    idx, x, y, g = radon_real() if real else radon_synthetic(**kw)
    return hierarchical_model(idx, x, y, g)


@pm.model
def logistic_model(X, y):
    """Hierarchical logistic regression with an explicit-sum linear predictor (small P)."""
    mu = yield pm.Normal("mu", 0.0, 1.0)
    tau = yield pm.HalfNormal("tau", 1.0)
    beta = yield pm.Normal("beta", mu, tau, batch_stack=X.shape[1])
    eta = 0.0
    for k in range(X.shape[1]):
        eta = eta + beta[k] * X[:, k]
    obs = yield pm.Bernoulli("obs", pm.math.sigmoid(eta), observed=y)


def logistic_data(n=200, p=3, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    beta = rng.normal(0, 0.7, p)
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return X, y


@pm.model
def poisson_model(idx, y, n_groups):
    """Group-level log rates with a Poisson likelihood (rugby-notebook-like glue: exp of gathered sums)."""
    intercept = yield pm.Normal("intercept", 0.0, 2.0)
    sd = yield pm.HalfCauchy("sd", 1.0)
    eff = yield pm.Normal("eff", 0.0, sd, batch_stack=n_groups)
    rate = pm.math.exp(intercept + pm.math.gather(eff, idx))
    obs = yield pm.Poisson("obs", rate, observed=y)


def poisson_data(n=150, g=6, seed=5):
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, g, n).astype(np.int32)  # deliberately NOT sorted
    eff = rng.normal(0, 0.5, g)
    y = rng.poisson(np.exp(0.8 + eff[idx])).astype(np.float64)
    return idx, y, g


@pm.model
def simple_model():
    norm = yield pm.Normal("norm", 0, 1)
    return norm


@pm.model
def simple_model_with_deterministic():
    norm = yield simple_model()
    determ = yield pm.Deterministic("determ", norm * 2)
    return determ


@pm.model
def simple_model_no_free_rvs():
    norm = yield pm.Normal("norm", 0, 1, observed=1)
    return norm


@pm.model
def nested_model(cond):
    x = yield pm.Normal("x", cond, 1)
    dx = yield pm.Deterministic("dx", x + 1)
    return dx


@pm.model
def outer_model():
    cond = yield pm.HalfNormal("cond", 1)
    dcond = yield pm.Deterministic("dcond", cond * 2)
    dx = yield nested_model(dcond)
    ddx = yield pm.Deterministic("ddx", dx)
    return ddx



def mvn_latent_data(k=6, seed=0):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((k, k))
    cov = A @ A.T + np.eye(k)
    return cov, rng.standard_normal(k)


def mvn_latent_model(cov=None, y=None):
    """A latent multivariate normal vector with a constant Cholesky factor under a Normal observation model
    (reference distributions/multivariate.py:331-375 MvNormalCholesky as a sampled variable)."""
    if cov is None:
        cov, y = mvn_latent_data()
    L = np.linalg.cholesky(cov)
    k = L.shape[0]

    @pm.model
    def mvn_latent():
        f = yield pm.MvNormalCholesky("f", loc=np.zeros(k), scale_tril=L)
        s = yield pm.HalfNormal("s", 1.0)
        yield pm.Normal("y", f, s, observed=y)

    return mvn_latent()


def same_structure_models():
    """Two models of one structure and layout whose variable names differ (they share a device model)."""
    @pm.model
    def first():
        a = yield pm.Normal("a", 0.0, 1.0)
        yield pm.Deterministic("twice_a", a * 2)
        yield pm.Normal("ya", a, 1.0, observed=np.array([0.5, -0.25]))

    @pm.model
    def second():
        b = yield pm.Normal("b", 0.0, 1.0)
        yield pm.Deterministic("twice_b", b * 2)
        yield pm.Normal("yb", b, 1.0, observed=np.array([0.5, -0.25]))

    return first(), second()
