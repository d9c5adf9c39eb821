"""Models shared by the tests, bench.py and __graft_entry__ (no reference files are read at run time)."""
import os

import numpy as np

import pymc4_b200 as pm

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


@pm.model
def hierarchical_model(idx, x, y, n_groups):
    """Radon varying-intercept / varying-slope model
    (/root/reference/notebooks/radon_hierarchical.ipynb:72-95)."""
    mu_a = yield pm.Normal(loc=0.0, scale=1, name="mu_alpha")
    sigma_a = yield pm.HalfCauchy(scale=1, name="sigma_alpha")
    mu_b = yield pm.Normal(loc=0.0, scale=1, name="mu_beta")
    sigma_b = yield pm.HalfCauchy(scale=1, name="sigma_beta")
    a = yield pm.Normal(loc=mu_a, scale=sigma_a, batch_stack=n_groups, name="alpha")
    b = yield pm.Normal(loc=mu_b, scale=sigma_b, batch_stack=n_groups, name="beta")
    eps = yield pm.HalfCauchy(scale=1, name="eps")
    radon_est = pm.math.gather(a, idx) + pm.math.gather(b, idx) * x
    y_like = yield pm.Normal(loc=radon_est, scale=eps, observed=y, name="y_like")


def radon_real():
    d = np.load(os.path.join(GOLDEN, "radon_919.npz"))
    return d["idx"], d["x"], d["y"], 85


def radon_synthetic(n_obs=919, n_groups=85, seed=20261019):
    """Synthetic radon-shaped data: sorted group index, x ~ Bernoulli(0.17),
    y = a[g] + b[g] x + eps N(0,1) with the notebook's posterior means as truth."""
    rng = np.random.default_rng(seed)
    # skewed group sizes: every group at least one observation
    w = rng.lognormal(0.0, 1.0, n_groups)
    sizes = 1 + rng.multinomial(n_obs - n_groups, w / w.sum())
    idx = np.repeat(np.arange(n_groups, dtype=np.int32), sizes)
    x = (rng.random(n_obs) < 0.17).astype(np.float64)
    a = rng.normal(1.5, 0.33, n_groups)
    b = rng.normal(-0.65, 0.35, n_groups)
    y = a[idx] + b[idx] * x + 0.72 * rng.standard_normal(n_obs)
    return idx, x, y, n_groups


def radon_large(n_obs=100_000, n_groups=5_000, seed=20261019):
    """Radon-shaped data scaled up: group sizes ~ Poisson(20) + 1, renormalised
    to n_obs observations; same generating parameters as radon_synthetic."""
    rng = np.random.default_rng(seed)
    sizes = rng.poisson(20.0, n_groups) + 1
    sizes = np.maximum(1, np.floor(sizes * (n_obs / sizes.sum())).astype(np.int64))
    sizes[: n_obs - sizes.sum()] += 1
    idx = np.repeat(np.arange(n_groups, dtype=np.int32), sizes)
    assert idx.size == n_obs
    x = (rng.random(n_obs) < 0.17).astype(np.float64)
    a = rng.normal(1.5, 0.33, n_groups)
    b = rng.normal(-0.65, 0.35, n_groups)
    y = a[idx] + b[idx] * x + 0.72 * rng.standard_normal(n_obs)
    return idx, x, y, n_groups


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


@pm.model
def logistic_glm_model(X, y):
    """Hierarchical logistic regression with a dense linear predictor X @ beta (BASELINE configs[2] shape;
    SURVEY section 8(d) item 3): beta_p ~ N(mu, tau), mu ~ N(0, 1), tau ~ HalfNormal(1), y ~ Bernoulli(sigmoid(X beta))."""
    mu = yield pm.Normal("mu", 0.0, 1.0)
    tau = yield pm.HalfNormal("tau", 1.0)
    beta = yield pm.Normal("beta", mu, tau, batch_stack=X.shape[1])
    obs = yield pm.Bernoulli("obs", pm.math.sigmoid(X @ beta), observed=y)


def logistic_glm_data(n=3000, p=20, seed=11):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    beta = rng.normal(0, 0.5, p)
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return X, y


def gp_data(n=64, d=1, seed=20261019, ls=1.0, amp=1.0, sigma=0.1):
    """BASELINE configs[3] shape: X ~ U(0, 10), y ~ MVN(0, ExpQuad(ls, amp)(X, X) + sigma^2 I)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0.0, 10.0, (n, d))
    d2 = np.sum(np.square(X[:, None, :] - X[None, :, :]), axis=-1)
    K = amp ** 2 * np.exp(-d2 / (2 * ls ** 2)) + sigma ** 2 * np.eye(n)
    y = np.linalg.cholesky(K) @ rng.standard_normal(n)
    return X, y


def gp_model(X, y):
    """Marginal GP regression spelled the way the reference allows (no marginal_likelihood, gp/gp.py:51-60)."""
    n = X.shape[0]

    @pm.model
    def gp_regression():
        ls = yield pm.HalfCauchy("ls", 5.0)
        amp = yield pm.HalfCauchy("amp", 5.0)
        sigma = yield pm.HalfNormal("sigma", 1.0)
        cov = pm.gp.cov.ExpQuad(length_scale=ls, amplitude=amp)(X, X) + sigma ** 2 * np.eye(n)
        yield pm.MvNormal("y", loc=0.0, covariance_matrix=pm.gp.util.stabilize(cov, 1e-6), observed=y)

    return gp_regression()
