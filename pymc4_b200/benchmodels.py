"""Models and synthetic data of the BASELINE.json configurations (bench.py, __graft_entry__, tools/, tests).

No reference files are read at run time.  Shapes: SURVEY section 8(d).
"""
import numpy as np

import pymc4_b200 as pm


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


@pm.model
def hierarchical_model_noncentered(idx, x, y, n_groups):
    """The same radon model with the group effects written as offsets, alpha = mu_alpha + sigma_alpha * alpha_raw
    (the parameterisation that removes the funnel between the group effects and their scale: with it the reference-default
    sampler settings converge -- split R-hat < 1.01, no divergent transitions -- which the centred spelling of the
    reference's notebook does not within 1000 draws)."""
    mu_a = yield pm.Normal(loc=0.0, scale=1, name="mu_alpha")
    sigma_a = yield pm.HalfCauchy(scale=1, name="sigma_alpha")
    mu_b = yield pm.Normal(loc=0.0, scale=1, name="mu_beta")
    sigma_b = yield pm.HalfCauchy(scale=1, name="sigma_beta")
    a_raw = yield pm.Normal(loc=0.0, scale=1.0, batch_stack=n_groups, name="alpha_raw")
    b_raw = yield pm.Normal(loc=0.0, scale=1.0, batch_stack=n_groups, name="beta_raw")
    eps = yield pm.HalfCauchy(scale=1, name="eps")
    a = mu_a + sigma_a * a_raw
    b = mu_b + sigma_b * b_raw
    radon_est = pm.math.gather(a, idx) + pm.math.gather(b, idx) * x
    y_like = yield pm.Normal(loc=radon_est, scale=eps, observed=y, name="y_like")
