"""GPU tests of the drop-in ``pm.sample`` API; they read like the reference's own tests
(tests/test_8schools.py:64-77, tests/test_sampling.py:8-20,64-123, tests/test_compound.py:124-146)."""
import numpy as np
import pytest

import pymc4_b200 as pm
from oracle import pm4_oracle
from pymc4_b200 import ir as I
from tests import models as M

pytestmark = pytest.mark.gpu


def test_model_logp_golden():
    """tests/test_8schools.py:24-61 with the same call."""
    logp, *_ = pm.mcmc.samplers.build_logp_and_deterministic_functions(
        M.schools_pm4(), observed={"obs": M.Y8}, state=None)
    init_value = logp(**{"schools_pm4/eta": np.zeros(8, np.float32), "schools_pm4/mu": np.zeros((), np.float32),
                         "schools_pm4/__log_tau": np.ones((), np.float32)}).cpu().numpy()
    np.testing.assert_allclose(init_value, -42.876_114, rtol=1e-6)
    # leading batch axes are vectorised like the reference's vectorize_logp_function
    batch = logp(np.zeros((5, 3, 8), np.float32), np.zeros((5, 3), np.float32), np.ones((5, 3), np.float32))
    assert tuple(batch.shape) == (5, 3) and np.allclose(batch.cpu().numpy(), init_value)


def test_sample_eight_schools_keys_and_shapes():
    chains, samples = 4, 100
    trace = pm.inference.sampling.sample(M.schools_pm4(), step_size=0.28, num_chains=chains, num_samples=samples,
                                         burn_in=50, xla=False, seed=3)
    post = trace.posterior
    for var_name in ("eta", "mu", "tau", "__log_tau"):
        assert f"schools_pm4/{var_name}" in post.keys()
    assert "schools_pm4" not in post  # keys without "/" are dropped
    assert post["schools_pm4/eta"].shape == (chains, samples, 8) and post["schools_pm4/mu"].shape == (chains, samples)
    np.testing.assert_allclose(post["schools_pm4/tau"], np.exp(post["schools_pm4/__log_tau"]), rtol=1e-6)
    ss = trace.sample_stats
    assert set(ss) == {"lp", "tree_size", "diverging", "energy", "mean_tree_accept"}
    assert ss["tree_size"].shape == (chains, samples) and ss["tree_size"].dtype == np.int32 and ss["tree_size"].min() >= 1
    assert ss["diverging"].dtype == np.bool_ and np.all(ss["mean_tree_accept"] <= 1e-6)
    assert np.array_equal(trace.observed_data["schools_pm4/obs"], M.Y8)


@pytest.mark.parametrize("xla", [True, False])
def test_sample_deterministics(xla):
    trace = pm.sample(model=M.simple_model_with_deterministic(), num_samples=10, num_chains=4, burn_in=100,
                      step_size=0.1, xla=xla)
    norm = "simple_model_with_deterministic/simple_model/norm"
    determ = "simple_model_with_deterministic/determ"
    np.testing.assert_allclose(trace.posterior[determ], trace.posterior[norm] * 2)


def test_sampling_with_deterministics_in_nested_models():
    trace = pm.sample(model=M.outer_model(), num_samples=10, num_chains=4, burn_in=100, step_size=0.1)
    mapping = {
        "outer_model/dcond": (["outer_model/__log_cond"], lambda x: np.exp(x) * 2),
        "outer_model/ddx": (["outer_model/nested_model/dx"], lambda x: x),
        "outer_model/nested_model/dx": (["outer_model/nested_model/x"], lambda x: x + 1),
    }
    for det, (inputs, op) in mapping.items():
        np.testing.assert_allclose(trace.posterior[det], op(*[trace.posterior[i] for i in inputs]), rtol=1e-6)


def test_sampling_with_no_free_rvs():
    with pytest.raises(ValueError):
        pm.sample(model=M.simple_model_no_free_rvs(), num_samples=1, num_chains=1, burn_in=1)


def test_sample_shapes_unvectorized_model():
    observed = np.zeros((5, 4), dtype="float32")

    @pm.model
    def model():
        mu = yield pm.Normal("mu", np.zeros(4), 1)
        scale = yield pm.HalfNormal("scale", 1)
        x = yield pm.Normal("x", mu, scale, batch_stack=5, observed=observed)

    trace = pm.sample(model=model(), num_samples=10, num_chains=4, burn_in=1, step_size=0.1)
    assert trace.posterior["model/mu"].shape == (4, 10, 4) and trace.posterior["model/__log_scale"].shape == (4, 10)


@pytest.mark.parametrize("sampler", ["nuts", "hmc", "nuts_simple", "hmc_simple", "rwm"])
def test_samplers_on_standard_normal_and_seed_determinism(sampler):
    """tests/test_compound.py:124-146: |posterior mean| < 0.1; the same seed gives the same trace."""
    kw = dict(sampler_type=sampler, num_chains=50, num_samples=300, burn_in=200, seed=5)
    if sampler.startswith("hmc"):
        kw.update(step_size=0.5, num_leapfrog_steps=5)
    if sampler == "rwm":
        kw.update(num_chains=200, num_samples=500)  # random walk: high autocorrelation
    t1 = pm.sample(M.simple_model(), **kw)
    t2 = pm.sample(M.simple_model(), **kw)
    x1, x2 = t1.posterior["simple_model/norm"], t2.posterior["simple_model/norm"]
    assert abs(x1.mean()) < 0.1 and abs(x1.std() - 1) < 0.1
    assert np.array_equal(x1, x2)
    t3 = pm.sample(M.simple_model(), **dict(kw, seed=6))
    assert not np.array_equal(x1, t3.posterior["simple_model/norm"])
    assert ("mean_accept" if sampler == "rwm" else "mean_tree_accept") in t1.sample_stats


def test_per_part_step_sizes_like_the_radon_notebook():
    """notebooks/radon_hierarchical.ipynb:123-146: step_size as a list of [C, *core] arrays."""
    C = 16
    model = M.radon_model(n_obs=200, n_groups=12)
    trace = pm.sample(model, num_chains=C, num_samples=20, burn_in=20, step_size=0.05, seed=1)
    keys = [k for k in trace.posterior if k.split("/")[-1] in ("mu_alpha", "mu_beta", "alpha", "beta", "__log_sigma_alpha",
                                                              "__log_sigma_beta", "__log_eps")]
    order = ["mu_alpha", "mu_beta", "alpha", "beta", "__log_sigma_alpha", "__log_sigma_beta", "__log_eps"]
    parts = []
    for name in order:
        x = trace.posterior["hierarchical_model/" + name]
        std = x.std(axis=(0, 1)) + 1e-3
        parts.append(np.broadcast_to(std, (C,) + std.shape).astype(np.float32))
    trace2 = pm.sample(model, num_chains=C, num_samples=30, burn_in=30, step_size=parts, seed=2)
    assert trace2.posterior["hierarchical_model/alpha"].shape == (C, 30, 12)
    assert np.all(np.isfinite(trace2.posterior["hierarchical_model/alpha"]))


def test_float64_mode_and_chain_offset():
    a = pm.sample(M.schools_pm4(), num_chains=4, num_samples=20, burn_in=20, seed=9, dtype="float64")
    assert a.posterior["schools_pm4/mu"].dtype == np.float64
    # chains 2..3 of a 4-chain run == a 2-chain run with chain_offset=2 under per-chain step sizes
    step = np.full((4, 1), 0.2)
    full = pm.sample(M.schools_pm4(), num_chains=4, num_samples=15, burn_in=15, seed=9, step_size=step)
    part = pm.sample(M.schools_pm4(), num_chains=2, num_samples=15, burn_in=15, seed=9, step_size=step[:2], chain_offset=2)
    np.testing.assert_array_equal(full.posterior["schools_pm4/eta"][2:], part.posterior["schools_pm4/eta"])


def test_include_log_likelihood():
    """pm.sample(..., include_log_likelihood=True): a log_likelihood group with one [chain, draw, *obs] array per
    observed variable (reference sampling.py:52-184 -> samplers.py:832-854, mcmc/utils.py:95-144)."""
    from scipy import stats

    chains, samples = 3, 20
    trace = pm.sample(M.schools_pm4(), step_size=0.28, num_chains=chains, num_samples=samples, burn_in=20,
                      seed=5, include_log_likelihood=True)
    assert "log_likelihood" in trace.groups()
    ll = trace.log_likelihood["schools_pm4/obs"]
    assert ll.shape == (chains, samples, 8)
    post = trace.posterior
    theta = post["schools_pm4/mu"][..., None] + post["schools_pm4/tau"][..., None] * post["schools_pm4/eta"]
    np.testing.assert_allclose(ll, stats.norm(theta, M.SIGMA8).logpdf(M.Y8), rtol=2e-4, atol=2e-4)
    # the free function takes the posterior in sampler layout [draw, chain, ...]
    logp, init, _, _, state = pm.mcmc.samplers.build_logp_and_deterministic_functions(M.schools_pm4())
    posterior = {k: np.swapaxes(post[k], 0, 1) for k in init}
    ll2 = pm.mcmc.samplers.calculate_log_likelihood(M.schools_pm4(), posterior, state)
    np.testing.assert_allclose(np.swapaxes(ll2["schools_pm4/obs"].cpu().numpy(), 0, 1), ll, rtol=1e-5, atol=1e-5)


def test_sample_posterior_predictive():
    """pm.sample_posterior_predictive(model, trace): a posterior_predictive group with one
    [chain, draw, *observed shape] array per observed variable (reference forward_sampling.py:230-427)."""
    chains, samples = 4, 200
    trace = pm.sample(M.schools_pm4(), step_size=0.28, num_chains=chains, num_samples=samples, burn_in=100, seed=7)
    out = pm.sample_posterior_predictive(M.schools_pm4(), trace, seed=1)
    assert out is not trace and "posterior_predictive" in out.groups() and "posterior" in out.groups()
    ppc = out.posterior_predictive
    assert sorted(ppc) == ["schools_pm4/obs"] and ppc["schools_pm4/obs"].shape == (chains, samples, 8)
    post = trace.posterior
    theta = post["schools_pm4/mu"][..., None] + post["schools_pm4/tau"][..., None] * post["schools_pm4/eta"]
    z = (ppc["schools_pm4/obs"] - theta) / M.SIGMA8
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05
    again = pm.sample_posterior_predictive(M.schools_pm4(), trace, seed=1, inplace=False)
    assert again.groups() == ["posterior_predictive"]
    assert np.array_equal(again.posterior_predictive["schools_pm4/obs"], ppc["schools_pm4/obs"])
    with pytest.raises(KeyError):
        pm.sample_posterior_predictive(M.schools_pm4(), trace, var_names=["schools_pm4/nope"])
    with pytest.raises(ValueError):
        pm.sample_posterior_predictive(M.schools_pm4(), trace, var_names=[])
    # var_names may name variables that are not re-drawn: sampled variables come back with their posterior values,
    # back-transformed variables / deterministics are recomputed from the draws (reference forward_sampling.py:318-323)
    mixed = pm.sample_posterior_predictive(M.schools_pm4(), trace, seed=1, inplace=False,
                                           var_names=["schools_pm4/obs", "schools_pm4/mu", "schools_pm4/tau"])
    pp = mixed.posterior_predictive
    assert sorted(pp) == ["schools_pm4/mu", "schools_pm4/obs", "schools_pm4/tau"]
    np.testing.assert_array_equal(pp["schools_pm4/mu"], post["schools_pm4/mu"])
    np.testing.assert_allclose(pp["schools_pm4/tau"], np.exp(post["schools_pm4/__log_tau"]), rtol=1e-5)
    assert np.array_equal(pp["schools_pm4/obs"], ppc["schools_pm4/obs"])


def test_sample_prior_predictive():
    """pm.sample_prior_predictive(model, sample_shape): a prior_predictive group [1, *sample_shape, *core] keyed by
    the untransformed names, deterministics included (reference forward_sampling.py:26-227 and its doctest)."""
    out = pm.sample_prior_predictive(M.schools_pm4(), sample_shape=(50, 40), seed=3)
    pp = out.prior_predictive
    assert out.groups() == ["prior_predictive"]
    assert sorted(pp) == ["schools_pm4/eta", "schools_pm4/mu", "schools_pm4/obs", "schools_pm4/tau"]
    assert pp["schools_pm4/eta"].shape == (1, 50, 40, 8) and pp["schools_pm4/tau"].shape == (1, 50, 40)
    assert pp["schools_pm4/tau"].min() > 0 and abs(pp["schools_pm4/eta"].std() - 1) < 0.05
    z = (pp["schools_pm4/obs"] - (pp["schools_pm4/mu"][..., None] + pp["schools_pm4/tau"][..., None] * pp["schools_pm4/eta"])) / M.SIGMA8
    assert abs(z.std() - 1) < 0.05
    again = pm.sample_prior_predictive(M.schools_pm4(), sample_shape=(50, 40), seed=3).prior_predictive
    assert all(np.array_equal(again[k], pp[k]) for k in pp)
    only = pm.sample_prior_predictive(M.schools_pm4(), sample_shape=7, sample_from_observed=False).prior_predictive
    assert sorted(only) == ["schools_pm4/eta", "schools_pm4/mu", "schools_pm4/tau"] and only["schools_pm4/mu"].shape == (1, 7)
    det = pm.sample_prior_predictive(M.outer_model(), sample_shape=100, seed=1).prior_predictive
    assert {"outer_model/cond", "outer_model/dcond", "outer_model/nested_model/x", "outer_model/nested_model/dx",
            "outer_model/ddx"} <= set(det)
    np.testing.assert_allclose(det["outer_model/dcond"], 2 * det["outer_model/cond"], rtol=1e-5)
    np.testing.assert_allclose(det["outer_model/nested_model/dx"], det["outer_model/nested_model/x"] + 1, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        pm.sample_prior_predictive(M.schools_pm4(), var_names=["schools_pm4/nope"])


def test_sample_dense_logistic_regression_default_nuts():
    """pm.sample with the default sampler (NUTS) on the dense logistic regression: lock-step trees, every leapfrog
    one batched tensor-core gradient."""
    X, y = M.logistic_glm_data(n=3000, p=20)
    rng = np.random.default_rng(11)
    rng.standard_normal((3000, 20))
    beta_true = rng.normal(0, 0.5, 20)
    trace = pm.sample(M.logistic_glm_model(X, y), num_chains=64, num_samples=100, burn_in=100, step_size=0.02, seed=3)
    beta = trace.posterior["logistic_glm_model/beta"]
    assert beta.shape == (64, 100, 20)
    err = np.abs(beta.mean((0, 1)) - beta_true)
    assert np.max(err) < 0.25 and np.mean(err) < 0.1
    assert trace.sample_stats["tree_size"].min() >= 1 and trace.sample_stats["diverging"].mean() < 0.05


def test_sample_dense_logistic_regression_with_hmc():
    """pm.sample(..., sampler_type="hmc") on the hierarchical logistic regression with X @ beta: the posterior mean
    of beta recovers the generating coefficients."""
    X, y = M.logistic_glm_data(n=3000, p=20)
    rng = np.random.default_rng(11)
    rng.standard_normal((3000, 20))
    beta_true = rng.normal(0, 0.5, 20)  # the generator of logistic_glm_data draws X first, then beta
    trace = pm.sample(M.logistic_glm_model(X, y), sampler_type="hmc", num_chains=64, num_samples=150, burn_in=150,
                      step_size=0.02, num_leapfrog_steps=10, seed=2)
    beta = trace.posterior["logistic_glm_model/beta"]
    assert beta.shape == (64, 150, 20)
    err = np.abs(beta.mean((0, 1)) - beta_true)
    assert np.max(err) < 0.25 and np.mean(err) < 0.1
    assert 0.5 < np.exp(np.minimum(trace.sample_stats["mean_tree_accept"], 0)).mean() <= 1.0


def test_sample_gp_regression_recovers_hyperparameters():
    """BASELINE configs[3] at test size: marginal GP regression spelled with ExpQuad + MvNormal (the reference has no
    marginal_likelihood), sampled with the default NUTS; the posterior covers the generating (l, a, sigma) = (1, 1, 0.1)."""
    X, y = M.gp_data(n=128, d=1)
    trace = pm.sample(M.gp_model(X, y), num_chains=32, num_samples=100, burn_in=150, step_size=0.05, seed=5)
    post = trace.posterior
    ls, amp, sg = (post["gp_regression/" + k] for k in ("ls", "amp", "sigma"))
    assert ls.shape == (32, 100)
    assert 0.6 < np.median(ls) < 1.6 and 0.4 < np.median(amp) < 2.5 and 0.07 < np.median(sg) < 0.14
    assert trace.sample_stats["diverging"].mean() < 0.1
    # forward sampling: the observed vector is redrawn from MvNormal(0, K(theta)) for every posterior draw
    ppc = pm.sample_posterior_predictive(M.gp_model(X, y), trace, seed=1).posterior_predictive["gp_regression/y"]
    assert ppc.shape == (32, 100, 128) and np.all(np.isfinite(ppc))
    assert 0.3 < ppc.var() < 4.0 and abs(ppc.mean()) < 0.3
    prior = pm.sample_prior_predictive(M.gp_model(X, y), sample_shape=50, seed=2).prior_predictive
    assert prior["gp_regression/y"].shape == (1, 50, 128) and prior["gp_regression/ls"].shape == (1, 50)


def test_refit_on_data_of_another_length_and_on_other_indices():
    """One model fitted on data sets of different length (a cross-validation fold), then on same-shaped data with
    other group indices: every fit gets its own row layout / plans -- no stale module, no half-updated device model
    (struct hash without the row layout; set_data committing before validation)."""
    from scipy import stats

    def fit(n, seed):
        y = np.random.default_rng(seed).normal(0.5, 1.0, n)

        @pm.model
        def m():
            mu = yield pm.Normal("mu", 0.0, 1.0)
            s = yield pm.HalfNormal("s", 1.0)
            yield pm.Normal("y", mu, s, observed=y)

        tr = pm.sample(m(), step_size=0.2, num_chains=4, num_samples=30, burn_in=30, seed=3, include_log_likelihood=True)
        ll = tr.log_likelihood["m/y"]
        assert ll.shape == (4, 30, n)
        mu, s = tr.posterior["m/mu"][..., None], tr.posterior["m/s"][..., None]
        np.testing.assert_allclose(ll, stats.norm(mu, s).logpdf(y), rtol=2e-4, atol=2e-4)

    fit(100, 0)
    fit(200, 1)
    fit(100, 2)  # same layout as the first fit, new values: the cached device model takes them through set_data

    rng = np.random.default_rng(0)
    n, g = 400, 12
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    for idx in (np.sort(rng.integers(0, g, n)), np.sort(np.minimum(rng.geometric(0.4, n) - 1, g - 1))):
        model = M.hierarchical_model(idx.astype(np.int32), x, y, g)
        ir, _, _ = I.lower_model(model)
        dm = pm.mcmc.samplers._get_device_model(ir, np.float32, 0)
        th = rng.normal(0, 0.3, (16, ir.D))
        lp, gr = dm.logp_grad(th.astype(np.float32), dtype=np.float32)
        lp_ref, g_ref = pm4_oracle.Oracle(ir).logp_grad(th, dtype=np.float64)
        np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=2e-5)
        assert np.max(np.abs(gr.cpu().numpy() - g_ref)) / np.max(np.abs(g_ref)) < 2e-5


def test_streamed_trace_equals_the_plain_trace():
    """pm4_model_set_trace_sink: the draws copied to the host in slabs while the sampler keeps running are the draws
    the device holds at the end (same seed, streamed vs fetched afterwards)."""
    model = M.radon_model(real=False)
    kw = dict(step_size=0.05, num_chains=96, num_samples=40, burn_in=30, seed=11)
    plain = pm.sample(model, trace_slabs=0, **kw)
    streamed = pm.sample(model, trace_slabs=5, **kw)
    for k, v in plain.posterior.items():
        np.testing.assert_array_equal(streamed.posterior[k], v)
    for k, v in plain.sample_stats.items():
        np.testing.assert_array_equal(streamed.sample_stats[k], v)


def test_sample_latent_mvnormal_cholesky():
    """pm.sample on a model with a latent MvNormalCholesky vector (constant factor): the posterior of f given Normal
    observations with the noise scale fixed is Gaussian in closed form."""
    cov, y = M.mvn_latent_data()
    k = cov.shape[0]
    L = np.linalg.cholesky(cov)
    s = 0.7

    @pm.model
    def m():
        f = yield pm.MvNormalCholesky("f", loc=np.zeros(k), scale_tril=L)
        yield pm.Normal("y", f, s, observed=y)

    tr = pm.sample(m(), step_size=0.3, num_chains=64, num_samples=400, burn_in=200, seed=4)
    f = tr.posterior["m/f"].reshape(-1, k)
    post_cov = np.linalg.inv(np.linalg.inv(cov) + np.eye(k) / s ** 2)
    post_mean = post_cov @ (y / s ** 2)
    np.testing.assert_allclose(f.mean(axis=0), post_mean, atol=0.05)
    np.testing.assert_allclose(np.cov(f.T), post_cov, atol=0.05)


@pytest.mark.parametrize("sampler_type", ["nuts", "hmc", "rwm"])
def test_thinning_keeps_every_kth_state(sampler_type):
    """tfp sample_chain(num_steps_between_results=k): k + 1 kernel steps per kept result (reference samplers.py:191-207
    forwards the keyword): the thinned trace is the strided view of the unthinned one from the same seed."""
    kw = dict(num_chains=6, burn_in=20, seed=9, sampler_type=sampler_type)
    full = pm.sample(M.schools_pm4(), num_samples=30, **kw)
    thin = pm.sample(M.schools_pm4(), num_samples=10, num_steps_between_results=2, **kw)
    for name in ("schools_pm4/eta", "schools_pm4/mu", "schools_pm4/tau"):
        assert thin.posterior[name].shape[1] == 10
        np.testing.assert_array_equal(thin.posterior[name], full.posterior[name][:, 2::3])
    for name, v in thin.sample_stats.items():
        np.testing.assert_array_equal(v, full.sample_stats[name][:, 2::3])


def test_step_sizes_that_differ_between_chains():
    """`step_size` with a leading chain axis (reference samplers.py:406 passes it to tfp unchanged): with adaptation off
    every chain keeps its own step size -- small steps accept almost always, a huge one almost never."""
    eps = np.array([0.02, 0.02, 5.0, 5.0])
    # (the sampler class without `num_adaptation_steps`: no adaptation, like the reference's _BaseSampler)
    tr = pm.mcmc.samplers.HMC(M.schools_pm4(), step_size=eps)(num_samples=200, num_chains=4, burn_in=0, seed=2)
    acc = np.exp(np.minimum(tr.sample_stats["mean_tree_accept"], 0.0)).mean(axis=1)
    assert acc[0] > 0.9 and acc[1] > 0.9 and acc[2] < 0.3 and acc[3] < 0.3, acc


def test_models_of_one_structure_with_other_names_share_a_device_model():
    """Device models are cached by structure and layout, not by variable names: a second model of the same structure
    gets its own names in the trace, in the deterministics, in the pointwise log-likelihood and in the compound
    step's blocks (the cached model was built for the first one)."""
    kw = dict(num_chains=4, num_samples=20, burn_in=10, seed=3, include_log_likelihood=True)
    first, second = M.same_structure_models()
    t1 = pm.sample(first, **kw)
    t2 = pm.sample(second, **kw)
    assert {"first/a", "first/twice_a"} <= set(t1.posterior) and "first/ya" in t1.log_likelihood
    assert {"second/b", "second/twice_b"} <= set(t2.posterior) and "second/yb" in t2.log_likelihood
    assert not any(k.startswith("first/") for k in list(t2.posterior) + list(t2.log_likelihood))
    np.testing.assert_array_equal(t1.posterior["first/a"], t2.posterior["second/b"])
    np.testing.assert_array_equal(t1.posterior["first/twice_a"], t2.posterior["second/twice_b"])
    t3 = pm.sample(M.same_structure_models()[1], sampler_type="compound", num_chains=4, num_samples=20, burn_in=10, seed=3)
    assert "second/b" in t3.posterior and t3.posterior["second/b"].shape == (4, 20)
