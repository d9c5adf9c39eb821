"""The compound step (discrete latents): host logic on the CPU, sampling on the GPU.

Mirrors the reference's tests/test_compound.py and tests/test_discrete.py (same models, same assertions) and adds a
mixture model whose posterior is known in closed form, plus a detailed-balance check of the device proposals."""
import numpy as np
import pytest

import pymc4_b200 as pm
from pymc4_b200.ir import lower_model
from pymc4_b200.mcmc.samplers import NUTS, RandomWalkM, reg_samplers


def compound_model():
    @pm.model()
    def compound_model():
        var1 = yield pm.Normal("var1", 0, 1)
        var2 = yield pm.Bernoulli("var2", 0.5)
        return var2

    return compound_model()


def simple_model():
    @pm.model()
    def simple_model():
        var1 = yield pm.Normal("var1", 0, 1)
        return var1

    return simple_model()


def categorical_same_shape():
    @pm.model
    def categorical_same_shape():
        yield pm.Categorical("var1", probs=[0.2, 0.4, 0.4])
        yield pm.Categorical("var2", probs=[0.1, 0.3, 0.6])
        yield pm.Categorical("var3", probs=[0.1, 0.1, 0.8])

    return categorical_same_shape()


def categorical_different_shape():
    @pm.model
    def categorical_different_shape():
        yield pm.Categorical("var1", probs=[0.2, 0.4, 0.4])
        yield pm.Categorical("var2", probs=[0.1, 0.3, 0.1, 0.5])
        yield pm.Categorical("var3", probs=[0.1, 0.1, 0.1, 0.2, 0.5])

    return categorical_different_shape()


def model_symmetric():
    @pm.model
    def model_symmetric():
        yield pm.Bernoulli("var1", 0.1)
        yield pm.Bernoulli("var2", 1 - 0.1)

    return model_symmetric()


Y_MIX = np.array([1.9, 2.6, 1.4, 2.2, 3.0])


def mixture_model():
    """z ~ Bernoulli(0.3), mu ~ N(0, 1), y_i ~ N(mu + 2 z, 1): P(z = 1 | y) and E[mu | y] are closed-form."""

    @pm.model
    def mixture_model():
        z = yield pm.Bernoulli("z", 0.3)
        mu = yield pm.Normal("mu", 0.0, 1.0)
        yield pm.Normal("y", mu + 2.0 * z, 1.0, observed=Y_MIX)

    return mixture_model()


def mixture_posterior():
    n = len(Y_MIX)
    out = {}
    for z, prior in ((0, 0.7), (1, 0.3)):
        r = Y_MIX - 2.0 * z
        # integrate mu ~ N(0, 1) out: r ~ N(0, I + 1 1^T)
        cov = np.eye(n) + np.ones((n, n))
        sign, logdet = np.linalg.slogdet(cov)
        out[z] = (np.log(prior) - 0.5 * r @ np.linalg.solve(cov, r) - 0.5 * logdet, r.sum() / (n + 1.0))
    w1 = 1.0 / (1.0 + np.exp(out[0][0] - out[1][0]))
    return w1, (1 - w1) * out[0][1] + w1 * out[1][1]


def discrete_models():
    """the three models of the reference's tests/test_discrete.py"""

    @pm.model()
    def model_with_discrete_categorical():
        disc = yield pm.Categorical("disc", probs=[0.1, 0.9])
        return disc

    @pm.model()
    def model_with_discrete_bernoulli():
        disc = yield pm.Bernoulli("disc", 0.9)
        return disc

    @pm.model()
    def model_with_discrete_and_continuous():
        yield pm.Categorical("disc", probs=[0.1, 0.9])
        norm = yield pm.Normal("mu", 0, 1)
        return norm

    return [model_with_discrete_categorical(), model_with_discrete_bernoulli(), model_with_discrete_and_continuous()]


CAT_PROBS = [0.1, 0.3, 0.1, 0.5]


def categorical_normal_model():
    @pm.model
    def m():
        yield pm.Categorical("k", probs=CAT_PROBS)
        yield pm.Normal("x", 1.0, 2.0)

    return m()


def poisson_latent_model():
    @pm.model
    def m():
        yield pm.Poisson("n", 4.0)

    return m()


def all_models():
    """(model, needs float64 kernels) for every model structure of this file (prebuilt by __graft_entry__.build())"""
    return ([(m, False) for m in (compound_model(), simple_model(), model_symmetric(), poisson_latent_model())]
            + [(m, True) for m in (mixture_model(), categorical_normal_model())] + [(m, False) for m in discrete_models()]
            + [(mixture_with_deterministic(), False)])


# ---------------------------------------------------------------------------------------------------------------
# host logic (CPU)
# ---------------------------------------------------------------------------------------------------------------
def test_free_discrete_variables_lower_to_frozen_coordinates():
    ir, _, _ = lower_model(compound_model())
    flags = {rv.name: rv.discrete for rv in ir.rvs}
    assert flags == {"compound_model/var1": False, "compound_model/var2": True}
    with pytest.raises(ValueError):  # reference samplers.py:62-66
        NUTS(compound_model())
    with pytest.raises(ValueError):
        reg_samplers["hmc"](compound_model())


def test_categorical_log_mass_polynomial_is_exact_at_every_class():
    from oracle import pm4_oracle

    probs1, probs2 = [0.2, 0.4, 0.4], [0.1, 0.3, 0.1, 0.2, 0.3]

    @pm.model
    def cat():
        yield pm.Categorical("a", probs=probs1)
        yield pm.Categorical("b", probs=probs2)
        yield pm.Normal("n", 0, 1)

    ir, _, _ = lower_model(cat())
    orc = pm4_oracle.Oracle(ir)
    for a in range(3):
        for b in range(5):
            lp, _ = orc.logp_grad(np.array([[a, b, 0.3]], float), dtype=np.float64)
            ref = np.log(probs1[a]) + np.log(probs2[b]) - 0.5 * 0.09 - 0.5 * np.log(2 * np.pi)
            assert abs(lp[0] - ref) < 1e-12
            lp32, _ = orc.logp_grad(np.array([[a, b, 0.3]], float), dtype=np.float32)
            assert abs(lp32[0] - ref) < 1e-5
    d = pm.Categorical.dist(probs=probs2)
    np.testing.assert_allclose(d.log_prob(np.arange(5.0)), np.log(probs2))
    draws = d.sample(20000, seed=1)
    np.testing.assert_allclose(np.bincount(draws.astype(int), minlength=5) / 20000.0, probs2, atol=0.015)


def test_proposal_functions_compare_like_the_reference():
    assert pm.categorical_uniform_fn(classes=3) == pm.categorical_uniform_fn(classes=3)
    assert not (pm.categorical_uniform_fn(classes=3) == pm.categorical_uniform_fn(classes=4))
    assert not (pm.categorical_uniform_fn(classes=3) == pm.categorical_uniform_fn(classes=3, name="other"))
    assert pm.bernoulli_fn() == pm.bernoulli_fn() and not (pm.bernoulli_fn() == pm.gaussian_round_fn())
    assert pm.gaussian_round_fn(scale=2.0) == pm.gaussian_round_fn(scale=2.0)
    assert not (pm.gaussian_round_fn(scale=2.0) == pm.gaussian_round_fn(scale=1.0))
    # host statement of the proposals (the device applies the same maps to its own random stream)
    fn = pm.bernoulli_fn()()
    out = fn([np.zeros(4000), np.ones(4000)], 3)
    assert set(np.unique(out[0])) == {0.0, 1.0} and abs(out[0].mean() - 0.5) < 0.05 and abs(out[1].mean() - 0.5) < 0.05
    out = pm.categorical_uniform_fn(classes=4)()([np.zeros(8000)], 5)
    np.testing.assert_allclose(np.bincount(out[0].astype(int), minlength=4) / 8000.0, 0.25, atol=0.03)
    out = pm.gaussian_round_fn()()([np.full(8000, 3.0)], 7)
    assert np.all(out[0] == np.round(out[0])) and abs(out[0].mean() - 3.0) < 0.1
    with pytest.raises(ValueError):
        pm.gaussian_round_fn(scale=[1.0, 2.0])()([np.zeros(2)] * 3, 0)


def test_sampler_merging():
    """reference tests/test_compound.py:174-236"""
    sampler = reg_samplers["compound"]
    sampler1 = sampler(categorical_same_shape())
    sampler1._assign_default_methods()
    sampler2 = sampler(categorical_different_shape())
    sampler2._assign_default_methods()
    assert len(sampler1.kernel_kwargs["compound_samplers"]) == 1
    assert len(sampler2.kernel_kwargs["compound_samplers"]) == 3
    sampler_methods1 = [("var1", RandomWalkM)]
    sampler_methods2 = [("var1", RandomWalkM, {"new_state_fn": pm.categorical_uniform_fn(classes=3)})]
    sampler_methods3 = [
        ("var1", RandomWalkM, {"new_state_fn": pm.categorical_uniform_fn(classes=3, name="smth_different")})
    ]
    sampler_methods4 = [
        ("var1", RandomWalkM, {"new_state_fn": pm.categorical_uniform_fn(classes=3, name="smth_different")}),
        ("var3", RandomWalkM, {"new_state_fn": pm.categorical_uniform_fn(classes=3, name="smth_different")}),
    ]
    sampler_methods5 = [("var1", RandomWalkM), ("var2", RandomWalkM), ("var3", RandomWalkM)]
    sampler_ = sampler(categorical_same_shape())
    sampler_._assign_default_methods(sampler_methods=sampler_methods1)
    assert len(sampler_.kernel_kwargs["compound_samplers"]) == 1
    sampler_._assign_default_methods(sampler_methods=sampler_methods2)
    assert len(sampler_.kernel_kwargs["compound_samplers"]) == 1
    sampler_._assign_default_methods(sampler_methods=sampler_methods3)
    assert len(sampler_.kernel_kwargs["compound_samplers"]) == 2
    sampler_._assign_default_methods(sampler_methods=sampler_methods4)
    assert len(sampler_.kernel_kwargs["compound_samplers"]) == 2
    sampler_ = sampler(categorical_different_shape())
    sampler_._assign_default_methods(sampler_methods=sampler_methods5)
    assert len(sampler_.kernel_kwargs["compound_samplers"]) == 3
    # the blocks partition the free variables
    flat = sorted(n for block in sampler_._block_vars for n in block)
    assert flat == sorted("categorical_different_shape/var%d" % k for k in (1, 2, 3))


def test_sampler_assignment_rules():
    from pymc4_b200.inference.sampling import auto_assign_sampler, check_proposal_functions

    assert auto_assign_sampler(simple_model()) == "nuts"
    assert auto_assign_sampler(compound_model()) == "compound"
    assert auto_assign_sampler(compound_model(), "rwm") == "rwm"
    assert check_proposal_functions(compound_model()) and not check_proposal_functions(simple_model())
    s = reg_samplers["compound"](compound_model())
    s._assign_default_methods()
    kinds = {tuple(v): k[0]._name for k, v in zip(s.kernel_kwargs["compound_samplers"], s._block_vars)}
    assert kinds == {("compound_model/var2",): "rwm", ("compound_model/var1",): "nuts"}
    with pytest.raises(ValueError):  # a gradient sampler for a variable without a gradient
        s._assign_default_methods(sampler_methods=[("var2", NUTS)])
    with pytest.raises(ValueError):
        s._assign_default_methods(sampler_methods=[("var2",)])
    s._assign_default_methods(sampler_methods=[(["var1", "var2"], RandomWalkM)])
    assert len(s.kernel_kwargs["compound_samplers"]) == 2  # var2 carries bernoulli_fn, var1 the default proposal


def test_oracle_compound_targets_the_closed_form_posterior():
    """The oracle's compound transition (the checker of the GPU replay test) is itself pinned to a closed form."""
    from oracle import pm4_oracle
    from pymc4_b200._cstructs import PM4_EPS_POOLED, make_adapt
    from pymc4_b200.distributions.state_functions import PM4_RW_BERNOULLI

    ir, _, _ = lower_model(mixture_model())
    disc = np.concatenate([np.full(rv.size, 1.0 if rv.discrete else 0.0) for rv in ir.rvs])
    kinds = np.where(disc > 0, PM4_RW_BERNOULLI, 0).astype(np.int32)
    blocks = [dict(kind="rwm", dim_scale=disc, rw_kind=kinds),
              dict(kind="nuts", dim_scale=1 - disc, step_size=0.1, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=100))]
    out = pm4_oracle.Oracle(ir).compound(blocks, np.zeros((384, ir.D)), 500, 150, seed=3)
    w1, mu_mean = mixture_posterior()
    z, mu = out["states"][:, :, disc > 0][..., 0], out["states"][:, :, disc == 0][..., 0]
    se_z = z.mean(axis=0).std() / np.sqrt(z.shape[1])
    se_mu = mu.mean(axis=0).std() / np.sqrt(mu.shape[1])
    assert abs(z.mean() - w1) < 4 * se_z + 0.005, (z.mean(), w1, se_z)
    assert abs(mu.mean() - mu_mean) < 4 * se_mu + 0.005, (mu.mean(), mu_mean, se_mu)
    assert set(np.unique(z)) == {0.0, 1.0}


# ---------------------------------------------------------------------------------------------------------------
# sampling (GPU)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("seed", [3, 5])
@pytest.mark.parametrize("sampler_type", ["hmc", "nuts", "rwm", "compound"])
def test_samplers_on_compound_model(seed, sampler_type):
    """reference tests/test_compound.py:87-103"""

    def _execute():
        trace = pm.sample(compound_model(), sampler_type=sampler_type, seed=seed)
        var1 = round(float(trace.posterior["compound_model/var1"].mean()), 1)
        var2 = np.sum(trace.posterior["compound_model/var2"]) / (1000 * 10)
        np.testing.assert_allclose(var1, 0.0, atol=0.1)
        np.testing.assert_allclose(var2, 0.5, atol=0.1)

    if sampler_type in ["compound", "rwm"]:
        _execute()
    else:
        with pytest.raises(ValueError):
            _execute()


@pytest.mark.gpu
@pytest.mark.parametrize("sampler_type", ["rwm", "compound"])
def test_compound_model_sampler_method(sampler_type):
    """reference tests/test_compound.py:106-121"""
    trace = pm.sample(compound_model(), sampler_type=sampler_type, sampler_methods=[("var2", RandomWalkM)], seed=3)
    np.testing.assert_allclose(round(float(trace.posterior["compound_model/var1"].mean()), 1), 0.0, atol=0.1)
    np.testing.assert_allclose(np.sum(trace.posterior["compound_model/var2"]) / (1000 * 10), 0.5, atol=0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("sampler_type", ["hmc", "nuts", "rwm", "compound", "nuts_simple", "hmc_simple"])
def test_samplers_on_simple_model(sampler_type):
    """reference tests/test_compound.py:124-135"""
    trace = pm.sample(simple_model(), sampler_type=sampler_type, seed=11)
    np.testing.assert_allclose(round(float(trace.posterior["simple_model/var1"].mean()), 1), 0.0, atol=0.1)


@pytest.mark.gpu
def test_compound_seed():
    """reference tests/test_compound.py:149-171"""
    trace1 = pm.sample(compound_model(), seed=3)
    trace2 = pm.sample(compound_model(), seed=3)
    trace3 = pm.sample(compound_model(), seed=4)
    for name in ("compound_model/var1", "compound_model/var2"):
        assert np.array_equal(trace1.posterior[name], trace2.posterior[name])
    assert not np.array_equal(trace1.posterior["compound_model/var1"], trace3.posterior["compound_model/var1"])
    assert "sample_stats" not in trace1.groups()  # the compound sampler's statistics are not traced (reference :168)


@pytest.mark.gpu
def test_compound_symmetric():
    """reference tests/test_compound.py:247-259"""
    trace = pm.sample(model_symmetric(), seed=3)
    np.testing.assert_allclose(np.mean(trace.posterior["model_symmetric/var1"]), 0.1, atol=0.1)
    np.testing.assert_allclose(np.mean(trace.posterior["model_symmetric/var2"]), 0.9, atol=0.1)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [3, 5, 7])
def test_discrete_sampling(seed):
    """reference tests/test_discrete.py:47-68"""

    for model in discrete_models():
        trace = pm.sample(model=model, sampler_type="compound", seed=seed)
        value = round(float(trace.posterior[model.name + "/disc"].mean()), 1)
        np.testing.assert_allclose(value, 0.9, atol=0.1)


@pytest.mark.gpu
def test_categorical_frequencies_and_trace_discrete():
    probs = CAT_PROBS
    trace = pm.sample(categorical_normal_model(), num_chains=256, num_samples=400, burn_in=100, seed=1, trace_discrete=["k"])
    k = trace.posterior["m/k"]
    assert k.dtype == np.int32 and k.min() >= 0 and k.max() <= 3
    np.testing.assert_allclose(np.bincount(k.ravel(), minlength=4) / k.size, probs, atol=0.01)
    x = trace.posterior["m/x"]
    assert abs(x.mean() - 1.0) < 0.05 and abs(x.std() - 2.0) < 0.05


@pytest.mark.gpu
def test_mixture_posterior_against_the_closed_form():
    """The continuous block (NUTS on mu given z) and the discrete block (Bernoulli flips of z given mu) together
    target the joint posterior: P(z = 1 | y) and E[mu | y] against the closed form."""
    w1, mu_mean = mixture_posterior()
    trace = pm.sample(mixture_model(), num_chains=512, num_samples=500, burn_in=200, seed=5)
    z, mu = trace.posterior["mixture_model/z"], trace.posterior["mixture_model/mu"]
    assert set(np.unique(z)) <= {0.0, 1.0}
    # chains are independent: the standard error comes from the chain means
    se_z = z.mean(axis=1).std() / np.sqrt(z.shape[0])
    se_mu = mu.mean(axis=1).std() / np.sqrt(mu.shape[0])
    assert abs(z.mean() - w1) < max(5 * se_z, 0.01), (z.mean(), w1, se_z)
    assert abs(mu.mean() - mu_mean) < max(5 * se_mu, 0.01), (mu.mean(), mu_mean, se_mu)
    # the same target with the continuous variable moved by random-walk Metropolis and by HMC
    for methods in ([("mu", RandomWalkM)], [("mu", reg_samplers["hmc"])]):
        tr = pm.sample(mixture_model(), num_chains=512, num_samples=1500, burn_in=300, seed=6, sampler_type="compound",
                       sampler_methods=methods)
        zz = tr.posterior["mixture_model/z"]
        assert abs(zz.mean() - w1) < 0.02, (methods, zz.mean(), w1)


@pytest.mark.gpu
def test_free_poisson_with_the_gaussian_round_proposal():
    trace = pm.sample(poisson_latent_model(), sampler_type="compound", num_chains=512, num_samples=1000, burn_in=200, seed=2,
                      sampler_methods=[("n", RandomWalkM, {"new_state_fn": pm.gaussian_round_fn()})])
    n = trace.posterior["m/n"]
    assert np.all(n == np.round(n)) and n.min() >= 0
    assert abs(n.mean() - 4.0) < 0.1 and abs(n.var() - 4.0) < 0.3


def _replay_blocks(ir, which):
    """Block records (host arrays) for DeviceModel.compound / Oracle.compound."""
    from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt
    from pymc4_b200.distributions.state_functions import PM4_RW_BERNOULLI, PM4_RW_CATEGORICAL

    disc = np.concatenate([np.full(rv.size, 1.0 if rv.discrete else 0.0) for rv in ir.rvs])
    kinds = np.zeros(ir.D, np.int32)
    if which == "categorical":
        kinds[disc > 0] = PM4_RW_CATEGORICAL | (len(CAT_PROBS) << 8)
        return [dict(kind="rwm", dim_scale=disc, rw_kind=kinds), dict(kind="rwm", dim_scale=1 - disc, rw_scale=1.5)]
    kinds[disc > 0] = PM4_RW_BERNOULLI
    first = dict(kind="rwm", dim_scale=disc, rw_kind=kinds)
    if which == "nuts_pooled":
        return [first, dict(kind="nuts", dim_scale=1 - disc, step_size=0.1, max_tree_depth=6,
                            adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=12))]
    if which == "nuts_per_chain":
        return [dict(kind="nuts", dim_scale=1 - disc, step_size=0.1, max_tree_depth=6,
                     adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=12)), first]
    return [first, dict(kind="hmc", dim_scale=(1 - disc) * 0.5, step_size=0.2, num_leapfrog_steps=4,
                        adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=12))]


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["nuts_pooled", "nuts_per_chain", "hmc", "categorical"])
def test_compound_replay_against_the_oracle(which):
    """FP64 replay of pm4_compound_run against the oracle's restatement (oracle/pm4_oracle_impl.h: pm4o_compound_run):
    same Philox streams, same block order -- the discrete coordinates agree exactly, the continuous ones and the
    log acceptance ratios to rounding, the adapted step sizes too."""
    from oracle import pm4_oracle
    from pymc4_b200._cabi import DeviceModel

    model = categorical_normal_model() if which == "categorical" else mixture_model()
    ir, _, _ = lower_model(model)
    blocks = _replay_blocks(ir, which)
    C, B, S = 24, 14, 26
    rng = np.random.default_rng(3)
    theta0 = np.zeros((C, ir.D))
    for rv in ir.rvs:
        if not rv.discrete:
            theta0[:, rv.offset: rv.offset + rv.size] = rng.normal(0, 0.3, (C, rv.size))
    ref = pm4_oracle.Oracle(ir).compound(blocks, theta0, S, B, seed=17, dtype=np.float64)
    out = DeviceModel(ir, device=0, f64=True).compound(blocks, theta0, S, B, seed=17, dtype=np.float64)
    st, lacc = out["states"].cpu().numpy(), out["log_accept"].cpu().numpy()
    disc = np.concatenate([np.full(rv.size, rv.discrete) for rv in ir.rvs])
    assert np.array_equal(st[:, :, disc], ref["states"][:, :, disc])
    assert len(np.unique(st[:, :, disc])) > 1  # the discrete block moves
    np.testing.assert_allclose(st[:, :, ~disc], ref["states"][:, :, ~disc], atol=1e-9)
    np.testing.assert_allclose(out["lp"].cpu().numpy(), ref["lp"], rtol=1e-10, atol=1e-10)
    finite = np.isfinite(ref["log_accept"])
    assert np.array_equal(np.isfinite(lacc), finite)
    np.testing.assert_allclose(lacc[finite], ref["log_accept"][finite], atol=1e-8)
    np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-9)


def mixture_with_deterministic():
    @pm.model
    def mixture_det():
        z = yield pm.Bernoulli("z", 0.3)
        mu = yield pm.Normal("mu", 0.0, 1.0)
        shifted = yield pm.Deterministic("shifted", mu + 2.0 * z)
        yield pm.Normal("y", shifted, 1.0, observed=Y_MIX)

    return mixture_det()


@pytest.mark.gpu
def test_compound_with_deterministics_log_likelihood_and_predictive():
    """The trace of a compound run carries deterministics, the pointwise log-likelihood and feeds posterior-predictive
    sampling like any other sampler's (reference samplers.py:136-170 is sampler-agnostic)."""
    model = mixture_with_deterministic()
    trace = pm.sample(model, num_chains=64, num_samples=200, burn_in=100, seed=4, include_log_likelihood=True)
    post = trace.posterior
    z, mu, sh = post["mixture_det/z"], post["mixture_det/mu"], post["mixture_det/shifted"]
    np.testing.assert_allclose(sh, mu + 2.0 * z, rtol=1e-6, atol=1e-6)
    ll = trace.log_likelihood["mixture_det/y"]
    assert ll.shape == (64, 200, len(Y_MIX))
    ref = -0.5 * (Y_MIX[None, None, :] - sh[..., None]) ** 2 - 0.5 * np.log(2 * np.pi)
    np.testing.assert_allclose(ll, ref, rtol=1e-4, atol=1e-4)
    ppc = pm.sample_posterior_predictive(model, trace)
    yp = ppc.posterior_predictive["mixture_det/y"]
    assert yp.shape == (64, 200, len(Y_MIX))
    assert abs(yp.mean() - sh.mean()) < 0.05 and abs(yp.std() - np.sqrt(1.0 + sh.var())) < 0.1
