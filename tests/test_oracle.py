"""CPU tests of the oracle (oracle/pm4_oracle.c) against the reference's golden values, scipy,
torch autograd and published known-answer vectors.  No GPU needed."""
import math

import numpy as np
import pymc4_b200 as pm
import pytest
import torch
from scipy import stats

from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_hmc_cfg, make_nuts_cfg
from pymc4_b200.ir import lower_model
from tests import models as M


def test_eight_schools_golden(oracle_mod, golden):
    """The reference's own golden value, tests/test_8schools.py:50-61."""
    ir, _, _ = lower_model(M.schools_pm4())
    th = ir.join_theta({"schools_pm4/eta": np.zeros(8), "schools_pm4/mu": 0.0, "schools_pm4/__log_tau": 1.0})
    o = oracle_mod.Oracle(ir)
    lp64, _ = o.logp_grad(th, dtype=np.float64)
    lp32, _ = o.logp_grad(th, dtype=np.float32)
    assert abs(lp64[0] - golden["eight_schools"]["scipy_f64"]) < 1e-12
    np.testing.assert_almost_equal(lp32[0], golden["eight_schools"]["reference_assert_f32"])  # same assert as the reference
    assert np.float32(lp64[0]) == np.float32(golden["eight_schools"]["reference_assert_f32"])


def test_radon_golden(oracle_mod, golden):
    ir, _, _ = lower_model(M.radon_model(real=True))
    assert [rv.name.split("/")[-1] for rv in ir.rvs] == [
        "mu_alpha", "mu_beta", "alpha", "beta", "__log_sigma_alpha", "__log_sigma_beta", "__log_eps"]
    o = oracle_mod.Oracle(ir)
    pts = np.stack([np.zeros(ir.D), np.array(golden["radon_919"]["theta_rand"])])
    lp, _ = o.logp_grad(pts, dtype=np.float64)
    np.testing.assert_allclose(lp, [golden["radon_919"]["logp_at_zero"], golden["radon_919"]["logp_at_rand"]], rtol=1e-13)


def _torch_radon_logp(theta, idx, x, y, G):
    mu_a, mu_b = theta[0], theta[1]
    a, b = theta[2:2 + G], theta[2 + G:2 + 2 * G]
    lsa, lsb, le = theta[2 + 2 * G], theta[3 + 2 * G], theta[4 + 2 * G]
    sa, sb, eps = torch.exp(lsa), torch.exp(lsb), torch.exp(le)
    N = torch.distributions.Normal
    HC = torch.distributions.HalfCauchy
    lp = N(0., 1.).log_prob(mu_a) + N(0., 1.).log_prob(mu_b)
    lp = lp + HC(1.).log_prob(sa) + lsa + HC(1.).log_prob(sb) + lsb + HC(1.).log_prob(eps) + le
    lp = lp + N(mu_a, sa).log_prob(a).sum() + N(mu_b, sb).log_prob(b).sum()
    lp = lp + N(a[idx] + b[idx] * x, eps).log_prob(y).sum()
    return lp


def test_radon_gradient_vs_torch_autograd(oracle_mod):
    idx, x, y, G = M.radon_real()
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, G))
    rng = np.random.default_rng(3)
    th = rng.normal(0, 0.3, ir.D)
    lp, g = oracle_mod.Oracle(ir).logp_grad(th, dtype=np.float64)
    t = torch.tensor(th, dtype=torch.float64, requires_grad=True)
    ref = _torch_radon_logp(t, torch.tensor(idx, dtype=torch.long), torch.tensor(x), torch.tensor(y), G)
    ref.backward()
    assert abs(lp[0] - ref.item()) < 1e-9 * abs(ref.item())
    np.testing.assert_allclose(g[0], t.grad.numpy(), rtol=1e-8, atol=1e-8)  # x/s - mu/s vs (x - mu)/s


@pytest.mark.parametrize("name", ["logistic", "poisson", "nested", "eight_schools"])
def test_gradient_vs_finite_differences(name, oracle_mod):
    mk = {"logistic": lambda: M.logistic_model(*M.logistic_data()), "poisson": lambda: M.poisson_model(*M.poisson_data()),
          "nested": lambda: M.outer_model(), "eight_schools": lambda: M.schools_pm4()}[name]
    ir, _, _ = lower_model(mk())
    o = oracle_mod.Oracle(ir)
    th = np.random.default_rng(4).normal(0, 0.5, ir.D)
    _, g = o.logp_grad(th)
    fd = np.zeros(ir.D)
    for j in range(ir.D):
        e = np.zeros(ir.D); e[j] = 1e-6
        fd[j] = (o.logp_grad(th + e)[0][0] - o.logp_grad(th - e)[0][0]) / 2e-6
    np.testing.assert_allclose(g[0], fd, rtol=2e-6, atol=2e-6)


def test_densities_vs_scipy(oracle_mod):
    """Same recipe as the reference's tests/test_sampling.py:43-61: log-density == scipy sums."""
    import pymc4_b200 as pm

    rng = np.random.default_rng(8)
    yn, yb, yp = rng.normal(size=7), (rng.random(9) < 0.4).astype(float), rng.poisson(3.0, 11).astype(float)

    @pm.model
    def m():
        mu = yield pm.Normal("mu", 0.5, 2.0)
        s = yield pm.HalfNormal("s", 1.5)
        c = yield pm.HalfCauchy("c", 0.7)
        yield pm.Normal("yn", mu, s, observed=yn)
        yield pm.Bernoulli("yb", pm.math.sigmoid(mu), observed=yb)
        yield pm.Poisson("yp", c + 1.0, observed=yp)

    ir, _, _ = lower_model(m())
    assert [rv.name for rv in ir.rvs] == ["m/mu", "m/__log_s", "m/__log_c"]
    for _ in range(5):
        mu, ls, lc = rng.normal(0, 0.7, 3)
        s, c = math.exp(ls), math.exp(lc)
        p = 1 / (1 + math.exp(-mu))
        ref = (stats.norm.logpdf(mu, 0.5, 2.0) + stats.halfnorm.logpdf(s, scale=1.5) + ls
               + stats.halfcauchy.logpdf(c, scale=0.7) + lc + stats.norm.logpdf(yn, mu, s).sum()
               + stats.bernoulli.logpmf(yb, p).sum() + stats.poisson.logpmf(yp, c + 1.0).sum())
        lp, _ = oracle_mod.Oracle(ir).logp_grad(np.array([mu, ls, lc]))
        np.testing.assert_allclose(lp[0], ref, rtol=1e-12)
        lp32, _ = oracle_mod.Oracle(ir).logp_grad(np.array([mu, ls, lc]), dtype=np.float32)
        np.testing.assert_allclose(lp32[0], ref, rtol=1e-5)  # the reference's own tolerance


def test_log_transform_density(oracle_mod):
    """tests/test_executor.py:277-298 restated: HalfNormal through the Log transform equals the
    density of the transformed distribution."""
    import pymc4_b200 as pm

    @pm.model
    def m():
        yield pm.HalfNormal("n", 1.0)

    ir, _, _ = lower_model(m())
    for z in (-1.3, 0.0, 0.4, 2.0):
        lp, _ = oracle_mod.Oracle(ir).logp_grad(np.array([z]))
        np.testing.assert_allclose(lp[0], stats.halfnorm.logpdf(math.exp(z)) + z, rtol=1e-12)


def test_philox_known_answers(oracle_mod):
    """Random123 kat_vectors for philox4x32-10."""
    assert oracle_mod.philox4x32_10([0] * 4, [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert oracle_mod.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert oracle_mod.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [
        0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    m = np.concatenate([oracle_mod.draw_momentum(7, c, t, 175, np.float64) for c in range(8) for t in range(8)])
    assert abs(m.mean()) < 0.03 and abs(m.std() - 1) < 0.03  # Box-Muller normals


# --- tfp.mcmc.nuts U-turn bookkeeping: published generator restated, against the closed form ---------
def _build_tree_uturn_instruction(max_depth, init_memory=0):
    def _buildtree(address, depth):
        if depth == 0:
            address += 1
            return address, address
        address_left, address_right = _buildtree(address, depth - 1)
        _, address_right = _buildtree(address_right, depth - 1)
        instruction.append((address_left, address_right))
        return address_left, address_right

    instruction = []
    _buildtree(init_memory, max_depth)
    return np.array(instruction, dtype=np.int32)


def _generate_efficient_write_read_instruction(instruction_array):
    nsteps = np.max(instruction_array) + 1
    mat = np.zeros((nsteps, nsteps))
    for a, b in instruction_array:
        mat[a, b] = 1
    for i in range(nsteps):
        temp = mat[i]
        endpoint = np.where(temp == 1)[0]
        if endpoint.size > 0:
            temp[:i] = -1
            temp[endpoint[-1] + 1:] = -1
            mat[i] = temp
        else:
            mat[i] = -1
    to_write = np.sum(mat > -1, axis=0)
    write = to_write - 1
    write[np.diag(mat) == -1] = max(to_write)
    read = []
    for i in range(nsteps):
        col = mat[:, i]
        if np.sum(col == 1) > 0:
            r = np.where(col[col >= 0] == 1)[0][0]
            read.append([r, r + np.sum(col == 1)])
        else:
            read.append([0, 0])
    return write, np.asarray(read)


@pytest.mark.parametrize("max_depth", [1, 3, 6, 10])
def test_uturn_checkpoint_closed_form(max_depth):
    """oracle and kernels use: even leaf i -> write slot popcount(i); odd leaf i -> read slots
    [popcount(i) - trailing_ones(i), popcount(i)).  Equal to TFP's instruction tables."""
    write, read = _generate_efficient_write_read_instruction(_build_tree_uturn_instruction(max_depth, init_memory=-1))
    assert len(write) == 1 << max_depth and max(write) == max_depth
    for i in range(1 << max_depth):
        pc = bin(i).count("1")
        t1 = 0
        while (i >> t1) & 1:
            t1 += 1
        if i % 2 == 0:
            assert write[i] == pc and tuple(read[i]) == (0, 0)
        else:
            assert write[i] == max_depth and tuple(read[i]) == (pc - t1, pc)


def test_dual_averaging_recursion(oracle_mod):
    """Hand-evaluate the tfp DualAveragingStepSizeAdaptation recursion for a chain whose
    acceptance probability is known: a Gaussian target with a tiny step accepts with probability ~1."""
    import pymc4_b200 as pm

    @pm.model
    def m():
        yield pm.Normal("x", 0.0, 1.0)

    ir, _, _ = lower_model(m())
    B = 6
    cfg = make_hmc_cfg(1, 1, B, step_size=1e-4, num_leapfrog_steps=1, seed=5, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B))
    out = oracle_mod.Oracle(ir).hmc(cfg, np.zeros(1), dtype=np.float64)
    eps0, err, log_avg, eps = 1e-4, 0.0, 0.0, 1e-4
    for t in range(B + 1):
        step = t + 1.0
        err += 0.75 - 1.0  # accept prob == 1 to ~1e-8 while eps stays tiny
        log_eps = math.log(10 * eps0) - err * math.sqrt(step) / ((10 + step) * 0.05)
        eta = step ** -0.75
        log_avg = eta * log_eps + (1 - eta) * log_avg
        eps = math.exp(log_eps) if t < B else math.exp(log_avg)
        if eps > 0.05:
            break  # beyond this the acceptance is no longer ~1
    else:
        np.testing.assert_allclose(out["step_size"][0], eps, rtol=1e-4)
    assert out["step_size"][0] > 1e-4  # the step size grew as the recursion demands


def test_shrinkage_target_and_per_chain_initial_step_sizes(oracle_mod):
    """tfp `shrinkage_target` replaces 10 x the initial step size in the recursion; a `step_size` with a leading chain
    axis starts every chain from its own value (each then shrinks towards 10 x ITS initial step size)."""
    import pymc4_b200 as pm

    @pm.model
    def m():
        yield pm.Normal("x", 0.0, 1.0)

    ir, _, _ = lower_model(m())
    B, target = 2, 3e-4
    cfg = make_hmc_cfg(1, 1, B, step_size=1e-4, num_leapfrog_steps=1, seed=5,
                       adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B, shrinkage_target=target))
    out = oracle_mod.Oracle(ir).hmc(cfg, np.zeros(1), dtype=np.float64)
    err, log_avg = 0.0, 0.0
    for t in range(B + 1):  # accept prob == 1 to ~1e-8 while eps stays tiny
        step = t + 1.0
        err += 0.75 - 1.0
        log_eps = math.log(target) - err * math.sqrt(step) / ((10 + step) * 0.05)
        log_avg = step ** -0.75 * log_eps + (1 - step ** -0.75) * log_avg
    np.testing.assert_allclose(out["step_size"][0], math.exp(log_avg), rtol=1e-5)
    # per-chain initial step sizes, no adaptation: the step sizes come back unchanged, and the first proposal of chain c
    # is the leapfrog with ITS step size
    from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN

    eps_c = np.array([0.05, 0.2, 0.8])
    cfg = make_hmc_cfg(3, 1, 0, step_size=0.05, num_leapfrog_steps=1, seed=9, adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, enabled=False))
    mom = np.ones((1, 3, 1))
    out = oracle_mod.Oracle(ir).hmc(cfg, np.full((3, 1), 0.5), dtype=np.float64, momenta=mom, step_sizes=eps_c)
    np.testing.assert_allclose(out["step_size"], eps_c)
    # one leapfrog from x = 0.5, p = 1 on the standard normal: x' = x + e (p - e x / 2)
    np.testing.assert_allclose(out["traj"][0, :, 0], 0.5 + eps_c * (1.0 - 0.5 * eps_c * 0.5), rtol=1e-12)


def test_nuts_standard_normal_and_determinism(oracle_mod):
    """Statistical sanity of the restated NUTS (reference tests/test_compound.py:124-146: |mean| < 0.1,
    same seed => same trace)."""
    import pymc4_b200 as pm

    @pm.model
    def m():
        yield pm.Normal("x", 0.0, 1.0, batch_stack=3)

    ir, _, _ = lower_model(m())
    cfg = make_nuts_cfg(32, 300, 100, step_size=0.1, seed=11, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=100))
    o = oracle_mod.Oracle(ir)
    a = o.nuts(cfg, np.zeros(3), dtype=np.float64)
    b = o.nuts(cfg, np.zeros(3), dtype=np.float64)
    assert np.array_equal(a["states"], b["states"]) and np.array_equal(a["leapfrogs"], b["leapfrogs"])
    x = a["states"].reshape(-1, 3)
    assert np.all(np.abs(x.mean(0)) < 0.1) and np.all(np.abs(x.var(0) - 1) < 0.15)
    assert a["diverging"].sum() == 0
    assert np.all(a["leapfrogs"] == (1 << a["depth"]) - 1) or np.all(a["leapfrogs"] <= (1 << a["depth"]) - 1)
    c = o.nuts(make_nuts_cfg(32, 300, 100, step_size=0.1, seed=12, adapt=cfg.adapt), np.zeros(3), dtype=np.float64)
    assert not np.array_equal(a["states"], c["states"])
    # per-chain adaptation gives different step sizes per chain, pooled a single one
    d = o.nuts(make_nuts_cfg(8, 10, 50, step_size=0.1, seed=1, adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=50)),
               np.zeros(3), dtype=np.float64)
    assert len(set(np.round(d["step_size"], 12))) > 1 and len(set(np.round(a["step_size"], 12))) == 1


def test_deterministics_and_float32_path(oracle_mod):
    ir, _, names = lower_model(M.outer_model())
    assert [d.name for d in ir.dets] == names
    th = np.random.default_rng(0).normal(size=(5, ir.D))
    out = oracle_mod.Oracle(ir).deterministics(th)
    k = {d.name: d.out_offset for d in ir.dets}
    cond = np.exp(th[:, ir.rv_by_name("outer_model/__log_cond").offset])
    x = th[:, ir.rv_by_name("outer_model/nested_model/x").offset]
    np.testing.assert_allclose(out[:, k["outer_model/dcond"]], 2 * cond)
    np.testing.assert_allclose(out[:, k["outer_model/nested_model/dx"]], x + 1)
    np.testing.assert_allclose(out[:, k["outer_model/ddx"]], x + 1)
    np.testing.assert_allclose(out[:, k["outer_model/cond"]], cond)


def test_pointwise_log_likelihood_vs_scipy(oracle_mod):
    """dist.log_prob(observed) per state, no reduction (/root/reference/pymc4/mcmc/samplers.py:832-854),
    against scipy for a Normal, a Bernoulli and a Poisson likelihood; the row sums up to the observed
    terms' share of the joint log-density."""
    from scipy import stats

    from pymc4_b200.ir import lower_model
    from tests import models as M

    rng = np.random.default_rng(3)
    # radon: y ~ Normal(a[idx] + b[idx] x, eps)
    idx, x, y, g = M.radon_synthetic(200, 11)
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, g))
    (obs,) = [t for t in ir.terms if t.observed]
    assert obs.label.endswith("y_like") and ir.ll_row == 200 and obs.shape == (200,)
    theta = rng.normal(0, 0.3, (5, ir.D))
    ll = oracle_mod.Oracle(ir).log_likelihood(theta)
    parts = ir.split_theta(theta)
    a, b = parts["hierarchical_model/alpha"], parts["hierarchical_model/beta"]
    eps = np.exp(parts["hierarchical_model/__log_eps"])
    want = stats.norm(a[:, idx] + b[:, idx] * x, eps[:, None]).logpdf(y)
    np.testing.assert_allclose(ll, want, rtol=1e-12, atol=1e-12)
    # leading axes are kept: [S, C, D] -> [S, C, ll_row]
    assert oracle_mod.Oracle(ir).log_likelihood(theta.reshape(5, 1, -1)).shape == (5, 1, 200)

    X, yb = M.logistic_data()
    ir, _, _ = lower_model(M.logistic_model(X, yb))
    theta = rng.normal(0, 0.5, (4, ir.D))
    beta = ir.split_theta(theta)["logistic_model/beta"]
    p = 1 / (1 + np.exp(-(beta @ X.T)))
    np.testing.assert_allclose(oracle_mod.Oracle(ir).log_likelihood(theta), stats.bernoulli(p).logpmf(yb), rtol=1e-10)

    idxp, yp, gp = M.poisson_data()
    ir, _, _ = lower_model(M.poisson_model(idxp, yp, gp))
    theta = rng.normal(0, 0.5, (4, ir.D))
    parts = ir.split_theta(theta)
    rate = np.exp(parts["poisson_model/intercept"][:, None] + parts["poisson_model/eff"][:, idxp])
    np.testing.assert_allclose(oracle_mod.Oracle(ir).log_likelihood(theta), stats.poisson(rate).logpmf(yp), rtol=1e-10)


def test_posterior_predictive_draws_have_the_right_moments(oracle_mod):
    """One draw of every observed variable per state (forward_sampling.py:230-427): the oracle's Philox-driven
    samplers against the analytic moments of Normal / Bernoulli / Poisson (inversion below rate 10, PTRS above)."""
    from pymc4_b200.ir import lower_model
    from tests import models as M

    rng = np.random.default_rng(9)
    idx, x, y, g = M.radon_synthetic(400, 7)
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, g))
    theta = np.tile(rng.normal(0, 0.3, ir.D), (3000, 1))
    d = oracle_mod.Oracle(ir).posterior_predictive(theta, seed=11)
    parts = ir.split_theta(theta[0])
    mu = parts["hierarchical_model/alpha"][idx] + parts["hierarchical_model/beta"][idx] * x
    eps = float(np.exp(parts["hierarchical_model/__log_eps"]))
    assert d.shape == (3000, 400)
    assert np.max(np.abs(d.mean(0) - mu)) < 5 * eps / np.sqrt(3000)
    assert abs(d.std() ** 2 / (eps ** 2 + mu.var()) - 1) < 0.02
    # same seed, same draws; another seed, other draws; states are independent streams
    assert np.array_equal(d, oracle_mod.Oracle(ir).posterior_predictive(theta, seed=11))
    assert not np.array_equal(d, oracle_mod.Oracle(ir).posterior_predictive(theta, seed=12))
    assert abs(np.corrcoef(d[0], d[1])[0, 1] - np.corrcoef(mu, mu + 0)[0, 1] * mu.var() / (mu.var() + eps ** 2)) < 0.15

    X, yb = M.logistic_data()
    ir, _, _ = lower_model(M.logistic_model(X, yb))
    theta = np.tile(rng.normal(0, 0.5, ir.D), (4000, 1))
    d = oracle_mod.Oracle(ir).posterior_predictive(theta, seed=3)
    p = 1 / (1 + np.exp(-(ir.split_theta(theta[0])["logistic_model/beta"] @ X.T)))
    assert set(np.unique(d)) <= {0.0, 1.0} and np.max(np.abs(d.mean(0) - p)) < 5 * 0.5 / np.sqrt(4000)

    idxp, yp, gp = M.poisson_data()
    ir, _, _ = lower_model(M.poisson_model(idxp, yp, gp))
    for intercept in (0.5, 3.5):  # rates ~1.6 (inversion) and ~33 (PTRS)
        th = np.zeros(ir.D)
        th[ir.rv_by_name("poisson_model/intercept").offset] = intercept
        d = oracle_mod.Oracle(ir).posterior_predictive(np.tile(th, (4000, 1)), seed=5)
        lam = np.exp(intercept)
        assert np.all(d >= 0) and np.all(d == np.round(d))
        assert abs(d.mean() / lam - 1) < 0.01 and abs(d.var() / lam - 1) < 0.03


def test_simple_adaptation_and_random_walk_metropolis(oracle_mod):
    """tfp SimpleStepSizeAdaptation (step *= or /= 1 + rate by the sign of accept - target) and
    RandomWalkMetropolis (state + scale * N(0, 1), symmetric MH) as the oracle restates them."""
    from pymc4_b200._cstructs import (PM4_ADAPT_SIMPLE, PM4_EPS_PER_CHAIN, PM4_PROPOSAL_RANDOM_WALK, make_adapt,
                                      make_hmc_cfg)
    from pymc4_b200.ir import lower_model
    from tests import models as M

    ir, _, _ = lower_model(M.simple_model())
    orc = oracle_mod.Oracle(ir)
    C, B, S = 64, 200, 400
    cfg = make_hmc_cfg(C, S, B, step_size=0.3, num_leapfrog_steps=3, seed=2,
                       adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B, kind=PM4_ADAPT_SIMPLE,
                                        adaptation_rate=0.01, target_accept_prob=0.75))
    out = orc.hmc(cfg, np.zeros(ir.D), dtype=np.float64)
    k = np.log(out["step_size"] / 0.3) / np.log(1.01)
    np.testing.assert_allclose(k, np.round(k), atol=1e-8)  # whole powers of (1 + rate)
    assert np.all(np.abs(k) <= B) and np.all((np.round(k) - B) % 2 == 0)  # B moves of +-1
    x = out["states"][..., 0]
    assert abs(x.mean()) < 0.1 and abs(x.std() - 1) < 0.1
    cfg = make_hmc_cfg(256, 2000, 500, step_size=1.0, num_leapfrog_steps=0, seed=4, adapt=make_adapt(enabled=False),
                       proposal=PM4_PROPOSAL_RANDOM_WALK, rw_scale=1.0)
    out = orc.hmc(cfg, np.zeros(ir.D), dtype=np.float64)
    x = out["states"][..., 0]
    assert abs(x.mean()) < 0.05 and abs(x.std() - 1) < 0.05
    acc = out["accepted"].mean()
    assert 0.6 < acc < 0.8  # theory for N(0,1) target with unit normal proposals: 2/pi * atan(2) = 0.705


def test_prior_predictive_ancestral_sampling(oracle_mod):
    """forward_sampling.py:26-227 restated: free variables in model order given the ones before, then the
    observed variables -- eight schools (eta ~ N(0,1), mu ~ N(1,10), tau ~ HalfNormal(2), obs ~ N(mu + tau eta, sigma))
    and the radon hierarchy (alpha[g] ~ N(mu_alpha, sigma_alpha))."""
    from pymc4_b200.ir import lower_model
    from tests import models as M

    ir, _, _ = lower_model(M.schools_pm4())
    x, obs = oracle_mod.Oracle(ir).prior_predictive(40000, seed=1)
    eta, mu, tau = x[:, :8], x[:, 8], x[:, 9]  # constrained values: tau itself, not log tau
    assert np.max(np.abs(eta.mean(0))) < 0.03 and np.max(np.abs(eta.std(0) - 1)) < 0.03
    assert abs(mu.mean() - 1) < 0.2 and abs(mu.std() / 10 - 1) < 0.03
    assert tau.min() > 0 and abs(tau.mean() / (2 * np.sqrt(2 / np.pi)) - 1) < 0.03
    want_var = 100 + 4 + M.SIGMA8 ** 2  # var(mu) + E[tau^2] var(eta) + sigma^2
    assert np.max(np.abs(obs.var(0) / want_var - 1)) < 0.06
    # conditional structure: obs - (mu + tau eta) is N(0, sigma)
    z = (obs - (mu[:, None] + tau[:, None] * eta)) / M.SIGMA8
    assert np.max(np.abs(z.std(0) - 1)) < 0.03 and np.max(np.abs(z.mean(0))) < 0.03

    idx, xx, y, g = M.radon_synthetic(120, 5)
    ir, _, _ = lower_model(M.hierarchical_model(idx, xx, y, g))
    x, obs = oracle_mod.Oracle(ir).prior_predictive(20000, seed=2)
    p = ir.split_theta(x)  # slices of the CONSTRAINED vector, named after the sampled (transformed) names
    mu_a, sig_a, alpha = p["hierarchical_model/mu_alpha"], p["hierarchical_model/__log_sigma_alpha"], p["hierarchical_model/alpha"]
    assert np.all(sig_a > 0) and abs(np.median(sig_a) - 1) < 0.05  # HalfCauchy(1): median 1
    z = (alpha - mu_a[:, None]) / sig_a[:, None]
    assert np.max(np.abs(z.mean(0))) < 0.03 and np.max(np.abs(z.std(0) - 1)) < 0.03
    assert obs.shape == (20000, 120) and np.all(np.isfinite(obs))


def test_dense_glm_term_lowering_and_oracle(oracle_mod):
    """Bernoulli(sigmoid(X @ beta)) lowers to ONE dense term (no tape) and the oracle evaluates it like the
    explicit-sum model and like scipy (the reference: tfd.Bernoulli(probs=...).log_prob, discrete.py:75-78)."""
    from scipy import stats

    from pymc4_b200.ir import TERM_KINDS, lower_model
    from tests import models as M

    X, y = M.logistic_glm_data(n=500, p=12)
    ir, _, _ = lower_model(M.logistic_glm_model(X, y))
    (g,) = ir.glm_terms
    assert g.kind == TERM_KINDS["bernoulli_logit_glm"] and g.n == 500 and g.glm.P == 12 and not g.ops and g.observed
    assert g.glm.base == ir.rv_by_name("logistic_glm_model/beta").offset and ir.ll_row == 500
    assert ir.team_warps == 1  # the dense term is not the teams' work
    rng = np.random.default_rng(2)
    theta = rng.normal(0, 0.5, (6, ir.D))
    lp, grad = oracle_mod.Oracle(ir).logp_grad(theta)
    parts = ir.split_theta(theta)
    mu, beta, tau = parts["logistic_glm_model/mu"], parts["logistic_glm_model/beta"], np.exp(parts["logistic_glm_model/__log_tau"])
    p = 1 / (1 + np.exp(-(beta @ X.T)))
    want = (stats.norm(0, 1).logpdf(mu) + stats.halfnorm(scale=1).logpdf(tau) + parts["logistic_glm_model/__log_tau"]
            + stats.norm(mu[:, None], tau[:, None]).logpdf(beta).sum(1) + stats.bernoulli(p).logpmf(y).sum(1))
    np.testing.assert_allclose(lp, want, rtol=1e-12)
    eps = 1e-6
    for j in (0, 3, ir.D - 1):
        d = np.zeros(ir.D); d[j] = eps
        fd = (oracle_mod.Oracle(ir).logp_grad(theta + d)[0] - oracle_mod.Oracle(ir).logp_grad(theta - d)[0]) / (2 * eps)
        np.testing.assert_allclose(grad[:, j], fd, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle_mod.Oracle(ir).log_likelihood(theta), stats.bernoulli(p).logpmf(y), rtol=1e-10)
    d = oracle_mod.Oracle(ir).posterior_predictive(np.tile(theta[0], (3000, 1)), seed=4)
    assert np.max(np.abs(d.mean(0) - p[0])) < 5 * 0.5 / np.sqrt(3000)
    # anything that is not exactly that pattern still needs the tape (and matvec has no tape op)
    import pymc4_b200 as pm

    @pm.model
    def shifted(X, y):
        beta = yield pm.Normal("beta", 0.0, 1.0, batch_stack=X.shape[1])
        yield pm.Bernoulli("obs", pm.math.sigmoid(X @ beta + 1.0), observed=y)

    with pytest.raises(NotImplementedError):
        lower_model(shifted(X, y))


def test_gp_marginal_likelihood_oracle_against_scipy(oracle_mod):
    """The GP term of the oracle: tfd.MultivariateNormalFullCovariance.log_prob of an ExpQuad + noise covariance
    (/root/reference/pymc4/gp/cov.py:399-475; the class docstring's 3-point example is the golden vector for the kernel
    matrix), against scipy's MVN density and central differences."""
    from scipy.stats import halfcauchy, halfnorm, multivariate_normal

    k = pm.gp.cov.ExpQuad(amplitude=1.0, length_scale=1.0, feature_ndims=1)
    K = k(np.array([[1.0], [2.0], [3.0]]), np.array([[3.0], [2.0], [1.0]]))
    np.testing.assert_allclose(K, [[0.13533528, 0.60653067, 1.0], [0.60653067, 1.0, 0.60653067], [1.0, 0.60653067, 0.13533528]],
                               rtol=1e-7)
    np.testing.assert_allclose(k.evaluate_kernel(np.array([[1.0], [2.0], [3.0]]), np.array([[3.0], [2.0], [1.0]])),
                               [0.13533528, 1.0, 0.13533528], rtol=1e-7)
    np.testing.assert_allclose(pm.gp.util.stabilize(np.eye(3)), 1.0001 * np.eye(3))
    X, y = M.gp_data(n=40, d=2)
    ir, _, _ = lower_model(M.gp_model(X, y))
    t = ir.gp_terms[0]
    assert t.kind == 7 and t.n == 1 and t.observed and t.gp.elems == (0, 1, 2) and ir.ll_row == 1
    np.testing.assert_allclose(ir.dataf[t.gp.par_blob:t.gp.par_blob + 8], [0, 1, 2, 0, 0, 0, 1, 1e-6])
    rng = np.random.default_rng(5)
    theta = rng.normal(0, 0.3, (4, 3))
    theta[:, 2] -= 1.5

    def ref(th):
        ls, amp, sg = np.exp(th)
        d2 = np.sum(np.square(X[:, None] - X[None]), -1)
        Kc = amp ** 2 * np.exp(-d2 / (2 * ls ** 2)) + (sg ** 2 + 1e-6) * np.eye(len(y))
        return (multivariate_normal(np.zeros(len(y)), Kc).logpdf(y) + halfcauchy(scale=5).logpdf(ls)
                + halfcauchy(scale=5).logpdf(amp) + halfnorm(scale=1).logpdf(sg) + th.sum())

    lp, g = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    np.testing.assert_allclose(lp, [ref(th) for th in theta], rtol=1e-11)
    eps = 1e-6
    num = np.array([[(ref(th + eps * e) - ref(th - eps * e)) / (2 * eps) for e in np.eye(3)] for th in theta])
    np.testing.assert_allclose(g, num, rtol=1e-5, atol=1e-5)
    ll = oracle_mod.Oracle(ir).log_likelihood(theta)
    assert ll.shape == (4, 1)
    # forward sampling of the observed vector: loc + chol(K) z -- the sample covariance over many states is K
    Xs, ys = M.gp_data(n=10, d=1)
    irs, _, _ = lower_model(M.gp_model(Xs, ys))
    th = np.tile(np.log([1.3, 0.8, 0.3]), (6000, 1))
    draws = oracle_mod.Oracle(irs).gp_predictive(th, seed=11)
    assert draws.shape == (6000, 10)
    d2 = np.square(Xs - Xs.T)
    Ks = 0.8 ** 2 * np.exp(-d2 / (2 * 1.3 ** 2)) + (0.3 ** 2 + 1e-6) * np.eye(10)
    np.testing.assert_allclose(np.cov(draws.T), Ks, atol=0.06)
    assert np.all(np.abs(draws.mean(0)) < 0.06)
    np.testing.assert_array_equal(draws[:5], oracle_mod.Oracle(irs).gp_predictive(th[:5], seed=11))  # keyed by (state, element)
    # a latent vector with a CONSTANT covariance lowers (quadratic form over the precision matrix as a potential) ...
    @pm.model
    def latent():
        f = yield pm.MvNormal("f", loc=0.0, covariance_matrix=np.eye(3))
        yield pm.Normal("y", f, 1.0, observed=np.zeros(3))

    irl, _, _ = lower_model(latent())
    lpl, _ = oracle_mod.Oracle(irl).logp_grad(np.array([[0.3, -0.2, 0.1]]), dtype=np.float64)
    np.testing.assert_allclose(lpl[0], 2 * stats.norm(0, 1).logpdf([0.3, -0.2, 0.1]).sum(), rtol=1e-12)

    # ... a latent vector whose covariance depends on free variables does not: it fails loudly at lowering time
    @pm.model
    def latent_gp():
        ls = yield pm.HalfNormal("ls", 1.0)
        f = yield pm.MvNormal("f", loc=0.0, covariance_matrix=pm.gp.cov.ExpQuad(length_scale=ls, amplitude=1.0)(Xs, Xs))
        yield pm.Normal("y", f, 1.0, observed=ys)

    with pytest.raises(NotImplementedError, match="marginal likelihood"):
        lower_model(latent_gp())


def test_gp_term_spellings_lower_to_the_same_term(oracle_mod):
    """Variants of the marginal-GP spelling the lowering accepts: noise as a variance (not squared), constant kernel
    parameters, a vector loc, several jitter terms, square() instead of ** 2 -- each checked against scipy."""
    from scipy.stats import halfnorm, multivariate_normal

    X, y = M.gp_data(n=24, d=1)
    n = len(y)
    loc = np.linspace(-0.5, 0.5, n)
    d2 = np.square(X - X.T)

    @pm.model
    def variance_noise():  # noise enters as a variance; the length scale is a constant
        amp = yield pm.HalfNormal("amp", 2.0)
        s2 = yield pm.HalfNormal("s2", 1.0)
        cov = pm.gp.cov.ExpQuad(1.7, amp)(X, X) + s2 * np.eye(n) + 1e-6 * np.eye(n) + 2e-6 * np.eye(n)
        yield pm.MvNormal("y", loc=loc, covariance_matrix=cov, observed=y)

    @pm.model
    def squared_noise():  # pm.math.square spelling, constant amplitude
        ls = yield pm.HalfNormal("ls", 2.0)
        sg = yield pm.HalfNormal("sg", 1.0)
        cov = pm.math.square(sg) * np.eye(n) + pm.gp.cov.ExpQuad(ls, 0.9)(X, X)
        yield pm.MvNormal("y", loc=0.0, covariance_matrix=cov, observed=y)

    ir1, _, _ = lower_model(variance_noise())
    g1 = ir1.gp_terms[0].gp
    par1 = ir1.dataf[g1.par_blob:g1.par_blob + 8]
    assert g1.elems == (-1, 0, 1) and par1[3] == 1.7 and par1[6] == 0.0 and abs(par1[7] - 3e-6) < 1e-18
    ir2, _, _ = lower_model(squared_noise())
    g2 = ir2.gp_terms[0].gp
    par2 = ir2.dataf[g2.par_blob:g2.par_blob + 8]
    assert g2.elems == (0, -1, 1) and par2[4] == 0.9 and par2[6] == 1.0 and par2[7] == 0.0
    th = np.array([[0.2, -1.1], [-0.3, -0.4]])
    lp1, gr1 = oracle_mod.Oracle(ir1).logp_grad(th, dtype=np.float64)
    lp2, gr2 = oracle_mod.Oracle(ir2).logp_grad(th, dtype=np.float64)

    def ref1(t):
        amp, s2 = np.exp(t)
        K = amp ** 2 * np.exp(-d2 / (2 * 1.7 ** 2)) + (s2 + 3e-6) * np.eye(n)
        return (multivariate_normal(loc, K).logpdf(y) + halfnorm(scale=2).logpdf(amp) + halfnorm(scale=1).logpdf(s2) + t.sum())

    def ref2(t):
        ls, sg = np.exp(t)
        K = 0.81 * np.exp(-d2 / (2 * ls ** 2)) + sg ** 2 * np.eye(n)
        return (multivariate_normal(np.zeros(n), K).logpdf(y) + halfnorm(scale=2).logpdf(ls) + halfnorm(scale=1).logpdf(sg) + t.sum())

    np.testing.assert_allclose(lp1, [ref1(t) for t in th], rtol=1e-11)
    np.testing.assert_allclose(lp2, [ref2(t) for t in th], rtol=1e-11)
    eps = 1e-6
    for ref, gr in ((ref1, gr1), (ref2, gr2)):
        num = np.array([[(ref(t + eps * e) - ref(t - eps * e)) / (2 * eps) for e in np.eye(2)] for t in th])
        np.testing.assert_allclose(gr, num, rtol=2e-5, atol=2e-5)
    # forward sampling adds the vector loc back
    draws = oracle_mod.Oracle(ir1).gp_predictive(np.tile(th[0], (4000, 1)), seed=3)
    np.testing.assert_allclose(draws.mean(0), loc, atol=0.12)

    @pm.model
    def two_kernels():
        ls = yield pm.HalfNormal("ls", 2.0)
        cov = pm.gp.cov.ExpQuad(ls, 1.0)(X, X) + pm.gp.cov.ExpQuad(2 * ls, 1.0)(X, X) + 0.1 * np.eye(n)
        yield pm.MvNormal("y", loc=0.0, covariance_matrix=cov, observed=y)

    with pytest.raises(NotImplementedError, match="marginal likelihood"):
        lower_model(two_kernels())


# ------------------------------------------------------------------------------------------
# invariants that pin the restated sampler WITHOUT TensorFlow Probability (the reference holds no golden
# chains and TF/TFP cannot be installed here): properties any correct NUTS / HMC / dual averaging has
# ------------------------------------------------------------------------------------------
def test_nuts_is_stationary_on_the_standard_normal(oracle_mod):
    """NUTS must leave N(0, I) invariant: per-dimension Kolmogorov-Smirnov test of thinned draws, moments, and the
    chi-square law of |x|^2 (a wrong multinomial weight, U-turn test or leapfrog shows up here)."""
    @pm.model
    def m():
        yield pm.Normal("x", 0.0, 1.0, batch_stack=4)

    ir, _, _ = lower_model(m())
    C, B, S = 64, 150, 400
    out = oracle_mod.Oracle(ir).nuts(make_nuts_cfg(C, S, B, step_size=0.3, seed=3,
                                                   adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)),
                                     np.zeros(4), dtype=np.float64)
    x = out["states"][::8]  # thin: NUTS draws on a Gaussian are nearly independent after a few transitions
    flat = x.reshape(-1, 4)
    for d in range(4):
        assert stats.kstest(flat[:, d], "norm").pvalue > 1e-3
    np.testing.assert_allclose(flat.mean(axis=0), 0.0, atol=5.0 / math.sqrt(flat.shape[0]))
    np.testing.assert_allclose(flat.var(axis=0), 1.0, atol=0.08)
    assert stats.kstest((flat ** 2).sum(axis=1), "chi2", args=(4,)).pvalue > 1e-3
    assert out["diverging"].mean() == 0.0
    assert 0.6 < np.exp(out["log_accept"]).mean() < 0.95  # dual averaging steers towards 0.75


def test_hmc_leapfrog_is_reversible_and_volume_preserving(oracle_mod):
    """Run L leapfrogs from (q, p), flip the momentum, run L more: the integrator must come back to q; the energy
    error of a short trajectory must shrink like eps^2 (second-order symplectic integrator)."""
    ir, _, _ = lower_model(M.radon_model(real=True))
    D, L = ir.D, 7
    rng = np.random.default_rng(1)
    q0 = rng.normal(0, 0.1, D)
    p0 = rng.standard_normal(D)
    orc = oracle_mod.Oracle(ir)

    def run(q, p, eps, L=L):
        cfg = make_hmc_cfg(1, 1, 0, step_size=eps, num_leapfrog_steps=L, seed=0, adapt=make_adapt(enabled=False))
        o = orc.hmc(cfg, q, dtype=np.float64, momenta=p[None, None, :], uniforms=np.full((1, 1), 1e-300))
        return o["traj"][0, 0], o["log_accept"][0, 0]

    def final_momentum(q, p, eps):  # restate the integrator for the momentum the ABI does not return
        lp, g = orc.logp_grad(q[None], dtype=np.float64)
        g = g[0]
        p = p + 0.5 * eps * g
        for _ in range(L):
            q = q + eps * p
            g = orc.logp_grad(q[None], dtype=np.float64)[1][0]
            p = p + eps * g
        return q, p - 0.5 * eps * g

    eps = 0.01
    q1, _ = run(q0, p0, eps)
    q1_ref, p1 = final_momentum(q0, p0, eps)
    np.testing.assert_allclose(q1, q1_ref, atol=1e-10)  # the C integrator is the restated one
    q2, _ = run(q1, -p1, eps)
    np.testing.assert_allclose(q2, q0, atol=1e-9)      # time reversal
    # same trajectory length, step halved twice: the energy error of a second-order integrator falls ~4x each time
    errs = [abs(run(q0, p0, e, n)[1]) for e, n in ((0.004, 5), (0.002, 10), (0.001, 20))]
    assert 3.0 < errs[0] / errs[1] < 5.5 and 3.0 < errs[1] / errs[2] < 5.5, errs


def test_sample_chain_step_accounting_and_averaged_step_size(oracle_mod):
    """tfp sample_chain / DualAveragingStepSizeAdaptation accounting: with
    num_adaptation_steps = A the update after transition t = A (0-based) installs exp(log_avg); every transition is
    traced when burn_in = 0, so the recursion can be replayed from the traced acceptance ratios."""
    @pm.model
    def m():
        yield pm.Normal("x", 0.0, 1.0, batch_stack=5)

    ir, _, _ = lower_model(m())
    A, S = 12, 20
    cfg = make_nuts_cfg(1, S, 0, step_size=0.1, seed=9, adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=A))
    out = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(5), dtype=np.float64)
    acc = np.exp(out["log_accept"][:, 0])
    assert out["log_accept"].shape == (S, 1)
    eps0, err, log_avg, eps = 0.1, 0.0, 0.0, 0.1
    for t in range(S):
        step = t + 1.0
        err += 0.75 - acc[t]
        log_eps = math.log(10 * eps0) - err * math.sqrt(step) / ((10 + step) * 0.05)
        eta = step ** -0.75
        log_avg = eta * log_eps + (1 - eta) * log_avg
        if t < A:
            eps = math.exp(log_eps)
        elif t == A:
            eps = math.exp(log_avg)
    np.testing.assert_allclose(out["step_size"][0], eps, rtol=1e-10)
    assert out["total_leapfrogs"][0] == out["leapfrogs"].sum()  # every transition of the call is accounted for


def test_diagonal_mass_matrix_windows_recover_the_scales(oracle_mod):
    """pm4_adapt_cfg_t.mass_windows: on a Gaussian whose coordinates differ by a factor of 400 in
    scale the adapted step scale is proportional to the standard deviations and the trees shrink from hundreds of
    leapfrogs to a handful; the draws are still N(0, diag(sigma^2))."""
    sig = np.array([0.05, 0.3, 1.0, 4.0, 20.0])

    @pm.model
    def m():
        yield pm.Normal("x", np.zeros(5), sig)

    ir, _, _ = lower_model(m())
    res = {}
    for K in (0, 4):
        cfg = make_nuts_cfg(64, 200, 300, step_size=0.1, seed=1,
                            adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=300, mass_windows=K))
        res[K] = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(5), dtype=np.float64)
    assert np.all(res[0]["mass_scale"] == 1.0)
    s = res[4]["mass_scale"]
    np.testing.assert_allclose(np.exp(np.log(s).mean()), 1.0, rtol=1e-12)  # geometric mean 1
    np.testing.assert_allclose(s / s[2], sig / sig[2], rtol=0.25)
    assert res[4]["leapfrogs"].mean() < 0.1 * res[0]["leapfrogs"].mean()
    x = res[4]["states"].reshape(-1, 5)
    np.testing.assert_allclose(x.std(axis=0), sig, rtol=0.05)
    assert res[4]["diverging"].mean() == 0.0


def test_multivariate_normal_with_constant_covariance_against_scipy(oracle_mod):
    """MvNormalCholesky as a LATENT vector and MvNormal as an observation with a free loc (row a16: reference
    distributions/multivariate.py:159-199, 331-375): the constant precision matrix turns the quadratic form into an
    element-wise potential; value and gradient against scipy / finite differences."""
    cov, y = M.mvn_latent_data()
    k = cov.shape[0]
    ir, _, _ = lower_model(M.mvn_latent_model(cov, y))
    assert ir.D == k + 1
    rng = np.random.default_rng(3)
    th = rng.normal(0, 0.5, (6, ir.D))
    lp, g = oracle_mod.Oracle(ir).logp_grad(th, dtype=np.float64)

    def ref(t):
        f, z = t[:k], t[k]
        return (stats.multivariate_normal(np.zeros(k), cov).logpdf(f) + stats.halfnorm(scale=1).logpdf(np.exp(z)) + z
                + stats.norm(f, np.exp(z)).logpdf(y).sum())

    np.testing.assert_allclose(lp, [ref(t) for t in th], rtol=1e-12)
    for c in range(2):
        for d in range(ir.D):
            e = np.zeros(ir.D); e[d] = 1e-6
            np.testing.assert_allclose(g[c, d], (ref(th[c] + e) - ref(th[c] - e)) / 2e-6, rtol=1e-5, atol=1e-6)

    @pm.model
    def obs_model():
        mu = yield pm.Normal("mu", 0.0, 1.0)
        yield pm.MvNormal("y", loc=mu, covariance_matrix=cov, observed=y)

    ir2, _, _ = lower_model(obs_model())
    t2 = rng.normal(0, 1, (4, 1))
    lp2, _ = oracle_mod.Oracle(ir2).logp_grad(t2, dtype=np.float64)
    np.testing.assert_allclose(lp2, [stats.norm(0, 1).logpdf(t[0]) + stats.multivariate_normal(np.full(k, t[0]), cov).logpdf(y)
                                     for t in t2], rtol=1e-12)
    d = pm.MvNormalCholesky("v", loc=np.zeros(k), scale_tril=np.linalg.cholesky(cov))
    np.testing.assert_allclose(d.log_prob(y), stats.multivariate_normal(np.zeros(k), cov).logpdf(y), rtol=1e-10)
