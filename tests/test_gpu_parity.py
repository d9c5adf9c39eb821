"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states:
  (a) per-state log_prob and gradients within 1e-5 relative in FP32, 1e-12 in FP64;
  (b) fixed-L HMC trajectories with injected identical momenta within 1e-4 after 100 steps;
  (c) NUTS posterior means / variances within 3 Monte Carlo standard errors, R-hat < 1.01;
  (d) integer tree-depth and divergence counts bit-exact on the injected-RNG replay.
"""
import numpy as np
import pytest

from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED, make_adapt, make_hmc_cfg, make_nuts_cfg
from pymc4_b200.ir import lower_model
from tests import models as M

pytestmark = pytest.mark.gpu


def _dm(ir, f64=False):
    from pymc4_b200._cabi import DeviceModel

    return DeviceModel(ir, device=0, f64=f64)


MODELS = {
    "eight_schools": lambda: M.schools_pm4(),
    "radon_real": lambda: M.radon_model(real=True),
    "radon_synth": lambda: M.radon_model(real=False),
    "logistic": lambda: M.logistic_model(*M.logistic_data()),
    "poisson": lambda: M.poisson_model(*M.poisson_data()),
    "nested": lambda: M.outer_model(),
    "mvn_latent": lambda: M.mvn_latent_model(),
}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


@pytest.mark.parametrize("name", list(MODELS))
def test_logp_grad_parity_fp32(name, oracle_mod):
    ir, _, _ = lower_model(MODELS[name]())
    rng = np.random.default_rng(11)
    theta = rng.normal(0, 0.4, (96, ir.D))
    theta[0] = 0.0  # the test point every chain starts from
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    lp, g = _dm(ir).logp_grad(theta.astype(np.float32), dtype=np.float32)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    assert np.all(np.isfinite(lp))
    # (a) 1e-5 relative, per state: log-density against its own magnitude, gradient norm-wise
    assert np.max(np.abs(lp - lp_ref) / np.maximum(np.abs(lp_ref), 1.0)) < 1e-5
    for c in range(theta.shape[0]):
        assert _rel(g[c], g_ref[c]) < 1e-5, (name, c)


@pytest.mark.parametrize("name", ["eight_schools", "radon_real", "logistic", "poisson", "mvn_latent"])
def test_logp_grad_parity_fp64(name, oracle_mod):
    ir, _, _ = lower_model(MODELS[name]())
    rng = np.random.default_rng(12)
    theta = rng.normal(0, 0.4, (40, ir.D))
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    lp, g = _dm(ir, f64=True).logp_grad(theta, dtype=np.float64)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    assert np.max(np.abs(lp - lp_ref) / np.maximum(np.abs(lp_ref), 1.0)) < 1e-12
    for c in range(theta.shape[0]):
        assert _rel(g[c], g_ref[c]) < 1e-12, (name, c)


def test_golden_logp_values(golden):
    # the reference's own golden value (tests/test_8schools.py:61) and the radon.csv values
    ir, _, _ = lower_model(M.schools_pm4())
    th = ir.join_theta({"schools_pm4/eta": np.zeros(8), "schools_pm4/mu": 0.0, "schools_pm4/__log_tau": 1.0})
    lp64, _ = _dm(ir, f64=True).logp_grad(th[None], dtype=np.float64)
    assert abs(float(lp64[0]) - golden["eight_schools"]["scipy_f64"]) < 1e-11
    assert np.float32(float(lp64[0])) == np.float32(golden["eight_schools"]["reference_assert_f32"])
    lp32, _ = _dm(ir).logp_grad(th[None].astype(np.float32), dtype=np.float32)
    np.testing.assert_allclose(float(lp32[0]), golden["eight_schools"]["reference_assert_f32"], rtol=1e-6)

    ir, _, _ = lower_model(M.radon_model(real=True))
    dm = _dm(ir, f64=True)
    D = ir.D
    pts = np.stack([np.zeros(D), np.array(golden["radon_919"]["theta_rand"])])
    lp, _ = dm.logp_grad(pts, dtype=np.float64)
    np.testing.assert_allclose(lp.cpu().numpy(), [golden["radon_919"]["logp_at_zero"], golden["radon_919"]["logp_at_rand"]],
                               rtol=1e-12)


@pytest.mark.parametrize("name", ["eight_schools", "radon_synth"])
def test_hmc_trajectory_parity(name, oracle_mod):
    """(b) one 100-leapfrog trajectory per chain from identical injected momenta."""
    ir, _, _ = lower_model(MODELS[name]())
    C, D, L = 16, ir.D, 100
    rng = np.random.default_rng(5)
    theta0 = rng.normal(0, 0.2, (C, D))
    momenta = rng.standard_normal((1, C, D))
    uniforms = rng.random((1, C))
    eps = 0.005 if name == "radon_synth" else 0.05  # inside the stable region of the integrator
    cfg = make_hmc_cfg(C, 1, 0, step_size=eps, num_leapfrog_steps=L, seed=1, adapt=make_adapt(enabled=False))
    ref = oracle_mod.Oracle(ir).hmc(cfg, theta0, dtype=np.float64, momenta=momenta, uniforms=uniforms)
    out = _dm(ir).hmc(cfg, theta0.astype(np.float32), dtype=np.float32, momenta=momenta, uniforms=uniforms, want_traj=True)
    traj = out["traj"].cpu().numpy()
    assert np.max(np.abs(traj - ref["traj"])) < 1e-4
    np.testing.assert_allclose(out["log_accept"].cpu().numpy(), ref["log_accept"], atol=5e-3, rtol=1e-4)
    # and the accept decisions + kept states agree
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
    assert np.max(np.abs(out["states"].cpu().numpy() - ref["states"])) < 1e-4


def test_hmc_chain_parity_with_adaptation(oracle_mod):
    """A short adaptive HMC chain (pooled and per-chain dual averaging) replayed from injected randomness."""
    ir, _, _ = lower_model(M.schools_pm4())
    C, D, B, S = 8, ir.D, 15, 10
    rng = np.random.default_rng(7)
    momenta = rng.standard_normal((B + S, C, D))
    uniforms = rng.random((B + S, C))
    for mode in (PM4_EPS_POOLED, PM4_EPS_PER_CHAIN):
        cfg = make_hmc_cfg(C, S, B, step_size=0.1, num_leapfrog_steps=3, seed=3,
                           adapt=make_adapt(mode=mode, num_adaptation_steps=B))
        ref = oracle_mod.Oracle(ir).hmc(cfg, np.zeros(D), dtype=np.float64, momenta=momenta, uniforms=uniforms)
        out = _dm(ir, f64=True).hmc(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta, uniforms=uniforms)
        assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
        np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=1e-9)
        np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-9)


def test_per_chain_initial_step_sizes_and_shrinkage_target_replay(oracle_mod):
    """`step_size` with a leading chain axis whose entries differ (cfg.step_size_dev) and tfp's `shrinkage_target`: FP64
    replay of NUTS and HMC against the oracle -- integer statistics bit-exact, adapted step sizes equal."""
    ir, _, _ = lower_model(M.schools_pm4())
    C, D, B, S = 10, ir.D, 20, 15
    rng = np.random.default_rng(33)
    momenta = rng.standard_normal((B + S, C, D))
    eps_c = np.linspace(0.05, 0.4, C)
    adapt = make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B, shrinkage_target=0.7)
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=4, adapt=adapt)
    ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float64, momenta=momenta, step_sizes=eps_c)
    out = _dm(ir, f64=True).nuts(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta, step_sizes=eps_c)
    for key in ("leapfrogs", "depth", "diverging"):
        assert np.array_equal(out[key].cpu().numpy(), ref[key]), key
    np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-7)
    assert np.ptp(ref["leapfrogs"][0]) > 0  # the chains really started from different step sizes
    # the default recursion (10 x the initial step) ends elsewhere
    dflt = oracle_mod.Oracle(ir).nuts(make_nuts_cfg(C, S, B, step_size=0.1, seed=4, adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B)),
                                      np.zeros(D), dtype=np.float64, momenta=momenta, step_sizes=eps_c)
    assert not np.allclose(dflt["step_size"], ref["step_size"])
    uniforms = rng.random((B + S, C))
    hcfg = make_hmc_cfg(C, S, B, step_size=0.1, num_leapfrog_steps=3, seed=3, adapt=adapt)
    ref = oracle_mod.Oracle(ir).hmc(hcfg, np.zeros(D), dtype=np.float64, momenta=momenta, uniforms=uniforms, step_sizes=eps_c)
    out = _dm(ir, f64=True).hmc(hcfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta, uniforms=uniforms, step_sizes=eps_c)
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=1e-9)
    np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-7)


@pytest.mark.parametrize("name,mode", [("eight_schools", PM4_EPS_POOLED), ("eight_schools", PM4_EPS_PER_CHAIN),
                                       ("radon_synth", PM4_EPS_PER_CHAIN)])
def test_nuts_replay_bit_exact_integers(name, mode, oracle_mod):
    """(d) FP64 replay with injected momenta: tree depth, leapfrog and divergence counts are bit-exact."""
    ir, _, _ = lower_model(MODELS[name]())
    C, D = 12, ir.D
    B, S = (25, 25) if name == "eight_schools" else (12, 10)
    rng = np.random.default_rng(21)
    momenta = rng.standard_normal((B + S, C, D))
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=99, adapt=make_adapt(mode=mode, num_adaptation_steps=B))
    ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float64, momenta=momenta)
    out = _dm(ir, f64=True).nuts(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta)
    for key in ("leapfrogs", "depth", "diverging"):
        assert np.array_equal(out[key].cpu().numpy(), ref[key]), key
    assert np.array_equal(out["total_leapfrogs"].cpu().numpy(), ref["total_leapfrogs"])
    # real-valued outputs: rounding differences (device FMA contraction, libm vs CUDA math) grow by a
    # constant factor per transition through the Hamiltonian dynamics; after B+S transitions they are
    # still far below anything that flips an integer decision
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=5e-3)
    np.testing.assert_allclose(out["energy"].cpu().numpy(), ref["energy"], rtol=1e-4, atol=5e-3)
    np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-4)
    # the first kept draws of a short non-adaptive run agree to round-off
    cfg0 = make_nuts_cfg(C, 4, 0, step_size=0.05, seed=99, adapt=make_adapt(enabled=False))
    ref0 = oracle_mod.Oracle(ir).nuts(cfg0, np.zeros(D), dtype=np.float64, momenta=momenta[:4])
    out0 = _dm(ir, f64=True).nuts(cfg0, np.zeros((C, D)), dtype=np.float64, momenta=momenta[:4])
    assert np.array_equal(out0["leapfrogs"].cpu().numpy(), ref0["leapfrogs"])
    np.testing.assert_allclose(out0["states"].cpu().numpy(), ref0["states"], atol=1e-10)


def test_philox_momentum_stream_matches_host(oracle_mod):
    """Device and host draw the same Philox momenta (FP64 Box-Muller, 1 ulp-level agreement)."""
    ir, _, _ = lower_model(M.radon_model(real=False))
    C, D = 4, ir.D
    cfg = make_hmc_cfg(C, 1, 0, step_size=1e-9, num_leapfrog_steps=1, seed=1234, adapt=make_adapt(enabled=False),
                       chain_offset=5)
    # with a vanishing step the proposal equals the start: recover p from the kinetic-energy difference is
    # not possible, so compare through the trajectory of a *unit* step on a flat direction instead:
    ref = oracle_mod.Oracle(ir).hmc(cfg, np.zeros(D), dtype=np.float64)
    out = _dm(ir, f64=True).hmc(cfg, np.zeros((C, D)), dtype=np.float64, want_traj=True)
    # theta' = theta + eps * (p + eps/2 * grad)  ->  identical momenta give identical proposals
    np.testing.assert_allclose(out["traj"].cpu().numpy(), ref["traj"], rtol=0, atol=1e-20)
    for c in range(C):
        host = oracle_mod.draw_momentum(1234, 5 + c, 0, D, dtype=np.float64)
        assert np.all(np.isfinite(host)) and abs(host.mean()) < 0.5


def _split_rhat(x):
    """x: [chain, draw] -> classic split R-hat."""
    C, S = x.shape
    h = S // 2
    parts = np.concatenate([x[:, :h], x[:, h:2 * h]], axis=0)
    m = parts.mean(axis=1)
    w = parts.var(axis=1, ddof=1).mean()
    b = h * m.var(ddof=1)
    return np.sqrt(((h - 1) / h * w + b / h) / w)


def test_nuts_posterior_matches_oracle(oracle_mod):
    """(c) GPU FP32 NUTS (Philox stream) vs the oracle's FP64 NUTS on eight schools."""
    ir, _, _ = lower_model(M.schools_pm4())
    D = ir.D
    B, S = 300, 1000
    Cg, Co = 512, 64
    adapt = make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B)
    out = _dm(ir).nuts(make_nuts_cfg(Cg, S, B, step_size=0.28, seed=2026, adapt=adapt), np.zeros((Cg, D), np.float32),
                       dtype=np.float32)
    ref = oracle_mod.Oracle(ir).nuts(make_nuts_cfg(Co, S, B, step_size=0.28, seed=77, adapt=adapt), np.zeros(D),
                                     dtype=np.float64)
    xs = out["states"].cpu().numpy().astype(np.float64)  # [S, C, D]
    xr = ref["states"]
    assert not np.any(np.isnan(xs))
    for fn in (lambda v: v, lambda v: v * v):  # means and second moments (-> variances)
        mg, mo = fn(xs).mean(axis=0), fn(xr).mean(axis=0)  # per-chain means [C, D]
        se = np.sqrt(mg.var(axis=0, ddof=1) / Cg + mo.var(axis=0, ddof=1) / Co)
        z = np.abs(mg.mean(axis=0) - mo.mean(axis=0)) / se
        assert np.all(z < 3.0), z
    for d in range(D):
        assert _split_rhat(xs[:, :, d].T) < 1.01
    # sampler health mirrors the oracle's
    assert abs(np.exp(out["log_accept"].cpu().numpy()).mean() - np.exp(ref["log_accept"]).mean()) < 0.03
    assert out["diverging"].float().mean().item() < 0.02


def test_radon_pooled_step_size_matches_the_oracle_regime(oracle_mod):
    """(c) on the HEADLINE configuration: radon-919, the reference-default scalar step size (one epsilon adapted from
    the mean acceptance of all chains), GPU FP32 (2048 chains) against the oracle's FP64 NUTS (128 chains).

    The reference-default policy does NOT converge on this model (one shared epsilon sticks ~1 % of the chains in the
    sigma_beta funnel: split R-hat ~1.2-1.3, ~9 % divergent transitions -- the oracle says the same, and the
    reference's own notebook works around it with per-dimension step sizes, radon_hierarchical.ipynb:123-146), so this
    test compares the REGIME, chain means as the independent unit: posterior means and second moments of the seven
    monitored columns within 3 standard errors, divergent fraction, mean acceptance, leapfrogs per draw, adapted step
    size and the fraction of chains that diverge on (almost) every transition."""
    ir, _, _ = lower_model(M.radon_model(real=False))
    D = ir.D
    B, S = 300, 300
    Cg, Co = 2048, 128
    adapt = make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)
    out = _dm(ir).nuts(make_nuts_cfg(Cg, S, B, step_size=0.1, seed=2026, adapt=adapt), np.zeros((Cg, D), np.float32),
                       dtype=np.float32)
    ref = oracle_mod.Oracle(ir).nuts(make_nuts_cfg(Co, S, B, step_size=0.1, seed=77, adapt=adapt), np.zeros(D),
                                     dtype=np.float64)
    cols = [0, 1, 2, 87, 172, 173, 174]  # mu_a, mu_b, alpha[0], beta[0], log sigma_a, log sigma_b, log eps
    xs = out["states"].cpu().numpy().astype(np.float64)[:, :, cols]
    xr = ref["states"][:, :, cols]
    assert np.all(np.isfinite(xs))
    for fn in (lambda v: v, lambda v: v * v):
        mg, mo = fn(xs).mean(axis=0), fn(xr).mean(axis=0)  # per-chain means [C, 7]
        se = np.sqrt(mg.var(axis=0, ddof=1) / Cg + mo.var(axis=0, ddof=1) / Co)
        z = np.abs(mg.mean(axis=0) - mo.mean(axis=0)) / se
        assert np.all(z < 3.0), z
    dg, do = out["diverging"].cpu().numpy().astype(np.float64), ref["diverging"].astype(np.float64)
    # per-chain divergence rates are the independent unit (a stuck chain diverges on every transition)
    se_div = np.sqrt(dg.mean(axis=0).var(ddof=1) / Cg + do.mean(axis=0).var(ddof=1) / Co)
    assert abs(dg.mean() - do.mean()) < 3.0 * se_div + 0.01, (dg.mean(), do.mean())
    stuck_g, stuck_o = (dg.mean(axis=0) > 0.9).mean(), (do.mean(axis=0) > 0.9).mean()
    assert abs(stuck_g - stuck_o) < 3.0 * np.sqrt(max(stuck_o, 1.0 / Co) / Co) + 0.01, (stuck_g, stuck_o)
    ag, ao = np.exp(out["log_accept"].cpu().numpy()).mean(), np.exp(ref["log_accept"]).mean()
    assert abs(ag - ao) < 0.03, (ag, ao)
    lg, lo = out["leapfrogs"].float().mean().item(), ref["leapfrogs"].mean()
    assert abs(lg - lo) / lo < 0.08, (lg, lo)
    eg, eo = float(out["step_size"][0]), float(ref["step_size"][0])
    assert abs(eg - eo) / eo < 0.15, (eg, eo)


def test_radon_noncentered_meets_check_c_in_full(oracle_mod):
    """(c) in FULL on a configuration that converges: the radon-919 data with the group effects written as offsets
    (benchmodels.hierarchical_model_noncentered), reference-default sampler settings (scalar step size 0.1, pooled dual
    averaging over the burn-in).  GPU FP32 (1024 chains) against the oracle's FP64 NUTS (96 chains): posterior means and
    variances of the seven monitored columns within 3 Monte Carlo standard errors, split R-hat < 1.01 on every one of
    them, divergent transitions equal to the oracle's (2.6 % after this short burn-in, 0.4 % after the bench's 1000), the same adapted step size and tree sizes."""
    import torch

    from pymc4_b200 import benchmodels as BM, diagnostics

    idx, x, y, g = BM.radon_synthetic(919, 85)
    ir, _, _ = lower_model(BM.hierarchical_model_noncentered(idx, x, y, g))
    D = ir.D
    B, S = 400, 400
    Cg, Co = 1024, 96
    adapt = make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)
    out = _dm(ir).nuts(make_nuts_cfg(Cg, S, B, step_size=0.1, seed=41, adapt=adapt), np.zeros((Cg, D), np.float32),
                       dtype=np.float32)
    ref = oracle_mod.Oracle(ir).nuts(make_nuts_cfg(Co, S, B, step_size=0.1, seed=43, adapt=adapt), np.zeros(D),
                                     dtype=np.float64)
    cols = [0, 1, 2, 87, 172, 173, 174]  # mu_a, mu_b, alpha_raw[0], beta_raw[0], log sigma_a, log sigma_b, log eps
    summ = diagnostics.summarize(out["states"], columns=cols, gather=False)
    assert summ["max_rhat"] < 1.01, summ["rhat"]
    dg, do = float(out["diverging"].float().mean()), float(ref["diverging"].mean())
    assert dg < 0.04 and abs(dg - do) < 0.01, (dg, do)  # 2.6 % after this short burn-in; 0.4 % after the bench's 1000 transitions
    summ_ref = diagnostics.summarize(torch.as_tensor(ref["states"], dtype=torch.float32), columns=cols, gather=False)
    assert summ_ref["max_rhat"] < 1.02, summ_ref["rhat"]  # 96 chains x 400 draws: the oracle agrees that it converges
    xs = out["states"].cpu().numpy().astype(np.float64)[:, :, cols]
    xr = ref["states"][:, :, cols]
    # means: chain means are the independent unit; variances: per-chain variances likewise
    for stat in (lambda v: v.mean(axis=0), lambda v: v.var(axis=0, ddof=1)):
        sg, so = stat(xs), stat(xr)  # [C, 7]
        se = np.sqrt(sg.var(axis=0, ddof=1) / Cg + so.var(axis=0, ddof=1) / Co)
        z = np.abs(sg.mean(axis=0) - so.mean(axis=0)) / se
        assert np.all(z < 3.0), z
    eg, eo = float(out["step_size"][0]), float(ref["step_size"][0])
    assert abs(eg - eo) / eo < 0.1, (eg, eo)
    lg, lo = out["leapfrogs"].float().mean().item(), ref["leapfrogs"].mean()
    assert abs(lg - lo) / lo < 0.08, (lg, lo)


@pytest.mark.parametrize("mode", [PM4_EPS_POOLED, PM4_EPS_PER_CHAIN])
def test_mass_matrix_windows_replay_against_the_oracle(mode, oracle_mod):
    """Diagonal mass-matrix windows (cross-chain spread -> step scale, dual averaging restarted): FP64 replay with
    injected momenta against the oracle -- integer tree statistics bit-exact, the adapted scale and step size equal."""
    ir, _, _ = lower_model(M.schools_pm4())
    # short on purpose: the windows couple ALL chains through the scale, so one rounding-flipped comparison in one
    # chain eventually reaches every chain (at B = 40 one of 288 tree sizes differs, the step size by 8e-7)
    C, D, B, S = 24, ir.D, 12, 12
    rng = np.random.default_rng(8)
    momenta = rng.standard_normal((B + S, C, D))
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=3, adapt=make_adapt(mode=mode, num_adaptation_steps=B, mass_windows=2))
    ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float64, momenta=momenta)
    out = _dm(ir, f64=True).nuts(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta)
    for key in ("leapfrogs", "depth", "diverging"):
        assert np.array_equal(out[key].cpu().numpy(), ref[key]), key
    np.testing.assert_allclose(out["mass_scale"].cpu().numpy(), ref["mass_scale"], rtol=1e-6)
    np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-5)
    assert np.ptp(ref["mass_scale"]) > 0.1  # the windows did move the scale
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=5e-3)


def test_mass_matrix_windows_shrink_the_radon_funnel_tail():
    """radon-919 under the reference-default shared step size leaves ~9 % divergent transitions and ~2 % of the chains
    diverging on (almost) every transition; with diagonal mass-matrix windows (the backend's counterpart of the
    notebook's hand-fed per-dimension step sizes, radon_hierarchical.ipynb:123-146) the same run diverges less than
    half as often, almost no chain stays stuck, the adapted step size is larger and the trees are shorter.  (The
    centred parameterisation still mixes slowly in log sigma_beta: split R-hat stays ~1.1 at 1000 draws either way.)"""
    ir, _, _ = lower_model(M.radon_model(real=False))
    C, B, S = 1024, 500, 400
    o = {}
    for K in (0, 4):
        cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=12, adapt=make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B, mass_windows=K))
        o[K] = _dm(ir).nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    div = {K: o[K]["diverging"].float() for K in o}
    d0, d4 = float(div[0].mean()), float(div[4].mean())
    s0, s4 = (float((div[K].mean(dim=0) > 0.9).float().mean()) for K in (0, 4))
    print("divergent %.4f -> %.4f, stuck chains %.4f -> %.4f, step size %.4f -> %.4f, leapfrogs/draw %.1f -> %.1f"
          % (d0, d4, s0, s4, float(o[0]["step_size"][0]), float(o[4]["step_size"][0]),
             float(o[0]["leapfrogs"].float().mean()), float(o[4]["leapfrogs"].float().mean())))
    assert d4 < 0.6 * d0 and s4 < 0.5 * s0 + 1e-3
    assert float(o[4]["step_size"][0]) > 1.2 * float(o[0]["step_size"][0])
    scale = o[4]["mass_scale"].cpu().numpy()
    np.testing.assert_allclose(np.exp(np.log(scale).mean()), 1.0, rtol=1e-4)
    assert scale[174] < 0.3 and scale[173] > 1.5  # log eps is pinned down by 919 observations, log sigma_beta is wide


def test_radon_fp32_replay_reports_exact_match_rate(oracle_mod):
    """(d) in FP32 on radon-919 with injected momenta: the integer tree statistics replay bit-exactly until a rounding
    difference flips a comparison (device FMA contraction, MUFU vs libm, another summation order); from there a chain
    runs its own valid trajectory.  The first transitions must match for (almost) every chain and the match rate is
    reported."""
    ir, _, _ = lower_model(M.radon_model(real=False))
    C, D, T = 64, ir.D, 6
    rng = np.random.default_rng(5)
    momenta = rng.standard_normal((T, C, D))
    cfg = make_nuts_cfg(C, T, 0, step_size=0.02, seed=4, adapt=make_adapt(enabled=False))
    ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float32, momenta=momenta)
    out = _dm(ir).nuts(cfg, np.zeros((C, D), np.float32), dtype=np.float32, momenta=momenta.astype(np.float32))
    same = out["leapfrogs"].cpu().numpy() == ref["leapfrogs"]  # [T, C]
    alive = np.cumprod(same, axis=0).astype(bool)  # still on the oracle's trajectory
    rate = alive.mean(axis=1)
    print("FP32 replay, radon-919: chains on the oracle's trajectory after transition 1..%d: %s" % (T, np.round(rate, 3)))
    assert rate[0] >= 0.95 and rate[1] >= 0.85, rate
    first = alive[0]
    np.testing.assert_allclose(out["states"].cpu().numpy()[0][first], ref["states"][0][first], atol=5e-4)
    assert np.array_equal(out["diverging"].cpu().numpy()[0][first], ref["diverging"][0][first])


# ------------------------------------------------------------------------------------------
# large-state path (BASELINE configs[4]: radon-shaped, 100k observations / 5k groups, D = 10 005):
# 16 warps per chain, 20 state elements per lane, hybrid tree scratch (hot slots on chip, rest in L2),
# plan pools read from global memory
# ------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def large_ir():
    idx, x, y, g = M.radon_large()
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, g))
    assert ir.D == 10005 and ir.team_warps == 16 and ir.part_k == 1
    return ir


def test_large_state_logp_grad_parity(large_ir, oracle_mod):
    ir = large_ir
    rng = np.random.default_rng(31)
    theta = rng.normal(0, 0.3, (6, ir.D))
    theta[0] = 0.0
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    lp, g = _dm(ir).logp_grad(theta.astype(np.float32), dtype=np.float32)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    assert np.max(np.abs(lp - lp_ref) / np.abs(lp_ref)) < 1e-5
    for c in range(theta.shape[0]):
        assert _rel(g[c], g_ref[c]) < 1e-5, c


def test_large_state_hmc_trajectory(large_ir, oracle_mod):
    ir = large_ir
    C, D, L = 3, ir.D, 50
    rng = np.random.default_rng(32)
    theta0 = rng.normal(0, 0.1, (C, D))
    momenta = rng.standard_normal((1, C, D))
    uniforms = rng.random((1, C))
    # the posterior of 100k observations is tight: the integrator is stable for eps well below 1e-3
    cfg = make_hmc_cfg(C, 1, 0, step_size=2e-4, num_leapfrog_steps=L, seed=1, adapt=make_adapt(enabled=False))
    ref = oracle_mod.Oracle(ir).hmc(cfg, theta0, dtype=np.float64, momenta=momenta, uniforms=uniforms)
    assert np.max(np.abs(ref["traj"])) < 10
    out = _dm(ir).hmc(cfg, theta0.astype(np.float32), dtype=np.float32, momenta=momenta, uniforms=uniforms, want_traj=True)
    assert np.max(np.abs(out["traj"].cpu().numpy() - ref["traj"])) < 1e-4
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])


def test_large_state_nuts_against_oracle(large_ir, oracle_mod):
    """A short FP32 NUTS run with injected momenta: tree sizes agree with the oracle's FP32 run wherever
    rounding does not flip a U-turn test (the bit-exact claim (d) is made in FP64 on the small models)."""
    ir = large_ir
    C, D, B, S = 4, ir.D, 3, 3
    rng = np.random.default_rng(33)
    momenta = rng.standard_normal((B + S, C, D))
    cfg = make_nuts_cfg(C, S, B, step_size=3e-4, seed=5, max_tree_depth=6,
                        adapt=make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B))
    ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float64, momenta=momenta)
    out = _dm(ir).nuts(cfg, np.zeros((C, D), dtype=np.float32), dtype=np.float32, momenta=momenta)
    got, want = out["leapfrogs"].cpu().numpy(), ref["leapfrogs"]
    assert np.mean(got == want) >= 0.75, (got, want)
    same = got == want
    assert np.max(np.abs(out["states"].cpu().numpy()[same] - ref["states"][same])) < 5e-2
    assert np.all(np.isfinite(out["states"].cpu().numpy()))


@pytest.mark.parametrize("name", ["eight_schools", "radon_synth", "logistic", "poisson"])
def test_pointwise_log_likelihood_parity(name, oracle_mod):
    """dist.log_prob(observed) per state on the device against the oracle, FP32 and FP64."""
    ir, _, _ = lower_model(MODELS[name]())
    rng = np.random.default_rng(41)
    theta = rng.normal(0, 0.4, (7, 5, ir.D))
    ref = oracle_mod.Oracle(ir).log_likelihood(theta, dtype=np.float64)
    got = _dm(ir).log_likelihood(theta.astype(np.float32), dtype=np.float32).cpu().numpy()
    assert got.shape == ref.shape == (7, 5, ir.ll_row) and ir.ll_row > 0
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) < 1e-5
    if name != "radon_synth":
        got64 = _dm(ir, f64=True).log_likelihood(theta, dtype=np.float64).cpu().numpy()
        np.testing.assert_allclose(got64, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["eight_schools", "radon_synth", "logistic", "poisson"])
def test_posterior_predictive_matches_oracle_stream(name, oracle_mod):
    """device draws of the observed variables against the oracle on the same Philox stream:
    continuous draws agree to rounding, integer draws are identical except where rounding flips a comparison."""
    ir, _, _ = lower_model(MODELS[name]())
    rng = np.random.default_rng(43)
    theta = rng.normal(0, 0.4, (6, 50, ir.D))
    if name == "poisson":
        theta[3:, :, ir.rv_by_name("poisson_model/intercept").offset] += 3.0  # rates above 10: the PTRS branch
    ref = oracle_mod.Oracle(ir).posterior_predictive(theta, seed=77, dtype=np.float64)
    # FP32 draws use 24-bit uniforms, FP64 draws 32/53-bit ones: compare each precision with its own oracle
    ref32 = oracle_mod.Oracle(ir).posterior_predictive(theta.astype(np.float32), seed=77, dtype=np.float32)
    got = _dm(ir).posterior_predictive(theta.astype(np.float32), seed=77, dtype=np.float32).cpu().numpy()
    assert got.shape == ref.shape
    if name in ("eight_schools", "radon_synth"):
        assert np.max(np.abs(got - ref32) / np.maximum(np.abs(ref32), 1.0)) < 2e-5
    else:
        assert np.mean(got == ref32) > 0.995 and np.all(got == np.round(got)) and got.min() >= 0
    if name != "radon_synth":
        got64 = _dm(ir, f64=True).posterior_predictive(theta, seed=77, dtype=np.float64).cpu().numpy()
        if name == "eight_schools":
            np.testing.assert_allclose(got64, ref, rtol=1e-11, atol=1e-11)
        else:
            assert np.mean(got64 == ref) > 0.9999


def test_simple_adaptation_and_random_walk_replay(oracle_mod):
    """hmc_simple / nuts_simple (tfp SimpleStepSizeAdaptation) and rwm (RandomWalkMetropolis),
    FP64 replay from injected randomness against the oracle."""
    from pymc4_b200._cstructs import PM4_ADAPT_SIMPLE, PM4_PROPOSAL_RANDOM_WALK

    ir, _, _ = lower_model(M.schools_pm4())
    C, D, B, S = 8, ir.D, 20, 10
    rng = np.random.default_rng(51)
    momenta = rng.standard_normal((B + S, C, D))
    uniforms = rng.random((B + S, C))
    orc, dm = oracle_mod.Oracle(ir), _dm(ir, f64=True)
    for mode in (PM4_EPS_POOLED, PM4_EPS_PER_CHAIN):
        adapt = make_adapt(mode=mode, num_adaptation_steps=B, kind=PM4_ADAPT_SIMPLE, adaptation_rate=0.05)
        cfg = make_hmc_cfg(C, S, B, step_size=0.1, num_leapfrog_steps=3, seed=3, adapt=adapt)
        ref = orc.hmc(cfg, np.zeros(D), dtype=np.float64, momenta=momenta, uniforms=uniforms)
        out = dm.hmc(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta, uniforms=uniforms)
        assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
        np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=1e-9)
        np.testing.assert_allclose(out["step_size"].cpu().numpy(), ref["step_size"], rtol=1e-12)
        assert len(np.unique(ref["step_size"])) == (1 if mode == PM4_EPS_POOLED else len(np.unique(ref["step_size"])))
        # the step size moved by whole powers of (1 + rate)
        k = np.log(ref["step_size"] / 0.1) / np.log(1.05)
        np.testing.assert_allclose(k, np.round(k), atol=1e-9)
        ncfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=3, adapt=adapt)
        nref = orc.nuts(ncfg, np.zeros(D), dtype=np.float64, momenta=momenta)
        nout = dm.nuts(ncfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta)
        for key in ("leapfrogs", "depth", "diverging"):
            assert np.array_equal(nout[key].cpu().numpy(), nref[key]), key
        np.testing.assert_allclose(nout["step_size"].cpu().numpy(), nref["step_size"], rtol=1e-12)
    # random-walk Metropolis: `momenta` injects the perturbations
    cfg = make_hmc_cfg(C, S, B, step_size=1.0, num_leapfrog_steps=0, seed=3, adapt=make_adapt(enabled=False),
                       proposal=PM4_PROPOSAL_RANDOM_WALK, rw_scale=0.3)
    ref = orc.hmc(cfg, np.zeros(D), dtype=np.float64, momenta=momenta, uniforms=uniforms)
    out = dm.hmc(cfg, np.zeros((C, D)), dtype=np.float64, momenta=momenta, uniforms=uniforms, want_traj=True)
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"]) and 0 < ref["accepted"].mean() < 1
    np.testing.assert_allclose(out["states"].cpu().numpy(), ref["states"], atol=1e-12)
    np.testing.assert_allclose(out["log_accept"].cpu().numpy(), ref["log_accept"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(out["traj"].cpu().numpy(), ref["traj"], atol=1e-12)


@pytest.mark.parametrize("name", ["eight_schools", "radon_synth", "poisson", "nested"])
def test_prior_predictive_matches_oracle_stream(name, oracle_mod):
    """ancestral samples of the whole model on the device against the oracle's identical Philox
    stream (FP32 against the FP32 oracle: the draws use precision-specific uniforms)."""
    ir, _, _ = lower_model(MODELS[name]())
    n = 300
    xr, obr = oracle_mod.Oracle(ir).prior_predictive(n, seed=9, dtype=np.float32)
    x, ob = _dm(ir).prior_predictive(n, seed=9, dtype=np.float32)
    x, ob = x.cpu().numpy(), ob.cpu().numpy()
    # heavy-tailed priors (HalfCauchy scales) amplify rounding: compare where the draw is moderate
    ok = np.abs(xr) < 1e3
    assert ok.mean() > 0.98
    assert np.max(np.abs(x - xr)[ok] / np.maximum(np.abs(xr[ok]), 1.0)) < 1e-4
    if name == "poisson":
        rows = np.all(np.abs(xr) < 3, axis=1)  # moderate rates: integer draws must coincide
        assert rows.sum() > 50 and np.mean(ob[rows] == obr[rows]) > 0.98
    elif ir.ll_row:
        rows = np.all(np.abs(xr) < 50, axis=1)
        assert np.max(np.abs(ob - obr)[rows] / np.maximum(np.abs(obr[rows]), 1.0)) < 1e-3


def test_full_size_invariants_radon_4096_chains():
    """BASELINE configs[1] at full size (919 observations / 85 counties, 4096 chains) through size-independent
    properties: the log-density the sampler tracks equals an independent evaluation of the traced states, tree
    sizes are consistent with tree depths, the run is reproducible, and a sharded run equals the unsharded one."""
    import torch

    ir, _, _ = lower_model(M.radon_model(real=False))
    dm = _dm(ir)
    C, B, S = 4096, 30, 8
    adapt = make_adapt(mode=PM4_EPS_POOLED, num_adaptation_steps=B)
    cfg = make_nuts_cfg(C, S, B, step_size=0.1, seed=123, adapt=adapt)
    out = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    states, lp = out["states"], out["lp"]
    assert tuple(states.shape) == (S, C, ir.D) and bool(torch.isfinite(states).all())
    lp2, _ = dm.logp_grad(states.reshape(-1, ir.D), dtype=np.float32, want_grad=False)
    rel = (lp2.reshape(S, C) - lp).abs() / lp.abs().clamp_min(1.0)
    assert float(rel.max()) < 1e-5
    leaps, depth = out["leapfrogs"], out["depth"]
    assert int(leaps.min()) >= 1 and int(depth.max()) <= 10
    assert bool((leaps <= (2 ** depth.long() - 1)).all()) and bool((leaps >= 2 ** (depth.long() - 1)).all())
    eps = out["step_size"]
    assert bool((eps == eps[0]).all()) and 1e-3 < float(eps[0]) < 1.0  # one pooled epsilon
    assert 0.3 < float(torch.exp(out["log_accept"]).mean()) <= 1.0
    # same seed, same trace (scheduling order does not enter any result)
    again = dm.nuts(cfg, np.zeros((C, ir.D), np.float32), dtype=np.float32)
    assert torch.equal(again["states"], states) and torch.equal(again["leapfrogs"], leaps)
    # per-chain step sizes: chains are independent, so a shard of the chains reproduces its slice exactly
    adapt_pc = make_adapt(mode=PM4_EPS_PER_CHAIN, num_adaptation_steps=B)
    full = dm.nuts(make_nuts_cfg(C, S, B, step_size=0.1, seed=7, adapt=adapt_pc), np.zeros((C, ir.D), np.float32),
                   dtype=np.float32)
    n, off = 1000, 2048
    shard = dm.nuts(make_nuts_cfg(n, S, B, step_size=0.1, seed=7, adapt=adapt_pc, chain_offset=off),
                    np.zeros((n, ir.D), np.float32), dtype=np.float32)
    assert torch.equal(shard["states"], full["states"][:, off:off + n])
    assert torch.equal(shard["leapfrogs"], full["leapfrogs"][:, off:off + n])
    assert int(full["total_leapfrogs"].sum()) > C * (B + S)


def test_full_size_invariants_logit_1m_observations():
    """BASELINE configs[2] at full size (N = 10^6, P = 100) through a size-independent property, additivity over
    observations: log-density and gradient of the full model minus those of the model holding the first half of the
    data are the likelihood terms of the second half, computed here in float64 NumPy."""
    n, p, C = 1_000_000, 100, 128
    X, y = M.logistic_glm_data(n=n, p=p, seed=20261019)
    ir, _, _ = lower_model(M.logistic_glm_model(X, y))
    ir_a, _, _ = lower_model(M.logistic_glm_model(X[: n // 2], y[: n // 2]))
    rng = np.random.default_rng(81)
    theta = rng.normal(0, 0.3, (C, ir.D)).astype(np.float32)
    lp, g = _dm(ir).logp_grad(theta, dtype=np.float32)
    lp_a, g_a = _dm(ir_a).logp_grad(theta, dtype=np.float32)
    base = ir.glm_terms[0].glm.base
    beta = theta[:, base: base + p].astype(np.float64)
    eta = X[n // 2:] @ beta.T                                     # [n/2, C]
    yb = y[n // 2:, None]
    ll_b = np.sum(yb * eta - np.logaddexp(0.0, eta), axis=0)
    gr_b = (yb - 1.0 / (1.0 + np.exp(-eta))).T @ X[n // 2:]       # [C, P]
    d_lp = lp.cpu().numpy().astype(np.float64) - lp_a.cpu().numpy().astype(np.float64)
    scale = np.abs(lp.cpu().numpy()).astype(np.float64)
    assert np.max(np.abs(d_lp - ll_b) / scale) < 1e-5
    d_g = (g.cpu().numpy().astype(np.float64) - g_a.cpu().numpy().astype(np.float64))[:, base: base + p]
    for c in range(C):
        assert _rel(d_g[c], gr_b[c]) < 2e-4, c   # 3xTF32 keeps ~22 bits of beta: 3.4e-5 at this size
    # nothing else of the gradient depends on the data
    other = np.ones(ir.D, bool)
    other[base: base + p] = False
    np.testing.assert_allclose(g.cpu().numpy()[:, other], g_a.cpu().numpy()[:, other], rtol=1e-5, atol=1e-5)


def test_full_size_invariants_gp_4096_points():
    """BASELINE configs[3] at full size (n = 4096): the log-density against LAPACK's Cholesky for one chain, and
    the gradient against central differences of the device log-density."""
    X, y = M.gp_data(n=4096, d=1, seed=20261019)
    ir, _, _ = lower_model(M.gp_model(X, y))
    dm = _dm(ir, f64=True)
    theta = np.array([[0.05, -0.1, -2.2], [0.3, 0.2, -1.8]])
    lp, g = dm.logp_grad(theta, dtype=np.float64)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    ls, amp, sg = np.exp(theta[0])
    d2 = np.square(X - X.T)
    K = amp ** 2 * np.exp(-d2 / (2 * ls ** 2)) + (sg ** 2 + 1e-6) * np.eye(4096)
    L = np.linalg.cholesky(K)
    w = np.linalg.solve(L, y)
    from scipy.stats import halfcauchy, halfnorm

    ref = (-0.5 * w @ w - np.log(np.diag(L)).sum() - 0.5 * 4096 * np.log(2 * np.pi) + halfcauchy(scale=5).logpdf(ls)
           + halfcauchy(scale=5).logpdf(amp) + halfnorm(scale=1).logpdf(sg) + theta[0].sum())
    np.testing.assert_allclose(lp[0], ref, rtol=1e-9)
    eps = 1e-5
    for k in range(3):
        e = np.zeros(3)
        e[k] = eps
        up, _ = dm.logp_grad(theta + e, dtype=np.float64, want_grad=False)
        dn, _ = dm.logp_grad(theta - e, dtype=np.float64, want_grad=False)
        num = (up.cpu().numpy() - dn.cpu().numpy()) / (2 * eps)
        np.testing.assert_allclose(g[:, k], num, rtol=2e-5, atol=1e-3)


# ------------------------------------------------------------------------------------------
# dense GLM likelihood (BASELINE configs[2] shape: hierarchical logistic regression, X @ beta): the likelihood
# and its gradient for all chains at once on the tcgen05 tensor cores in 3xTF32 (csrc/pm4_glm.cu), HMC in
# lock-step over the chains (csrc/pm4_lockstep.cuh)
# ------------------------------------------------------------------------------------------
def test_tcgen05_gemm_tile_3xtf32():
    """The MMA building block alone: one 128 x 128 tile, plain TF32 (1e-3) against 3xTF32 (1e-6)."""
    import ctypes

    import torch

    from pymc4_b200 import _cabi

    L = _cabi.load_library()
    torch.manual_seed(0)
    for K in (8, 40, 104):
        A, B = torch.randn(128, K, device="cuda"), torch.randn(128, K, device="cuda")
        ref = A.double() @ B.double().T
        errs = []
        for passes in (1, 3):
            C = torch.zeros(128, 128, device="cuda")
            rc = L.pm4_glm_gemm_tile(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), 128, K, passes, None)
            assert rc == 0
            errs.append(float((C.double() - ref).abs().max() / ref.abs().max()))
        assert 1e-5 < errs[0] < 2e-3 and errs[1] < 5e-6, errs
        # the same products with the A operand held in tensor memory (negative `passes`): the form the fused GLM kernel uses
        errs = []
        for passes in (-1, -3):
            C = torch.zeros(128, 128, device="cuda")
            rc = L.pm4_glm_gemm_tile(A.data_ptr(), K, B.data_ptr(), K, C.data_ptr(), 128, K, passes, None)
            assert rc == 0
            errs.append(float((C.double() - ref).abs().max() / ref.abs().max()))
        assert 1e-5 < errs[0] < 2e-3 and errs[1] < 5e-6, errs


@pytest.mark.parametrize("n,p,C", [(3000, 20, 96), (700, 100, 300), (500, 10, 40), (333, 7, 5)])
def test_dense_glm_logp_grad_parity(n, p, C, oracle_mod):
    X, y = M.logistic_glm_data(n=n, p=p)
    ir, _, _ = lower_model(M.logistic_glm_model(X, y))
    rng = np.random.default_rng(61)
    theta = rng.normal(0, 0.4, (C, ir.D))
    theta[0] = 0.0
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    lp, g = _dm(ir).logp_grad(theta.astype(np.float32), dtype=np.float32)
    lp, g = lp.cpu().numpy(), g.cpu().numpy()
    assert np.max(np.abs(lp - lp_ref) / np.abs(lp_ref)) < 1e-5
    for c in range(C):
        assert _rel(g[c], g_ref[c]) < 1e-5, c
    lp_only, _ = _dm(ir).logp_grad(theta.astype(np.float32), dtype=np.float32, want_grad=False)
    np.testing.assert_allclose(lp_only.cpu().numpy(), lp, rtol=1e-6)
    # post-processing kernels know the dense term too
    ll = _dm(ir).log_likelihood(theta[:8].astype(np.float32), dtype=np.float32).cpu().numpy()
    np.testing.assert_allclose(ll, oracle_mod.Oracle(ir).log_likelihood(theta[:8]), rtol=1e-4, atol=1e-5)


def test_dense_glm_lockstep_hmc_parity(oracle_mod):
    """(b) for the GLM path: a 30-leapfrog trajectory from injected momenta, then an adaptive chain."""
    X, y = M.logistic_glm_data(n=2000, p=20)
    ir, _, _ = lower_model(M.logistic_glm_model(X, y))
    C, D, L = 40, ir.D, 30
    rng = np.random.default_rng(62)
    theta0 = rng.normal(0, 0.2, (C, D))
    momenta = rng.standard_normal((1, C, D))
    uniforms = rng.random((1, C))
    cfg = make_hmc_cfg(C, 1, 0, step_size=0.01, num_leapfrog_steps=L, seed=1, adapt=make_adapt(enabled=False))
    ref = oracle_mod.Oracle(ir).hmc(cfg, theta0, dtype=np.float64, momenta=momenta, uniforms=uniforms)
    out = _dm(ir).hmc(cfg, theta0.astype(np.float32), dtype=np.float32, momenta=momenta, uniforms=uniforms, want_traj=True)
    assert np.max(np.abs(out["traj"].cpu().numpy() - ref["traj"])) < 1e-4
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
    # adaptive run (per-chain and pooled dual averaging): same Philox stream as the oracle's FP32 run
    B, S = 30, 20
    for mode in (PM4_EPS_POOLED, PM4_EPS_PER_CHAIN):
        cfg = make_hmc_cfg(C, S, B, step_size=0.02, num_leapfrog_steps=5, seed=9,
                           adapt=make_adapt(mode=mode, num_adaptation_steps=B))
        ref = oracle_mod.Oracle(ir).hmc(cfg, np.zeros(D), dtype=np.float32)
        out = _dm(ir).hmc(cfg, np.zeros((C, D), np.float32), dtype=np.float32)
        # a single accept decision that rounding flips sends that chain (pooled: every chain, through epsilon) on
        # another path, so the runs are compared statistically: first transitions identical, then the same regime
        acc, racc = out["accepted"].cpu().numpy(), ref["accepted"]
        assert np.mean(acc == racc) > 0.8 and abs(acc.mean() - racc.mean()) < 0.1
        np.testing.assert_allclose(np.median(out["step_size"].cpu().numpy()), np.median(ref["step_size"]), rtol=0.25)
        np.testing.assert_allclose(out["states"].cpu().numpy().mean((0, 1)), ref["states"].mean((0, 1)), atol=0.15)


def test_dense_glm_lockstep_nuts_parity(oracle_mod):
    """NUTS on the GLM path runs in lock-step over the chains (one batched gradient per leapfrog); same Philox
    streams and the same tree arithmetic as the oracle: the first transitions replay exactly (integer tree
    statistics), later ones agree in distribution (one flipped comparison sends a chain down another path)."""
    X, y = M.logistic_glm_data(n=2000, p=20)
    ir, _, _ = lower_model(M.logistic_glm_model(X, y))
    C, D = 48, ir.D
    rng = np.random.default_rng(63)
    theta0 = rng.normal(0, 0.2, (C, D))
    # fixed step size, two transitions: trees are replayed leaf for leaf
    cfg = make_nuts_cfg(C, 2, 0, step_size=0.02, seed=5, adapt=make_adapt(enabled=False), max_tree_depth=6)
    ref = oracle_mod.Oracle(ir).nuts(cfg, theta0, dtype=np.float32)
    out = _dm(ir).nuts(cfg, theta0.astype(np.float32), dtype=np.float32)
    lf, rlf = out["leapfrogs"].cpu().numpy(), ref["leapfrogs"]
    assert np.mean(lf[0] == rlf[0]) > 0.9 and np.mean(lf == rlf) > 0.8
    same = lf[0] == rlf[0]
    assert np.array_equal(out["depth"].cpu().numpy()[0][same], ref["depth"][0][same])
    assert np.array_equal(out["diverging"].cpu().numpy()[0][same], ref["diverging"][0][same])
    d0 = np.abs(out["states"].cpu().numpy()[0] - ref["states"][0]).max(axis=1)
    assert np.median(d0[same]) < 1e-4
    np.testing.assert_allclose(out["energy"].cpu().numpy()[0][same], ref["energy"][0][same], rtol=1e-3, atol=1e-2)
    assert np.array_equal(out["total_leapfrogs"].cpu().numpy(), lf.sum(0))
    # the leapfrogs of a transition are enqueued in batches, the dense evaluation of an iteration nobody asked for falls
    # through on a device-side gate: bit-identical to one host round trip per leapfrog (PM4_LOCKSTEP_BATCH=0)
    import os

    os.environ["PM4_LOCKSTEP_BATCH"] = "0"
    try:
        one = _dm(ir).nuts(cfg, theta0.astype(np.float32), dtype=np.float32)
    finally:
        del os.environ["PM4_LOCKSTEP_BATCH"]
    for key in ("states", "leapfrogs", "depth", "diverging", "energy", "lp"):
        assert np.array_equal(out[key].cpu().numpy(), one[key].cpu().numpy()), key
    # adaptive runs, pooled and per-chain: same regime as the oracle
    B, S = 40, 30
    for mode in (PM4_EPS_POOLED, PM4_EPS_PER_CHAIN):
        cfg = make_nuts_cfg(C, S, B, step_size=0.02, seed=9, adapt=make_adapt(mode=mode, num_adaptation_steps=B),
                            max_tree_depth=6)
        ref = oracle_mod.Oracle(ir).nuts(cfg, np.zeros(D), dtype=np.float32)
        out = _dm(ir).nuts(cfg, np.zeros((C, D), np.float32), dtype=np.float32)
        np.testing.assert_allclose(np.median(out["step_size"].cpu().numpy()), np.median(ref["step_size"]), rtol=0.25)
        np.testing.assert_allclose(out["leapfrogs"].cpu().numpy().mean(), ref["leapfrogs"].mean(), rtol=0.25)
        np.testing.assert_allclose(np.exp(out["log_accept"].cpu().numpy()).mean(), np.exp(ref["log_accept"]).mean(), atol=0.08)
        np.testing.assert_allclose(out["states"].cpu().numpy().mean((0, 1)), ref["states"].mean((0, 1)), atol=0.15)
        assert out["diverging"].cpu().numpy().mean() < 0.05


@pytest.mark.parametrize("n,d", [(40, 1), (200, 2), (256, 1)])
def test_gp_marginal_likelihood_logp_grad_parity(n, d, oracle_mod):
    """MvNormal(0, ExpQuad(l, a)(X, X) + sigma^2 I) observed.  The device factorises K per chain in
    float64 (blocked Cholesky, inverse factor, contraction with dK); compared with the oracle's plain loops."""
    X, y = M.gp_data(n=n, d=d)
    ir, _, _ = lower_model(M.gp_model(X, y))
    assert len(ir.gp_terms) == 1 and ir.lockstep
    C = 12
    rng = np.random.default_rng(71)
    theta = rng.normal(0, 0.4, (C, ir.D))
    theta[:, 2] -= 1.5
    theta[0] = 0.0
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    dm64 = _dm(ir, f64=True)
    lp, g = dm64.logp_grad(theta, dtype=np.float64)
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-8, atol=1e-7)
    # float32 sampling dtype: theta, lp and grad are float32, the factorisation stays float64
    lp32, g32 = _dm(ir).logp_grad(theta.astype(np.float32), dtype=np.float32)
    lp_r32, g_r32 = oracle_mod.Oracle(ir).logp_grad(theta.astype(np.float32).astype(np.float64), dtype=np.float64)
    np.testing.assert_allclose(lp32.cpu().numpy(), lp_r32, rtol=2e-6)
    for c in range(C):
        assert _rel(g32[c].cpu().numpy(), g_r32[c]) < 1e-5
    lp_only, _ = dm64.logp_grad(theta, dtype=np.float64, want_grad=False)
    np.testing.assert_allclose(lp_only.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-8)
    ll = dm64.log_likelihood(theta[:5], dtype=np.float64).cpu().numpy()
    np.testing.assert_allclose(ll, oracle_mod.Oracle(ir).log_likelihood(theta[:5]), rtol=1e-10, atol=1e-8)
    assert ll.shape == (5, 1)


def test_gp_chains_go_through_a_small_workspace_in_chunks(oracle_mod, monkeypatch):
    """2 n^2 doubles of workspace per chain: when not all chains fit, they are processed in chunks."""
    X, y = M.gp_data(n=200, d=1)
    ir, _, _ = lower_model(M.gp_model(X, y))
    rng = np.random.default_rng(73)
    theta = rng.normal(0, 0.3, (13, ir.D))
    theta[:, 2] -= 1.5
    lp_ref, g_ref = oracle_mod.Oracle(ir).logp_grad(theta, dtype=np.float64)
    monkeypatch.setenv("PM4_GP_WORK_MB", "6")  # 1.06 MB per chain at n = 200 (padded to 256): 5 chains per chunk
    lp, g = _dm(ir, f64=True).logp_grad(theta, dtype=np.float64)
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(g.cpu().numpy(), g_ref, rtol=1e-8, atol=1e-7)
    monkeypatch.setenv("PM4_GP_WORK_MB", "0")
    with pytest.raises(RuntimeError, match="workspace"):
        _dm(ir, f64=True).logp_grad(theta, dtype=np.float64)


def test_gp_forward_sampling_matches_the_oracle_stream(oracle_mod):
    """sample_posterior_predictive for the MvNormal observation of a GP model: loc + chol(K(theta)) z per state, z from
    the Philox stream the oracle draws too."""
    X, y = M.gp_data(n=150, d=2)
    ir, _, _ = lower_model(M.gp_model(X, y))
    rng = np.random.default_rng(74)
    theta = rng.normal(0, 0.3, (9, ir.D))
    theta[:, 2] -= 1.0
    ref = oracle_mod.Oracle(ir).gp_predictive(theta, seed=5)
    out = _dm(ir, f64=True).gp_predictive(theta, seed=5, dtype=np.float64).cpu().numpy()
    assert out.shape == (9, 150)
    np.testing.assert_allclose(out, ref, rtol=1e-9, atol=1e-9)
    out32 = _dm(ir).gp_predictive(theta.astype(np.float32), seed=5, dtype=np.float32).cpu().numpy()
    ref32 = oracle_mod.Oracle(ir).gp_predictive(theta.astype(np.float32).astype(np.float64), seed=5)
    np.testing.assert_allclose(out32, ref32, rtol=1e-4, atol=1e-5)
    assert not np.allclose(out, _dm(ir, f64=True).gp_predictive(theta, seed=6, dtype=np.float64).cpu().numpy())


def test_gp_lockstep_hmc_and_nuts_against_oracle(oracle_mod):
    X, y = M.gp_data(n=96, d=1)
    ir, _, _ = lower_model(M.gp_model(X, y))
    C, D = 24, ir.D
    rng = np.random.default_rng(72)
    theta0 = rng.normal(0, 0.1, (C, D))
    theta0[:, 2] -= 1.0
    momenta = rng.standard_normal((1, C, D))
    uniforms = rng.random((1, C))
    cfg = make_hmc_cfg(C, 1, 0, step_size=0.01, num_leapfrog_steps=10, seed=1, adapt=make_adapt(enabled=False))
    ref = oracle_mod.Oracle(ir).hmc(cfg, theta0, dtype=np.float64, momenta=momenta, uniforms=uniforms)
    out = _dm(ir).hmc(cfg, theta0.astype(np.float32), dtype=np.float32, momenta=momenta, uniforms=uniforms, want_traj=True)
    assert np.max(np.abs(out["traj"].cpu().numpy() - ref["traj"])) < 1e-4
    assert np.array_equal(out["accepted"].cpu().numpy(), ref["accepted"])
    cfg = make_nuts_cfg(C, 30, 40, step_size=0.05, seed=4, adapt=make_adapt(num_adaptation_steps=40), max_tree_depth=5)
    ref = oracle_mod.Oracle(ir).nuts(cfg, theta0[0] * 0, dtype=np.float32)
    out = _dm(ir).nuts(cfg, np.zeros((C, D), np.float32), dtype=np.float32)
    np.testing.assert_allclose(out["leapfrogs"].cpu().numpy().mean(), ref["leapfrogs"].mean(), rtol=0.3)
    np.testing.assert_allclose(out["states"].cpu().numpy().mean((0, 1)), ref["states"].mean((0, 1)), atol=0.25)
