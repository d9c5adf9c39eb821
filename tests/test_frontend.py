"""Host-logic tests: coroutine models, name scopes, executors, lowering to the IR, execution plans.
Mirrors the contracts the reference pins in tests/test_executor.py, tests/test_utils.py and
tests/test_sampling.py (names, observed handling, errors, nested models, deterministics)."""
import numpy as np
import pytest

import pymc4_b200 as pm
from pymc4_b200 import flow, ir as I
from pymc4_b200.flow.executor import EvaluationError
from pymc4_b200.ir import lower_model
from pymc4_b200.utils import NameParts, biwrap
from tests import models as M


def test_name_parts_and_scopes():
    p = NameParts.from_name("a/b/__log_tau")
    assert (p.path, p.transform_name, p.untransformed_name) == (("a", "b"), "log", "tau")
    assert p.full_untransformed_name == "a/b/tau" and p.full_original_name == "a/b/__log_tau" and p.is_transformed
    assert NameParts.is_valid_untransformed_name("tau") and not NameParts.is_valid_untransformed_name("__log_tau")
    with pytest.raises(ValueError):
        NameParts.from_name("_bad")
    with pm.name_scope("outer"):
        with pm.name_scope(None):
            assert pm.variable_name("x") == "outer/x"
            assert pm.scopes.transformed_variable_name("log", "x") == "outer/__log_x"
    assert pm.variable_name(None) is None and pm.variable_name("") is None


def test_biwrap_both_spellings():
    @biwrap
    def deco(fn, *, k=1):
        return (fn, k)

    def f():
        pass

    assert deco(f) == (f, 1)
    assert deco(k=3)(f) == (f, 3)


def test_model_decorator_and_names():
    @pm.model
    def plain():
        yield pm.Normal("n", 0, 1)

    @pm.model(name="custom", keep_return=False)
    def named():
        x = yield pm.Normal("n", 0, 1)
        return x

    assert plain().name == "plain" and named().name == "custom" and named(name="other").name == "other"
    with pytest.raises(ValueError):
        plain(name="__log_x")
    _, st = pm.evaluate_model(named())
    assert "custom" not in st.deterministics_values and "custom/n" in st.untransformed_values


def test_executor_states_and_values():
    _, st = pm.evaluate_model_transformed(
        M.schools_pm4(), values={"schools_pm4/eta": np.zeros(8), "schools_pm4/mu": 0.0, "schools_pm4/__log_tau": 1.0})
    assert set(st.transformed_values) == {"schools_pm4/__log_tau"}
    assert set(st.observed_values) == {"schools_pm4/obs"}
    np.testing.assert_allclose(st.untransformed_values["schools_pm4/tau"], np.e)
    np.testing.assert_allclose(st.collect_log_prob(), -42.87611367240224, rtol=1e-13)  # reference golden
    assert len(st.potentials) == 1 and "schools_pm4" in st.deterministics_values
    # the untransformed executor refuses transformed values (reference executor.py:569-574)
    with pytest.raises(ValueError):
        pm.evaluate_model(M.schools_pm4(), values={"schools_pm4/__log_tau": 1.0})


def test_meta_executor_initial_state():
    st, det_names = pm.initialize_sampling_state(M.radon_model(real=True))
    keys = list(st.all_unobserved_values)
    assert keys[:4] == ["hierarchical_model/mu_alpha", "hierarchical_model/mu_beta",
                        "hierarchical_model/alpha", "hierarchical_model/beta"]  # untransformed first
    assert keys[4:] == ["hierarchical_model/__log_sigma_alpha", "hierarchical_model/__log_sigma_beta",
                        "hierarchical_model/__log_eps"]
    assert all(np.all(np.asarray(v) == 0) for v in st.all_unobserved_values.values())  # test point: theta = 0
    assert det_names == ["hierarchical_model/sigma_alpha", "hierarchical_model/sigma_beta", "hierarchical_model/eps"]


def test_observed_handling_and_errors():
    @pm.model
    def m():
        x = yield pm.Normal("x", 0, 1, observed=1.0)
        y = yield pm.Normal("y", x, 1)
        return y

    _, st = pm.evaluate_model(m())
    assert "m/x" in st.observed_values and "m/y" in st.untransformed_values
    # suppress the observation -> it becomes a sampled (posterior-predictive) variable
    _, st = pm.evaluate_model(m(), observed={"m/x": None})
    assert "m/x" in st.untransformed_values and "m/x" in st.posterior_predictives
    # replace the observation
    _, st = pm.evaluate_model(m(), observed={"m/x": 5.0})
    assert st.observed_values["m/x"] == 5.0
    # a value for an observed variable without suppressing it is an error
    with pytest.raises(EvaluationError):
        pm.evaluate_model(m(), values={"m/x": 2.0})
    with pytest.raises(ValueError):
        pm.Normal(None, 0, 1, observed=1.0)

    @pm.model
    def dup():
        yield pm.Normal("x", 0, 1)
        yield pm.Normal("x", 0, 1)

    with pytest.raises(EvaluationError):
        pm.evaluate_model(dup())

    @pm.model
    def bad_yield():
        yield 3

    with pytest.raises((EvaluationError, flow.StopExecution)):
        pm.evaluate_model(bad_yield())

    @pm.model
    def bad_shape():
        yield pm.Normal("x", np.zeros(3), 1, observed=np.zeros(4))

    with pytest.raises(EvaluationError):
        pm.evaluate_model(bad_shape())


def test_nested_models_and_deterministics():
    _, st = pm.evaluate_model_transformed(M.outer_model())
    assert set(st.deterministics_values) == {"outer_model/dcond", "outer_model/nested_model/dx",
                                             "outer_model/nested_model", "outer_model/ddx", "outer_model"}
    assert set(st.transformed_values) == {"outer_model/__log_cond"}
    ir, init, names = lower_model(M.outer_model())
    assert [rv.name for rv in ir.rvs] == ["outer_model/nested_model/x", "outer_model/__log_cond"]
    assert names[-1] == "outer_model/cond" and {d.name for d in ir.dets} == set(names)


def test_lowering_eight_schools():
    ir, init, names = lower_model(M.schools_pm4())
    assert ir.D == 10 and [(rv.name, rv.offset, rv.size, rv.transform) for rv in ir.rvs] == [
        ("schools_pm4/eta", 0, 8, I.TR_IDENTITY), ("schools_pm4/mu", 8, 1, I.TR_IDENTITY),
        ("schools_pm4/__log_tau", 9, 1, I.TR_EXP)]
    kinds = [t.kind for t in ir.terms]
    assert kinds == [I.TERM_KINDS["normal"], I.TERM_KINDS["normal"], I.TERM_KINDS["halfnormal"],
                     I.TERM_KINDS["normal"], I.TERM_KINDS["potential"]]
    assert [t.n for t in ir.terms] == [8, 1, 1, 8, 1]
    # structure hash ignores data values but not structure
    y2 = M.Y8 + 1

    @pm.model
    def schools_pm4():
        eta = yield pm.Normal("eta", 0, 1, batch_stack=8)
        mu = yield pm.Normal("mu", 1, 10)
        tau = yield pm.HalfNormal("tau", 2.0)
        obs = yield pm.Normal("obs", mu + tau * eta, M.SIGMA8, observed=y2)
        return obs

    ir2, _, _ = lower_model(schools_pm4())
    assert ir2.struct_hash == ir.struct_hash and not np.array_equal(ir2.dataf, ir.dataf)
    ir3, _, _ = lower_model(M.radon_model())
    assert ir3.struct_hash != ir.struct_hash


def test_errors_of_the_logp_builder():
    from pymc4_b200.mcmc.samplers import build_logp_and_deterministic_functions

    with pytest.raises(ValueError):  # no free variables (reference samplers.py:748-751)
        lower_model(M.simple_model_no_free_rvs())
    with pytest.raises(TypeError):  # not a Model (reference samplers.py:737-742)
        build_logp_and_deterministic_functions(lambda: None)
    with pytest.raises(ValueError):  # state and observed together (reference samplers.py:743-744)
        build_logp_and_deterministic_functions(M.simple_model(), observed={"a": 1}, state=flow.SamplingState())

    @pm.model
    def discrete_free():
        yield pm.Bernoulli("b", 0.5)

    with pytest.raises(ValueError):  # discrete + gradient sampler (reference samplers.py:62-66)
        pm.mcmc.samplers.NUTS(discrete_free())
    with pytest.raises(KeyError):  # unknown sampler (reference tests/test_sampling.py:138-141)
        pm.sample(M.simple_model(), sampler_type="unknown")
    with pytest.raises(NotImplementedError):  # kernel options the CUDA kernels do not implement fail loudly
        pm.mcmc.kernels.NoUTurnSampler(None, 0.1, unrolled_leapfrog_steps=2)


def test_kwarg_routing_by_signature():
    s = pm.mcmc.samplers.NUTS(M.simple_model(), step_size=0.3, max_tree_depth=7, target_accept_prob=0.8,
                              num_adaptation_steps=5, xla_fixture=True, dtype="float64")
    assert s.kernel_kwargs["step_size"] == 0.3 and s.kernel_kwargs["max_tree_depth"] == 7
    assert s.adaptation_kwargs["target_accept_prob"] == 0.8 and s.adaptation_kwargs["num_adaptation_steps"] == 5
    assert "xla_fixture" not in s.kernel_kwargs and s.backend_kwargs == {"dtype": "float64"}
    h = pm.mcmc.samplers.HMC(M.simple_model())
    assert h.kernel_kwargs == {"step_size": 0.1, "num_leapfrog_steps": 3}
    assert s.stat_names == ["lp", "tree_size", "diverging", "energy", "mean_tree_accept"]


def test_sampler_names_come_from_its_own_lowered_model():
    """Device models are shared between models of one structure and layout (`_get_device_model`); the IR they carry
    was lowered for the first of them.  A sampler therefore reads names from the IR of ITS model (`_ir`), and two
    models that differ only in names do hash to the same structure."""
    import types

    @pm.model
    def other_names():
        value = yield pm.Normal("value", 0, 1)
        return value

    ir_a, _, _ = I.lower_model(M.simple_model())
    ir_b, _, _ = I.lower_model(other_names())
    assert ir_a.struct_hash == ir_b.struct_hash and ir_a.layout_key() == ir_b.layout_key()
    assert [rv.name for rv in ir_a.rvs] != [rv.name for rv in ir_b.rvs]
    s = pm.mcmc.samplers.NUTS(other_names())
    s._device_model = types.SimpleNamespace(ir=ir_a)  # the cached device model of the first model
    assert s._ir is ir_a  # nothing lowered yet: the device model's
    s._ir_own = ir_b
    assert s._ir is ir_b and [rv.name for rv in s._ir.rvs] == ["other_names/value"]


@pytest.mark.parametrize("sorted_idx,team", [(True, 1), (False, 1), (True, 4), (False, 2)])
def test_execution_plan_is_a_permutation(sorted_idx, team):
    rng = np.random.default_rng(1)
    n, G = 700, 37
    idx = rng.integers(0, G, n).astype(np.int32)
    idx[:G] = np.arange(G)
    if sorted_idx:
        idx = np.sort(idx)
    x, y = rng.normal(size=n), rng.normal(size=n)
    ir, _, _ = lower_model(M.hierarchical_model(idx, x, y, G), team_warps=team)
    assert ir.team_warps == team
    TL = 32 * team
    (pl,) = ir.plans
    n_slices, meta_off, recs_off, n_term = ir.plani[ir.phdr_off:ir.phdr_off + 4]
    assert (meta_off, recs_off, n_term) == (pl.meta_off, pl.recs_off, n) and n_slices == pl.n_slices
    meta = ir.plani[pl.meta_off:]
    slices = meta[:4 * n_slices].reshape(-1, 4)
    # two gather ops (alpha, beta): target, length and two part indices fill one int4 per slot
    slots = meta[4 * n_slices:4 * n_slices + 4 * TL * n_slices].reshape(-1, 4)
    recs = ir.planf[pl.recs_off:pl.recs_off + int(slices[:, 1].sum()) * TL * pl.width].reshape(-1, TL, pl.width)
    assert list(slices[:, 2]) == list(np.concatenate([[0], np.cumsum(slices[:, 1])])[:-1])
    seen = []
    per_target = {}
    for slot, (tgt, ln, pa, pb) in enumerate(slots):
        s, lane = divmod(slot, TL)
        full, mx, first, _ = slices[s]
        assert ln <= mx and (tgt < 0) == (ln == 0) and ln >= full
        if tgt >= 0:
            rows = recs[first:first + ln, lane]
            seen.append(rows)
            per_target.setdefault(int(tgt), []).append((ln, int(pa), int(pb)))
            assert np.all(recs[first + ln:first + mx, lane] == 0)  # padding
    got = np.concatenate(seen)
    want = np.stack([y, x], axis=1)  # record order = op order of the data loads: value y first, then x
    assert got.shape == want.shape
    assert np.array_equal(np.sort(got.view([("a", float), ("b", float)]), axis=0),
                          np.sort(want.view([("a", float), ("b", float)]), axis=0))
    # the part array: K fixed entries per gradient element, overflow blocks, log-density partials;
    # every entry has at most one writer and every contribution is reachable from its owner
    K = ir.part_k
    R = -(-ir.D // TL)
    elem = ir.plani[ir.elem_off:ir.elem_off + TL * R]
    ofirst, ocnt = elem >> 12, elem & 4095
    assert np.all(elem[ir.D:] == 0)

    assert K == 4 and np.all(ofirst[ocnt > 0] % 4 == 0)

    def owned(j):
        return list(range(K * j, K * j + K)) + list(range(ofirst[j], ofirst[j] + 4 * ocnt[j]))

    spart = ir.plani[ir.spart_off:ir.spart_off + (len(ir.scalar_pos) + 1) * team].reshape(-1, team)
    writers = np.zeros(ir.part_reals, dtype=int)
    sizes = np.bincount(idx, minlength=G)
    a_off, b_off = ir.rv_by_name("hierarchical_model/alpha").offset, ir.rv_by_name("hierarchical_model/beta").offset
    for tgt, chunks in per_target.items():
        assert sum(l for l, _, _ in chunks) == sizes[tgt]
        for off, col in ((a_off, 1), (b_off, 2)):
            ps = [c[col] for c in chunks]
            assert sorted(ps) == owned(off + tgt)[:len(ps)]  # filled in order: fixed entries first
            writers[ps] += 1
    for k, pos in enumerate(ir.scalar_pos):  # scalar parameters: one partial per warp
        assert set(spart[k]) <= set(owned(pos))
        writers[spart[k]] += 1
    writers[spart[-1]] += 1  # log-density partials, one per warp, one aligned vector
    assert writers.max() == 1 and spart[-1][0] % 4 == 0 and list(spart[-1]) == list(range(spart[-1][0], spart[-1][0] + team))
    assert ir.part_reals >= spart[-1][-1] + 1  # entries nobody writes are zeroed once per launch


def test_step_size_conventions():
    s = pm.mcmc.samplers.NUTS(M.schools_pm4())

    class _DM:
        pass

    dm = _DM()
    dm.ir, _, _ = lower_model(M.schools_pm4())
    s._device_model = dm
    from pymc4_b200._cstructs import PM4_EPS_PER_CHAIN, PM4_EPS_POOLED

    assert s._step_size_policy(0.28, 4) == (0.28, None, PM4_EPS_POOLED)
    assert s._step_size_policy(np.full((4, 1), 0.2), 4) == (0.2, None, PM4_EPS_PER_CHAIN) and s._chain_step_sizes is None
    # chains that start from different step sizes: the array goes to the device next to the first value
    assert s._step_size_policy(np.array([0.1, 0.2, 0.3, 0.4]), 4) == (0.1, None, PM4_EPS_PER_CHAIN)
    np.testing.assert_allclose(s._chain_step_sizes, [0.1, 0.2, 0.3, 0.4])
    with pytest.raises(ValueError):
        s._step_size_policy(np.array([0.1, -0.2, 0.3, 0.4]), 4)
    # tfp's shrinkage_target reaches the adaptation record
    s2 = pm.mcmc.samplers.NUTS(M.schools_pm4(), num_adaptation_steps=10, shrinkage_target=0.5)
    s2._device_model = dm
    assert s2._adapt_cfg(PM4_EPS_POOLED).shrinkage_target == 0.5 and s._adapt_cfg(PM4_EPS_POOLED).shrinkage_target == 0.0
    parts = [np.full((4, 8), 0.5), np.full((4,), 0.25), np.full((4,), 2.0)]
    eps0, scale, mode = s._step_size_policy(parts, 4)
    assert eps0 == 1.0 and mode == PM4_EPS_PER_CHAIN
    np.testing.assert_allclose(scale, [0.5] * 8 + [0.25, 2.0])


def test_trace_to_arviz_layout():
    post = {"m/x": np.arange(24.0).reshape(3, 4, 2), "m": np.zeros((3, 4))}  # [draw, chain, core]
    stats = {"lp": np.arange(12.0).reshape(3, 4)}
    tr = pm.trace_to_arviz(post, stats, observed_data={"m/y": np.ones(5)})
    assert set(tr.posterior) == {"m/x"}  # keys without "/" are dropped (reference utils.py:122)
    assert tr.posterior["m/x"].shape == (4, 3, 2) and tr.sample_stats["lp"].shape == (4, 3)
    assert tr.posterior["m/x"][1, 2, 0] == post["m/x"][2, 1, 0]
    assert tr.groups() == ["posterior", "sample_stats", "observed_data"]


def test_expquad_covariance_function_mirrors_the_reference():
    """pm.gp.cov.ExpQuad against the reference's own expectations (/root/reference/tests/test_gp_cov.py:306-391,
    pymc4/gp/cov.py:430-438): docstring matrix, the 2-point value behind test_scaled_cov (2 x 0.01831564 = 0.03663128),
    active_dims slicing, argument validation, stabilize."""
    import pymc4_b200 as pm

    k = pm.gp.cov.ExpQuad(1.0)
    x = np.array([[1.0, 2.0], [3.0, 4.0]])
    np.testing.assert_allclose(2 * k(x, x), [[2.0, 0.03663128], [0.03663128, 2.0]], rtol=1e-6)
    np.testing.assert_allclose(k(x[:, :1], x[:, :1]), pm.gp.cov.ExpQuad(1.0, 1.0, 1, 1)(x, x))       # first feature only
    np.testing.assert_allclose(pm.gp.cov.ExpQuad(1.0, 1.0, 1, [[1]])(x, x), k(x[:, 1:], x[:, 1:]))  # listed column
    rng = np.random.default_rng(0)
    X1, X2 = rng.normal(size=(20, 5)), rng.normal(size=(7, 5))
    ref = pm.gp.cov.ExpQuad(2.0, 1.5)
    cov = pm.gp.cov.ExpQuad(2.0, 1.5, 1, 3)(X1, X2)
    assert cov.shape == (20, 7)
    np.testing.assert_allclose(cov, ref(X1[:, :3], X2[:, :3]))
    np.testing.assert_allclose(pm.gp.cov.ExpQuad(2.0, 1.5, 1, [[0, 4]])(X1, X2), ref(X1[:, [0, 4]], X2[:, [0, 4]]))
    np.testing.assert_allclose(ref(X1, X1, diag=True, to_dense=False), np.full(20, 1.5 ** 2))
    np.testing.assert_allclose(ref.evaluate_kernel(X1[:7], X2), np.diag(ref(X1[:7], X2)))
    assert np.all(np.linalg.eigvalsh(pm.gp.util.stabilize(ref(X1, X1))) > 0)
    sd = np.array([1.0, 2.0, 0.5, 1.0, 4.0])
    np.testing.assert_allclose(pm.gp.cov.ExpQuad(1.0, scale_diag=sd)(X1, X2), k(X1 / sd, X2 / sd))
    with pytest.raises(ValueError, match=r"expected 'feature_ndims' to be an integer"):
        pm.gp.cov.ExpQuad(1.0, 1.0, 0)
    with pytest.raises(ValueError, match=r"active_dims' contain more entries than number of feature dimensions"):
        pm.gp.cov.ExpQuad(1.0, 1.0, 1, [1, 2, 3, 4, 5])
    # MvNormal on the host: scipy's density, shapes
    from scipy.stats import multivariate_normal

    K = pm.gp.util.stabilize(ref(X1, X1), 1e-6)
    y = rng.normal(size=20)
    d = pm.MvNormal("y", loc=0.5, covariance_matrix=K)
    assert d.event_shape == (20,) and d.batch_shape == ()
    np.testing.assert_allclose(d.log_prob(y), multivariate_normal(np.full(20, 0.5), K).logpdf(y), rtol=1e-9)
    assert d.sample(seed=1).shape == (20,)


def test_forward_sampling_argument_errors_mirror_the_reference():
    """The validations of sample_prior_predictive / sample_posterior_predictive the reference tests
    (/root/reference/tests/test_forward_sampling.py:242-248, 324-335); they run before any device work."""
    import re

    from pymc4_b200 import forward_sampling
    from pymc4_b200.mcmc.utils import trace_to_arviz

    @pm.model
    def model():
        sd = yield pm.HalfNormal("sd", 1.0)
        yield pm.Normal("y", 0.0, sd, observed=np.zeros(3))

    model_func = model()
    expected = ("Some of the supplied var_names are not defined in the supplied model {}.\nList of unknown var_names: {}"
                .format(model_func, ["X"]))
    with pytest.raises(ValueError, match=re.escape(expected)):
        pm.sample_prior_predictive(model_func, var_names=["X", "model/y"], sample_shape=())
    trace = trace_to_arviz({"model/__log_sd": np.zeros((1, 10)), "model/sd": np.ones((1, 10))})
    with pytest.raises(ValueError):
        forward_sampling.sample_posterior_predictive(model(), trace, var_names=[])
    with pytest.raises(KeyError):
        forward_sampling.sample_posterior_predictive(model(), trace, var_names=["name not in model!"])
    with pytest.raises(TypeError):
        pm.sample_posterior_predictive("not a model", trace)
    with pytest.raises(TypeError):
        pm.sample_prior_predictive("not a model")
    with pytest.raises(TypeError, match="only supports `pymc4.Model` objects"):
        pm.sample("not a model", sampler_type="nuts")


def test_dense_predictor_spellings_and_the_intercept_hint():
    """`X @ beta` and `pm.math.matvec(X, beta)` both lower to the dense GLM term; a linear predictor that is not the whole
    sigmoid argument is refused with the work-around in the message."""
    X, y = M.logistic_glm_data(n=50, p=4)

    @pm.model
    def with_matmul():
        beta = yield pm.Normal("beta", 0.0, 1.0, batch_stack=4)
        yield pm.Bernoulli("y", probs=pm.math.sigmoid(X @ beta), observed=y)

    @pm.model
    def with_intercept():
        beta = yield pm.Normal("beta", 0.0, 1.0, batch_stack=4)
        b0 = yield pm.Normal("b0", 0.0, 1.0)
        yield pm.Bernoulli("y", probs=pm.math.sigmoid(pm.math.matvec(X, beta) + b0), observed=y)

    ir, _, _ = lower_model(with_matmul())
    assert len(ir.glm_terms) == 1 and ir.glm_terms[0].glm.P == 4 and ir.lockstep
    with pytest.raises(NotImplementedError, match="append a column of ones"):
        lower_model(with_intercept())
    Xi = np.concatenate([X, np.ones((50, 1))], axis=1)

    @pm.model
    def intercept_as_a_column():
        beta = yield pm.Normal("beta", 0.0, 1.0, batch_stack=5)
        yield pm.Bernoulli("y", probs=pm.math.sigmoid(Xi @ beta), observed=y)

    ir2, _, _ = lower_model(intercept_as_a_column())
    assert ir2.glm_terms[0].glm.P == 5


def test_struct_hash_covers_the_observed_row_layout():
    """Two data sets of different length for one model must not share a kernel module: the generated pointwise
    log-likelihood / predictive kernels bake LL_ROW and every observed term's row offset in."""
    import pymc4_b200 as pm
    from pymc4_b200 import codegen
    from pymc4_b200.ir import lower_model

    def make(n):
        y = np.linspace(-1.0, 1.0, n)

        @pm.model
        def m():
            mu = yield pm.Normal("mu", 0.0, 1.0)
            s = yield pm.HalfNormal("s", 1.0)
            yield pm.Normal("y", mu, s, observed=y)

        return lower_model(m())[0]

    a, b, a2 = make(100), make(200), make(100)
    assert a.struct_hash != b.struct_hash
    assert codegen.module_path(a) != codegen.module_path(b)
    assert a.struct_hash == a2.struct_hash and a.layout_key() == a2.layout_key()
    assert "LL_ROW = 100" in codegen.generate_source(a) and "LL_ROW = 200" in codegen.generate_source(b)


def test_layout_key_separates_data_of_another_shape_or_plan():
    """Same structure, same shapes, other gather indices: the plan pools may differ in size, and then the cached device
    model must be rebuilt, not patched."""
    from pymc4_b200.ir import lower_model

    rng = np.random.default_rng(0)
    n, g = 400, 12
    x = rng.standard_normal(n)
    y = rng.standard_normal(n)
    idx_a = np.sort(rng.integers(0, g, n)).astype(np.int32)
    idx_b = np.sort(np.minimum(rng.geometric(0.4, n) - 1, g - 1)).astype(np.int32)  # very different group sizes
    ir_a = lower_model(M.hierarchical_model(idx_a, x, y, g))[0]
    ir_b = lower_model(M.hierarchical_model(idx_b, x, y, g))[0]
    assert ir_a.struct_hash == ir_b.struct_hash
    same_pools = np.asarray(ir_a.plani).shape == np.asarray(ir_b.plani).shape and np.asarray(ir_a.planf).shape == np.asarray(ir_b.planf).shape
    assert (ir_a.layout_key() == ir_b.layout_key()) == same_pools


def test_lowering_cache_follows_the_data():
    """lower_model is cached per model object behind a fingerprint of the arrays the model can reach: the same object
    lowers once, an in-place edit of its data lowers again, another model object never shares the entry."""
    idx, x, y, g = M.radon_synthetic()
    y = y.copy()
    m = M.hierarchical_model(idx, x, y, g)
    a = lower_model(m)
    assert lower_model(m)[0] is a[0]
    y[3] += 1.0
    c = lower_model(m)
    assert c[0] is not a[0] and not np.array_equal(c[0].dataf, a[0].dataf)
    assert lower_model(M.hierarchical_model(idx, x, y, g))[0] is not c[0]
    assert lower_model(m, observed={})[0] is not lower_model(m, observed={})[0]  # overrides are never cached


# ---- owner-computes plan of the warp path (pymc4_b200/owned.py), walked on the CPU the way pm4_warp.cuh walks it ----
def _decode_owned(ir):
    from pymc4_b200 import owned as OW

    lay = ir.owned
    hdr_at = ir.phdr_off + 4 * max(len(ir.plans), 1)
    en, i_off, i_len, f_off, f_len, NF, Q, n = [int(v) for v in ir.plani[hdr_at:hdr_at + OW.HDR_INTS]]
    assert en == 1 and n == ir.terms[lay.term].n
    sec = np.asarray(ir.plani[i_off:i_off + i_len], dtype=np.int64)
    RI, KS = lay.RI, lay.KS
    perm = sec[:RI * 32]
    p = RI * 32
    own_len = sec[p:p + max(KS, 1) * 32]
    p += max(KS, 1) * 32
    nsl = KS + NF + 1
    sl_base = sec[p:p + nsl]
    p += nsl + ((-nsl) % 4)
    fmeta = sec[p:p + NF * 32 * 4].reshape(NF, 32, 4)
    W = max(lay.width, 1)
    rows = int(sl_base[-1])
    rec = np.asarray(ir.planf[f_off:f_off + f_len]).reshape(-1)[:rows * 32 * W * 2].reshape(rows, 32, W, 2)
    return dict(NF=NF, Q=Q, perm=perm, own_len=own_len, sl_base=sl_base, fmeta=fmeta, rec=rec)


@pytest.mark.parametrize("G,n,conc", [(85, 919, None), (31, 200, 0.3), (33, 64, 0.3), (100, 2000, 0.3), (5, 1000, 0.3)])
def test_owned_plan_visits_every_observation_once_on_its_groups_lane(G, n, conc):
    """The owned plan (records as pair rows per lane and slice, the group -> (slice, lane) map, the foreign pieces of
    the biggest groups with their segmented shuffle reduction) is walked here exactly like `terms_o` walks it: every
    observation must be visited once, by a lane that holds (or is handed) the parameters of ITS group, padding must
    be zero, and the segmented reduction must deliver every owner the total of its pieces."""
    from pymc4_b200 import benchmodels as BM

    if conc is None:
        idx, x, y, G = BM.radon_synthetic(n, G)
    else:
        rng = np.random.default_rng(5 + G)
        sizes = 1 + rng.multinomial(n - G, rng.dirichlet(np.full(G, conc)))
        idx = np.repeat(np.arange(G, dtype=np.int32), sizes)
        x, y = (rng.random(n) < 0.3).astype(float), rng.standard_normal(n)
    ir, _, _ = lower_model(BM.hierarchical_model(idx, x, y, G))
    lay = ir.owned
    assert lay is not None and lay.NV == 2 and lay.G == G
    d = _decode_owned(ir)
    KS, NF, Q, base0 = lay.KS, d["NF"], d["Q"], lay.vec_bases[0]
    got = [[] for _ in range(G)]
    group_of_lane = {}
    for s in range(KS):
        rows = int(d["sl_base"][s + 1] - d["sl_base"][s])
        for lane in range(32):
            pj, cnt = int(d["perm"][s * 32 + lane]), int(d["own_len"][s * 32 + lane])
            if pj < 0:
                assert cnt == 0
                continue
            g = pj - base0
            for v, b in enumerate(lay.vec_bases):  # every gathered vector of the group: same lane, same slice
                assert d["perm"][(v * KS + s) * 32 + lane] == b + g
            if s == 0:
                group_of_lane[lane] = g
            assert cnt <= 2 * rows
            for r in range(2 * rows):
                cell = d["rec"][d["sl_base"][s] + r // 2, lane, :, r % 2]
                if r < cnt:
                    got[g].append(tuple(cell))
                else:
                    assert not cell.any()
    for f in range(NF):
        meta = d["fmeta"][f]
        rows = int(d["sl_base"][KS + f + 1] - d["sl_base"][KS + f])
        part = np.zeros(32)
        for lane in range(32):
            cnt, owner = int(meta[lane][0]), int(meta[lane][1])
            assert cnt <= 2 * rows
            for r in range(cnt):
                got[group_of_lane[owner]].append(tuple(d["rec"][d["sl_base"][KS + f] + r // 2, lane, :, r % 2]))
            part[lane] = cnt
        for st in range(Q):  # the kernel's segmented reduction: step st adds lane + 2^st where the mask bit is set
            part = np.array([part[l] + (part[l + (1 << st)] if int(meta[l][2]) >> st & 1 else 0) for l in range(32)])
        for owner in range(32):
            head = int(meta[owner][3])
            if head >= 0:
                assert int(meta[head][1]) == owner
                assert part[head] == sum(int(m[0]) for m in meta if int(m[1]) == owner and int(m[0]) > 0)
    t = ir.terms[lay.term]
    by_dst = {o.dst: o for o in t.ops}
    cols = [np.asarray(ir.dataf[by_dst[c].blob: by_dst[c].blob + t.n]) for c in t.plan.record_ops]
    for g in range(G):
        assert sorted(got[g]) == sorted(zip(*[c[idx == g] for c in cols])), g
    assert sum(len(v) for v in got) == n
