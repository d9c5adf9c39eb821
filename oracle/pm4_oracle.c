/* pm4_oracle.c -- CPU oracle for the pm.sample NUTS/HMC hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (pymc4_b200) never does and fails
 * loudly when its CUDA library is missing.
 *
 * What it restates (plain C, float and double):
 *   - the joint log-density of a lowered model: sum of reduce_sum(dist.log_prob(value)) plus
 *     potentials (/root/reference/pymc4/flow/executor.py:171-187), evaluated in unconstrained
 *     space with the transform log-det potentials (flow/transformed_executor.py:69-105), and its
 *     gradient by reverse-mode differentiation of the same tape;
 *   - the sampler the reference hands that density to
 *     (/root/reference/pymc4/mcmc/samplers.py:172-189): tfp.mcmc.sample_chain over
 *     NoUTurnSampler / HamiltonianMonteCarlo wrapped in DualAveragingStepSizeAdaptation.
 *
 * Provenance of the sampler arithmetic: it lives in the third-party dependency
 * tensorflow-probability (`tfp-nightly`, UNPINNED in /root/reference/requirements.txt:4; the API
 * the reference uses dates it to TFP ~0.12 / TF ~2.4), which is absent from /root/reference and
 * not installable here.  The algorithm is restated from the published TFP sources
 * (tensorflow_probability/python/mcmc/{nuts.py, sample.py, hmc.py,
 * dual_averaging_step_size_adaptation.py, internal/leapfrog_integrator.py}); SURVEY.md App. A is
 * the spec.  Parity pinning: the log-density part is pinned by the reference's own golden value
 * (tests/test_8schools.py:61, -42.876114) and scipy cross-checks (tests/test_sampling.py:43-61);
 * the NUTS/HMC internals are "parity unpinned" -- the reference holds no golden chains, tree
 * sizes, divergence counts or step-size traces (tests/test_8schools.py:65-66 says so) and TFP
 * cannot be run here.  Random numbers are a counter-based Philox4x32-10 stream shared with the
 * CUDA kernels (not TF's stateless RNG), so oracle and device consume identical draws.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/pm4b200.h"

/* purposes: 4th word of the Philox counter (must match csrc/pm4_philox.cuh) */
#define PM4_RNG_MOMENTUM 0u
#define PM4_RNG_DIRECTION 1u
#define PM4_RNG_LEAF 2u
#define PM4_RNG_MERGE 3u
#define PM4_RNG_HMC 4u
#define PM4_RNG_PREDICTIVE 5u

void pm4o_philox4x32_10(const uint32_t ctr_in[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ---- float instantiation ---- */
#define REAL float
#define REAL_IS_DOUBLE 0
#define NAME(x) x##_f32
#define REXP expf
#define RLOG logf
#define RLOG1P log1pf
#define RSQRT sqrtf
#define RPOW powf
#define RTANH tanhf
#define RABS fabsf
#define RLGAMMA lgammaf
#define RCOS cosf
#define RSIN sinf
#define RTAN tanf
#define RFLOOR floorf
#include "pm4_oracle_impl.h"
#undef REAL
#undef REAL_IS_DOUBLE
#undef NAME
#undef REXP
#undef RLOG
#undef RLOG1P
#undef RSQRT
#undef RPOW
#undef RTANH
#undef RABS
#undef RLGAMMA
#undef RCOS
#undef RSIN
#undef RTAN
#undef RFLOOR

/* ---- double instantiation ---- */
#define REAL double
#define REAL_IS_DOUBLE 1
#define NAME(x) x##_f64
#define REXP exp
#define RLOG log
#define RLOG1P log1p
#define RSQRT sqrt
#define RPOW pow
#define RTANH tanh
#define RABS fabs
#define RLGAMMA lgamma
#define RCOS cos
#define RSIN sin
#define RTAN tan
#define RFLOOR floor
#include "pm4_oracle_impl.h"

/* number of OpenMP threads of the chain loops (a launcher such as torchrun exports OMP_NUM_THREADS=1) */
#ifdef _OPENMP
#include <omp.h>
void pm4o_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int pm4o_max_threads(void) { return omp_get_max_threads(); }
#else
void pm4o_set_num_threads(int n) { (void)n; }
int pm4o_max_threads(void) { return 1; }
#endif

/* ---- dtype-dispatching entry points (signatures mirror include/pm4b200.h, host pointers) ---- */
int pm4o_logp_grad(const pm4_ir_t* ir, int dtype, int C, const void* theta, void* lp, void* grad) {
  return dtype == PM4_F64 ? pm4o_logp_grad_f64(ir, C, theta, lp, grad)
                          : pm4o_logp_grad_f32(ir, C, theta, lp, grad);
}

int pm4o_log_likelihood(const pm4_ir_t* ir, int dtype, int64_t M, int ll_row, const void* theta, void* out) {
  return dtype == PM4_F64 ? pm4o_log_likelihood_f64(ir, M, ll_row, theta, out)
                          : pm4o_log_likelihood_f32(ir, M, ll_row, theta, out);
}

int pm4o_posterior_predictive(const pm4_ir_t* ir, int dtype, int64_t M, int ll_row, uint64_t seed,
                              const void* theta, void* out) {
  return dtype == PM4_F64 ? pm4o_posterior_predictive_f64(ir, M, ll_row, seed, theta, out)
                          : pm4o_posterior_predictive_f32(ir, M, ll_row, seed, theta, out);
}

/* One draw of the observed vector of GP term `which` per state: out[m, :] = loc + chol(K(theta_m)) z_m with
 * z_m[i] = the first Box-Muller normal of Philox(state m, element 0x40000000 + (which << 24) + i, PM4_RNG_PREDICTIVE; seed)
 * (MvNormal.sample of /root/reference/pymc4/distributions/multivariate.py:159-199 inside sample_posterior_predictive,
 * forward_sampling.py:230-427; TF's own RNG stream is not reproduced).  theta [M, D] and out [M, n] are doubles. */
int pm4o_gp_predictive(const pm4_ir_t* ir, int which, int64_t M, uint64_t seed, const double* theta, double* out) {
  if (!ir || !theta || !out || which < 0) return PM4_EINVAL;
  const pm4_term_t* tm = NULL;
  int seen = 0;
  for (int t = 0; t < ir->n_terms; ++t)
    if (ir->terms[t].kind == PM4_T_MVN_EXPQUAD && seen++ == which) { tm = &ir->terms[t]; break; }
  if (!tm) return PM4_EINVAL;
  const int n = tm->glm_base, d = tm->glm_p, D = ir->D;
  const double *X = ir->dataf + tm->glm_x, *par = ir->dataf + tm->gp_par, *loc = ir->dataf + (int64_t)par[8];
  const int el = (int)par[0], ea = (int)par[1], en = (int)par[2], sq = par[6] != 0.0;
  double* x = (double*)malloc(sizeof(double) * (size_t)(D + 1));
  double* K = (double*)malloc(sizeof(double) * (size_t)n * n);
  double* z = (double*)malloc(sizeof(double) * (size_t)n);
  int rc = PM4_OK;
  for (int64_t m = 0; m < M; ++m) {
    constrain_f64(ir, theta + (size_t)m * D, x);
    const double l = el >= 0 ? x[el] : par[3], a = ea >= 0 ? x[ea] : par[4], v = en >= 0 ? x[en] : par[5];
    const double dg = (sq ? v * v : v) + par[7];
    for (int i = 0; i < n; ++i)
      for (int j = 0; j <= i; ++j) {
        double d2 = 0;
        for (int f = 0; f < d; ++f) { const double df = X[(size_t)i * d + f] - X[(size_t)j * d + f]; d2 += df * df; }
        K[(size_t)i * n + j] = a * a * exp(-d2 / (2.0 * l * l)) + (i == j ? dg : 0.0);
      }
    for (int j = 0; j < n; ++j) {
      double s = K[(size_t)j * n + j];
      for (int k = 0; k < j; ++k) s -= K[(size_t)j * n + k] * K[(size_t)j * n + k];
      const double ljj = sqrt(s); /* NaN when K is not positive definite: propagates into the draws */
      K[(size_t)j * n + j] = ljj;
      for (int i = j + 1; i < n; ++i) {
        double t = K[(size_t)i * n + j];
        for (int k = 0; k < j; ++k) t -= K[(size_t)i * n + k] * K[(size_t)j * n + k];
        K[(size_t)i * n + j] = t / ljj;
      }
    }
    for (int i = 0; i < n; ++i) {
      uint32_t ctr[4] = {(uint32_t)m, (uint32_t)((uint64_t)m >> 32), 0x40000000u + ((uint32_t)which << 24) + (uint32_t)i,
                         PM4_RNG_PREDICTIVE}, u[4];
      double n1;
      pm4o_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), u);
      box_muller_f64(u[0], u[1], &z[i], &n1);
    }
    for (int i = 0; i < n; ++i) {
      double t = 0;
      for (int k = 0; k <= i; ++k) t += K[(size_t)i * n + k] * z[k];
      out[(size_t)m * n + i] = loc[i] + t;
    }
  }
  free(x); free(K); free(z);
  return rc;
}

int pm4o_prior_predictive(const pm4_ir_t* ir, int dtype, int64_t M, int ll_row, uint64_t seed, void* x_out,
                          void* obs_out) {
  return dtype == PM4_F64 ? pm4o_prior_predictive_f64(ir, M, ll_row, seed, x_out, obs_out)
                          : pm4o_prior_predictive_f32(ir, M, ll_row, seed, x_out, obs_out);
}

int pm4o_deterministics(const pm4_ir_t* ir, int dtype, int64_t M, const void* theta, void* out) {
  return dtype == PM4_F64 ? pm4o_deterministics_f64(ir, M, theta, out)
                          : pm4o_deterministics_f32(ir, M, theta, out);
}

int pm4o_nuts_run(const pm4_ir_t* ir, int dtype, const pm4_nuts_cfg_t* cfg, const void* theta0,
                  const void* step_scale, const void* momenta, void* states, void* lp,
                  int32_t* leapfrogs, uint8_t* diverging, void* energy, void* log_accept,
                  int32_t* depth, void* step_size_out, int64_t* total_leapfrogs, void* mass_scale_out) {
  return dtype == PM4_F64
             ? pm4o_nuts_run_f64(ir, cfg, theta0, step_scale, momenta, states, lp, leapfrogs,
                                 diverging, energy, log_accept, depth, step_size_out, total_leapfrogs, mass_scale_out)
             : pm4o_nuts_run_f32(ir, cfg, theta0, step_scale, momenta, states, lp, leapfrogs,
                                 diverging, energy, log_accept, depth, step_size_out, total_leapfrogs, mass_scale_out);
}

int pm4o_compound_run(const pm4_ir_t* ir, int dtype, int C, int S, int B, uint64_t seed, int chain_offset,
                      const pm4_compound_block_t* blocks, int n_blocks, const void* theta0, void* states, void* lp,
                      void* step_size, void* log_accept, int32_t* leapfrogs) {
  return dtype == PM4_F64 ? pm4o_compound_run_f64(ir, C, S, B, seed, chain_offset, blocks, n_blocks, theta0, states, lp,
                                                  step_size, log_accept, leapfrogs)
                          : pm4o_compound_run_f32(ir, C, S, B, seed, chain_offset, blocks, n_blocks, theta0, states, lp,
                                                  step_size, log_accept, leapfrogs);
}

int pm4o_hmc_run(const pm4_ir_t* ir, int dtype, const pm4_hmc_cfg_t* cfg, const void* theta0,
                 const void* step_scale, const void* momenta, const void* uniforms, void* states,
                 void* lp, void* log_accept, uint8_t* accepted, void* step_size_out, void* traj) {
  return dtype == PM4_F64 ? pm4o_hmc_run_f64(ir, cfg, theta0, step_scale, momenta, uniforms, states,
                                             lp, log_accept, accepted, step_size_out, traj)
                          : pm4o_hmc_run_f32(ir, cfg, theta0, step_scale, momenta, uniforms, states,
                                             lp, log_accept, accepted, step_size_out, traj);
}

/* momentum draw exposed for tests (device/host stream equality) */
void pm4o_draw_momentum(int dtype, uint64_t seed, uint32_t chain, uint32_t t, int D, void* p) {
  if (dtype == PM4_F64) draw_momentum_f64(seed, chain, t, D, (double*)p);
  else draw_momentum_f32(seed, chain, t, D, (float*)p);
}

int pm4o_abi_version(void) { return PM4_ABI_VERSION; }
