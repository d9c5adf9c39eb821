/* pm4b200.h -- C ABI of the B200-native NUTS/HMC sampling backend for PyMC4's pm.sample path.
 *
 * The reference has no FFI: its seam is the Python call
 *   results, sample_stats = self._run_chains(init_state, burn_in)
 * (/root/reference/pymc4/mcmc/samplers.py:136, body :172-189) which hands the vectorised
 * log-density (samplers.py:117-119, built at :729-817 from flow/executor.py:171-187) to
 * tfp.mcmc.sample_chain(NoUTurnSampler | HamiltonianMonteCarlo wrapped in
 * DualAveragingStepSizeAdaptation).  This header is what a maintainer binds (ctypes) in place
 * of that call; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types.
 *   - every function returns 0 on success or a negative PM4_E* code; pm4_last_error() gives
 *     the thread-local message.
 *   - "dev" pointers are device memory owned by the caller (e.g. torch tensors) and must stay
 *     alive until the stream has been synchronised; "host" pointers are ordinary host memory.
 *   - dtype: PM4_F32 or PM4_F64 selects the arithmetic type of every `void*` real buffer.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and is asynchronous unless stated otherwise.
 */
#ifndef PM4B200_H
#define PM4B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PM4_ABI_VERSION 6

/* error codes */
#define PM4_OK 0
#define PM4_EINVAL (-22)   /* bad argument / malformed IR                      */
#define PM4_ENOMOD (-2)    /* specialised kernel module missing or mismatched  */
#define PM4_ECUDA (-5)     /* a CUDA runtime call failed                       */
#define PM4_ENOMEM (-12)
#define PM4_ENOSYS (-38)   /* requested feature not built into this module     */

/* arithmetic type of real-valued buffers */
#define PM4_F32 0
#define PM4_F64 1

/* ------------------------------------------------------------------------------------------
 * Flat model IR: what flow.TransformedSamplingExecutor's log_prob pass
 * (/root/reference/pymc4/flow/transformed_executor.py:22-138, executor.py:171-187) is lowered to.
 *
 *   theta[D]   unconstrained state of one chain; free variables are laid out in the reference's
 *              trace order (untransformed first, then transformed; executor.py:24-35,163).
 *   x[D]       constrained values, x = shift + scale * f(theta), f per variable (transform).
 *   terms      joint log-density = sum over terms of sum over the term's n elements of
 *              lpdf_kind(value_i, params_i)  (kind POTENTIAL: just value_i).
 *   ops        per-term straight-line tape evaluated for element i = 0..n-1 in SSA registers.
 * ------------------------------------------------------------------------------------------ */

/* transforms (pymc4/distributions/transforms.py:137-181) */
#define PM4_TR_IDENTITY 0
#define PM4_TR_EXP 1      /* Log / LowerBound / UpperBound: x = shift + scale*exp(z)      */
#define PM4_TR_SIGMOID 2  /* Sigmoid / Interval:            x = shift + scale*sigmoid(z)  */

typedef struct pm4_rv {
  int32_t offset;    /* first position in theta                 */
  int32_t size;      /* number of scalars (prod of core shape)  */
  int32_t transform; /* PM4_TR_*                                */
  int32_t _pad;
  double shift, scale;
} pm4_rv_t;

/* term kinds (pymc4/distributions/continuous.py:49,176,620; discrete.py:32,452) */
#define PM4_T_POTENTIAL 0  /* value                                   */
#define PM4_T_NORMAL 1     /* value, loc, scale                       */
#define PM4_T_HALFNORMAL 2 /* value, scale                            */
#define PM4_T_HALFCAUCHY 3 /* value, scale  (loc = 0)                 */
#define PM4_T_BERNOULLI 4  /* value, probs                            */
#define PM4_T_POISSON 5    /* value, rate                             */
#define PM4_T_BERNOULLI_LOGIT_GLM 6 /* Bernoulli(probs = sigmoid(X beta)) observed: a dense term with no tape;
                                       X [n, glm_p] row-major at dataf[glm_x], y at dataf[glm_y], beta = x[glm_base ..]
                                       (evaluated for all chains at once on the tensor cores, csrc/pm4_glm.cu) */

#define PM4_T_MVN_EXPQUAD 7 /* y ~ MvNormal(loc, K) observed, K_ij = a^2 exp(-|X_i - X_j|^2 / (2 l^2)) + (noise + jitter) [i == j]:
                               the marginal likelihood of GP regression (/root/reference/pymc4/gp/cov.py:399-475,
                               distributions/multivariate.py:159-199), a dense term with no tape and n = 1 (one density
                               value per state).  The dense-term fields are reused: glm_base = number of points,
                               glm_p = feature dimension d, X [points, d] row-major at dataf[glm_x], y - loc at
                               dataf[glm_y]; dataf[gp_par .. gp_par + 9) = (l_elem, a_elem, noise_elem, l_const, a_const,
                               noise_const, noise_squared, jitter, loc_off): *_elem is the position in x (constrained
                               space) of the scalar free variable, or -1 to use the constant; the diagonal term is
                               (noise_squared ? v * v : v) + jitter; loc [points] at dataf[loc_off] (forward sampling).  Evaluated for all chains at once by batched
                               float64 Cholesky kernels (csrc/pm4_gp.cu). */

typedef struct pm4_term {
  int32_t kind;        /* PM4_T_*                                       */
  int32_t n;           /* number of elements                            */
  int32_t op_begin;    /* [op_begin, op_end) into ops                   */
  int32_t op_end;
  int32_t n_regs;      /* SSA registers used by the tape                */
  int32_t value_reg;   /* register holding the value                    */
  int32_t param_reg[3];/* registers holding the parameters (-1 unused)  */
  int32_t plan;        /* index into plans, or -1                       */
  int32_t observed;    /* 1: the term is the density of an observed variable (pointwise log-likelihood) */
  int32_t ll_offset;   /* position of element 0 inside one state's log-likelihood row (observed terms)  */
  int32_t glm_base;    /* PM4_T_BERNOULLI_LOGIT_GLM: theta offset of beta[0]                            */
  int32_t glm_p;       /*   number of covariates                                                        */
  int64_t glm_x;       /*   dataf offset of X [n, glm_p]                                                */
  int64_t glm_y;       /*   dataf offset of y [n]                                                       */
  int64_t gp_par;      /* PM4_T_MVN_EXPQUAD: dataf offset of the 9-double parameter record                */
} pm4_term_t;

/* Device execution plan of a term that gathers from free variables: its elements regrouped by
 * gather target into length-sorted chunks dealt 32*team_warps at a time to the lanes of the
 * chain's team of warps (pymc4_b200/plan.py).  Data-dependent, hence run-time pools; the CPU
 * oracle ignores plans. */
typedef struct pm4_plan {
  int32_t term;       /* the term it belongs to                        */
  int32_t n_gather;   /* gather loads sharing the plan's index array   */
  int32_t width;      /* reals per record: 1, 2 or 4                   */
  int32_t _pad;
  int64_t meta;       /* offset of the plan metadata in plani          */
  int64_t recs;       /* offset of the records in planf (in reals)     */
} pm4_plan_t;

/* tape opcodes */
#define PM4_OP_LD_CONST 0 /* dst = imm                                                     */
#define PM4_OP_LD_DATA 1  /* dst = dataf[blob + (mode==ELEM ? i : 0)]                      */
#define PM4_OP_LD_X 2     /* dst = x[base + (SCALAR:0 | ELEM:i | GATHER:datai[blob+i])]     */
#define PM4_OP_LD_Z 3     /* same addressing, reads theta                                   */
#define PM4_OP_ADD 10
#define PM4_OP_SUB 11
#define PM4_OP_MUL 12
#define PM4_OP_DIV 13
#define PM4_OP_POW 14
#define PM4_OP_NEG 20
#define PM4_OP_EXP 21
#define PM4_OP_LOG 22
#define PM4_OP_LOG1P 23
#define PM4_OP_SQRT 24
#define PM4_OP_SQUARE 25
#define PM4_OP_SIGMOID 26
#define PM4_OP_SOFTPLUS 27
#define PM4_OP_TANH 28
#define PM4_OP_ABS 29
#define PM4_OP_RECIP 30
#define PM4_OP_LGAMMA 31

#define PM4_MODE_SCALAR 0
#define PM4_MODE_ELEM 1
#define PM4_MODE_GATHER 2

typedef struct pm4_op {
  int32_t opcode;
  int32_t dst, a, b; /* SSA register numbers (-1 unused)          */
  int32_t mode;      /* PM4_MODE_* for loads                      */
  int32_t base;      /* theta/x offset for LD_X / LD_Z            */
  int64_t blob;      /* offset into dataf (LD_DATA) / datai (GATHER) */
  double imm;        /* LD_CONST                                  */
} pm4_op_t;

/* a traced output: deterministics and back-transformed variables
 * (/root/reference/pymc4/mcmc/samplers.py:796-809) */
typedef struct pm4_det {
  int32_t n;
  int32_t op_begin, op_end;
  int32_t n_regs;
  int32_t value_reg;
  int32_t out_offset; /* position of element 0 inside one draw's output row */
} pm4_det_t;

typedef struct pm4_ir {
  int32_t abi_version; /* PM4_ABI_VERSION */
  int32_t D;
  int32_t n_rv;
  int32_t n_terms;
  int32_t n_ops;
  int32_t n_dets;
  int32_t det_row;     /* scalars per draw produced by the dets tapes */
  int32_t _pad;
  const pm4_rv_t* rvs;
  const pm4_term_t* terms;
  const pm4_op_t* ops;
  const pm4_det_t* dets;
  int64_t n_dataf;
  const double* dataf;  /* constant real data pool (observed values, covariates) */
  int64_t n_datai;
  const int32_t* datai; /* constant index pool (gather indices)                  */
  uint64_t struct_hash; /* hash of everything except n, blob offsets and data    */
  int32_t n_plans;
  int32_t _pad2;
  const pm4_plan_t* plans;
  int64_t n_planf;
  const double* planf;  /* plan records (permuted copies of observed data), multiple of 4 */
  int64_t n_plani;
  const int32_t* plani; /* plan metadata, multiple of 4                                   */
  /* team layout: a chain is evaluated by `team_warps` warps; adjoint partial sums meet in a
   * per-chain `part` array described by two tables at the end of plani (pymc4_b200/plan.py) */
  int32_t team_warps;   /* 1, 2, 4, 8, 16 or 32; must match the kernel module              */
  int32_t part_reals;   /* length of the per-chain part array (multiple of 4)              */
  int64_t elem_off;     /* plani offset of elem_tab[32*team_warps*R]: overflow first << 12 | count */
  int64_t spart_off;    /* plani offset of spart[(n_scalar + 1) * team_warps]                */
  int64_t phdr_off;     /* plani offset of the per-plan int4 headers                         */
} pm4_ir_t;

typedef struct pm4_model pm4_model_t;

/* Create a model on `device` from the IR (copied; caller may free it afterwards) and bind the
 * specialised kernel module built for ir->struct_hash (`module_path`: a shared object produced
 * by pymc4_b200.codegen; it is dlopen'ed and its embedded hash is checked).  Synchronous. */
int pm4_model_create(const pm4_ir_t* ir, const char* module_path, int device, pm4_model_t** out);
int pm4_model_destroy(pm4_model_t* m);
int pm4_model_dim(const pm4_model_t* m);

/* Replace the constant data pools (same sizes) -- new observed values for the same model
 * (reference: pm.sample(..., observed={...}), samplers.py:85-108).  Synchronous. */
int pm4_model_set_data(pm4_model_t* m, const double* dataf, int64_t n_dataf,
                       const int32_t* datai, int64_t n_datai);

/* Companion of pm4_model_set_data: replace the execution-plan pools (ir->plans / planf / plani,
 * same sizes) after the observed data or gather indices changed.  Synchronous. */
int pm4_model_set_plans(pm4_model_t* m, const pm4_ir_t* ir);

/* Parity check (a): joint log-density and its gradient in unconstrained space for C chains.
 * Replaces parallel_logpfn + GradientTape (samplers.py:117-119, 762-770).
 *   theta_dev [C, D], lp_dev [C], grad_dev [C, D] (grad_dev may be NULL). */
int pm4_logp_grad(pm4_model_t* m, int dtype, int C, const void* theta_dev, void* lp_dev,
                  void* grad_dev, void* stream);

/* Traced outputs for a batch of states: theta_dev [M, D] -> out_dev [M, det_row]. */
int pm4_deterministics(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev,
                       void* out_dev, void* stream);

/* Pointwise log-likelihood of the observed variables for a batch of states
 * (calculate_log_likelihood, /root/reference/pymc4/mcmc/samplers.py:832-854: dist.log_prob(observed)
 * per draw and chain, no reduction): theta_dev [M, D] -> out_dev [M, ll_row], the observed terms'
 * elements side by side (pm4_term_t.ll_offset).  pm4_model_ll_row gives ll_row. */
int pm4_log_likelihood(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev,
                       void* out_dev, void* stream);
int pm4_model_ll_row(const pm4_model_t* m);

/* Posterior predictive: one draw of every observed variable from its distribution for every state
 * (sample_posterior_predictive, /root/reference/pymc4/forward_sampling.py:230-427: the model is
 * re-evaluated with the posterior values and the observed variables are sampled).
 * theta_dev [M, D] -> out_dev [M, ll_row] (row layout of pm4_log_likelihood).  Philox counters are
 * (state, element), key = seed: the same seed gives the same draws whatever the launch geometry. */
int pm4_posterior_predictive(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev, uint64_t seed,
                             void* out_dev, void* stream);

/* Prior predictive: M ancestral samples of the whole model (sample_prior_predictive,
 * /root/reference/pymc4/forward_sampling.py:26-227: the model evaluated with nothing given).
 * x_dev [M, D]: CONSTRAINED values of the free variables (theta layout); obs_dev [M, ll_row] or NULL:
 * draws of the observed variables (sample_from_observed). */
int pm4_prior_predictive(pm4_model_t* m, int dtype, int64_t M, uint64_t seed, void* x_dev, void* obs_dev,
                         void* stream);

/* Step-size policy (tfp DualAveragingStepSizeAdaptation) */
#define PM4_EPS_POOLED 0    /* one scalar shared by all chains (reference default, scalar step_size) */
#define PM4_EPS_PER_CHAIN 1 /* step_size carries a leading chain axis: independent adaptation       */

/* adaptation scheme: tfp DualAveragingStepSizeAdaptation (samplers "hmc", "nuts") or
 * tfp SimpleStepSizeAdaptation (samplers "hmc_simple", "nuts_simple", samplers.py:318-339, 422-447):
 * step *= (1 + adaptation_rate) when the mean acceptance exceeds the target, else /= (1 + rate) */
#define PM4_ADAPT_DUAL_AVERAGING 0
#define PM4_ADAPT_SIMPLE 1

typedef struct pm4_adapt_cfg {
  int32_t mode;                 /* PM4_EPS_*                                  */
  int32_t num_adaptation_steps; /* = burn_in (inference/sampling.py:155-156)  */
  double target_accept_prob;    /* 0.75                                       */
  double exploration_shrinkage; /* 0.05                                       */
  double step_count_smoothing;  /* 10                                         */
  double decay_rate;            /* 0.75                                       */
  int32_t enabled;              /* 0: fixed step size                         */
  int32_t kind;                 /* PM4_ADAPT_*                                */
  double adaptation_rate;       /* 0.01 (PM4_ADAPT_SIMPLE)                    */
  /* Diagonal mass-matrix adaptation (NUTS; SURVEY 8(f) item 4 -- NOT in the reference API: its radon notebook feeds
   * the posterior standard deviations in by hand as per-dimension step sizes, radon_hierarchical.ipynb:123-146).
   * K = mass_windows > 0: before transition w_k = num_adaptation_steps * k / (K + 1), k = 1 .. K, the per-dimension
   * step scale becomes the CROSS-CHAIN standard deviation of the unconstrained states (divided by its geometric mean)
   * and dual averaging restarts from the current epsilon.  A step scale s_d with unit-Gaussian momenta is the diagonal
   * mass matrix diag(1 / s_d^2).  0 = off (the reference behaviour). */
  int32_t mass_windows;
  int32_t _pad;
  double shrinkage_target;      /* tfp `shrinkage_target`: the step size dual averaging shrinks towards; 0 = the tfp default,
                                   10 x the initial step size                 */
} pm4_adapt_cfg_t;

/* proposals of the Metropolis-Hastings kernel run by pm4_hmc_run */
#define PM4_PROPOSAL_LEAPFROG 0     /* Hamiltonian Monte Carlo (samplers "hmc", "hmc_simple")                    */
#define PM4_PROPOSAL_RANDOM_WALK 1  /* tfp RandomWalkMetropolis with random_walk_normal_fn (sampler "rwm",
                                       samplers.py:450-478): state + scale * N(0, 1), gradient-free; the
                                       `momenta` argument then injects the perturbations               */

typedef struct pm4_hmc_cfg {
  int32_t C;                  /* chains                                                    */
  int32_t num_results;        /* S                                                         */
  int32_t num_burnin;         /* B : total transitions = B + S (tfp sample_chain, App. A.0) */
  int32_t num_leapfrog_steps; /* L (reference default 3, samplers.py:305)                  */
  double step_size;           /* initial epsilon                                           */
  uint64_t seed;
  pm4_adapt_cfg_t adapt;
  int32_t warps_per_block;    /* warps per thread block (a multiple of the model's team size); 0 = library default */
  int32_t chain_offset;       /* global index of chain 0 (RNG stream id) when chains are sharded */
  int32_t proposal;           /* PM4_PROPOSAL_*                                            */
  int32_t _pad;
  double rw_scale;            /* PM4_PROPOSAL_RANDOM_WALK: standard deviation of the normal perturbation (1.0) */
  const void* step_size_dev;  /* [C] reals of the call's dtype or NULL: initial epsilon of every chain (a `step_size` with a
                                 leading chain axis whose entries differ; per-chain adaptation mode only)          */
} pm4_hmc_cfg_t;

/* Fixed-L HMC (tfp HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC); App. A.4).
 *   theta0_dev [C, D]            initial state (every chain may differ)
 *   step_scale_dev [D] or NULL   per-dimension multiplier of epsilon
 *   momenta_dev [B+S, C, D] or NULL   injected momenta (parity check (b)); NULL = Philox
 *   uniforms_dev [B+S, C] or NULL     injected accept uniforms; NULL = Philox
 * outputs (any may be NULL): states_dev [S, C, D], lp_dev [S, C], log_accept_dev [S, C],
 *   accepted_dev [S, C] (u8), step_size_dev [C] final epsilon, traj_dev [B+S, C, D] the *proposed*
 *   end point of every trajectory (parity check (b) compares these). */
int pm4_hmc_run(pm4_model_t* m, int dtype, const pm4_hmc_cfg_t* cfg, const void* theta0_dev,
                const void* step_scale_dev, const void* momenta_dev, const void* uniforms_dev,
                void* states_dev, void* lp_dev, void* log_accept_dev, uint8_t* accepted_dev,
                void* step_size_dev, void* traj_dev, void* stream);

/* ---- compound step (discrete latents; reference pymc4/mcmc/samplers.py:490-726 CompoundStep, tf_support.py:52-163
 * _CompoundStepTF.one_step, proposals distributions/state_functions.py) -------------------------------------------
 * The free variables are partitioned into blocks; one compound transition runs ONE transition of every block's kernel
 * in order, each conditioned on the current values of all other dimensions (the reference's
 * _target_log_prob_fn_part_compound).  A block owns the dimensions whose `dim_scale` is non-zero: the other dimensions
 * are frozen (no momentum, no proposal).  Gradient blocks adapt their own step size; random-walk blocks may give every
 * dimension one of the proposals below (`rw_kind` NULL = the normal perturbation for all of them). */
#define PM4_BLOCK_NUTS 0
#define PM4_BLOCK_HMC 1
#define PM4_BLOCK_RWM 2
#define PM4_RW_NORMAL 0         /* x + rw_scale * N(0,1)         tfp random_walk_normal_fn (samplers.py:450-478)     */
#define PM4_RW_BERNOULLI 1      /* (x + Bernoulli(1/2)) mod 2    state_functions.py:118-146 bernoulli_fn             */
#define PM4_RW_CATEGORICAL 2    /* a fresh uniform class         state_functions.py:79-115 categorical_uniform_fn; classes in bits 8.. */
#define PM4_RW_GAUSSIAN_ROUND 3 /* round(x + rw_scale * N(0,1))  state_functions.py:149-196 gaussian_round_fn        */

typedef struct pm4_compound_block {
  int32_t kind;               /* PM4_BLOCK_*                                                     */
  int32_t num_leapfrog_steps; /* PM4_BLOCK_HMC                                                   */
  int32_t max_tree_depth;     /* PM4_BLOCK_NUTS                                                  */
  int32_t _pad;
  double max_energy_diff;     /* PM4_BLOCK_NUTS                                                  */
  double step_size;           /* gradient blocks: initial epsilon                                */
  double rw_scale;            /* PM4_BLOCK_RWM: standard deviation of the normal perturbation    */
  pm4_adapt_cfg_t adapt;      /* gradient blocks (mass_windows must be 0)                        */
  const void* dim_scale_dev;  /* [D] reals: 0 = not in this block, else per-dimension step scale */
  const int32_t* rw_kind_dev; /* [D] or NULL: PM4_RW_* | (classes << 8) of every dimension       */
} pm4_compound_block_t;

/* B + S compound transitions of C chains from theta0_dev [C, D].  Outputs (any may be NULL): states_dev [S, C, D],
 * lp_dev [S, C] joint log-density of the kept states, step_size_dev [n_blocks, C] final epsilon of every block (1 for
 * random-walk blocks), log_accept_dev [n_blocks, S, C] log acceptance ratio of every block's transition. */
int pm4_compound_run(pm4_model_t* m, int dtype, int32_t C, int32_t num_results, int32_t num_burnin, uint64_t seed,
                     int32_t chain_offset, int32_t warps_per_block, const pm4_compound_block_t* blocks, int32_t n_blocks,
                     const void* theta0_dev, void* states_dev, void* lp_dev, void* step_size_dev, void* log_accept_dev,
                     void* stream);

typedef struct pm4_nuts_cfg {
  int32_t C;
  int32_t num_results;     /* S */
  int32_t num_burnin;      /* B */
  int32_t max_tree_depth;  /* 10  (tfp NoUTurnSampler default)      */
  double max_energy_diff;  /* 1000                                   */
  double step_size;        /* initial epsilon (reference default 0.1, samplers.py:406) */
  uint64_t seed;
  pm4_adapt_cfg_t adapt;
  int32_t warps_per_block; /* 0 = library default */
  int32_t chain_offset;    /* global index of chain 0 (RNG stream id) when chains are sharded */
  const void* step_size_dev; /* [C] reals of the call's dtype or NULL: initial epsilon of every chain (see pm4_hmc_cfg_t) */
} pm4_nuts_cfg_t;

/* NUTS (tfp NoUTurnSampler: iterative tree doubling, generalised U-turn with checkpoints,
 * multinomial sampling, biased progressive merge) under dual averaging.
 * Replaces _BaseSampler._run_chains for sampler "nuts" (samplers.py:172-189, 342-419).
 *   theta0_dev [C, D]; step_scale_dev [D] or NULL; momenta_dev [B+S, C, D] or NULL (replay).
 * outputs (any may be NULL), all for the S kept draws, reference layout [S, C, ...]:
 *   states_dev [S,C,D], lp_dev [S,C], leapfrogs_dev [S,C] i32, diverging_dev [S,C] u8,
 *   energy_dev [S,C], log_accept_dev [S,C], depth_dev [S,C] i32, step_size_dev [C] final epsilon,
 *   total_leapfrogs_dev [C] i64: leapfrogs over ALL B+S transitions (throughput accounting).
 * Device memory of a run (every pm4_*_run entry): the caller owns the inputs and outputs; the run's scratch
 * (chain states, tree scratch, queues) comes from the device's default stream-ordered pool
 * (cudaMallocAsync / cudaFreeAsync on `stream`), is zeroed before use, and the library raises that pool's release
 * threshold to 256 MiB once per device so that the scratch of one run is still mapped for the next one. */
int pm4_nuts_run(pm4_model_t* m, int dtype, const pm4_nuts_cfg_t* cfg, const void* theta0_dev,
                 const void* step_scale_dev, const void* momenta_dev, void* states_dev,
                 void* lp_dev, int32_t* leapfrogs_dev, uint8_t* diverging_dev, void* energy_dev,
                 void* log_accept_dev, int32_t* depth_dev, void* step_size_dev,
                 int64_t* total_leapfrogs_dev, void* stream);

/* Per-dimension step scale [D] (unconstrained order) the last pm4_nuts_run with mass_windows > 0 ended with,
 * copied to out_dev (dtype real); PM4_EINVAL when no such run has happened. */
int pm4_model_get_mass_scale(pm4_model_t* m, int dtype, void* out_dev, void* stream);

/* Trace streaming (one-shot, consumed by the next pm4_nuts_run of a model without dense terms): the kept draws
 * are produced in `n_slabs` launches and every finished slab of `states` [S, C, D] is copied to `states_host`
 * (pinned host memory, same layout and dtype) on `copy_stream` while the next slab is sampled.  The reference's
 * _sample pulls the whole trace to the host at the end (`.numpy()`, samplers.py:155-160; trace_to_arviz,
 * mcmc/utils.py:95-144): for the radon shape that is 2.9 GB per call.  The caller synchronises `copy_stream`
 * before it reads the buffer.  states_host == NULL detaches. */
int pm4_model_set_trace_sink(pm4_model_t* m, void* states_host, int n_slabs, void* copy_stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU: chains of one pm.sample call sharded over the GPUs of one box, one process per GPU.
 * The sampler has ONE exchange step: with a scalar step_size (PM4_EPS_POOLED) tfp's
 * DualAveragingStepSizeAdaptation adapts one epsilon from the mean acceptance of ALL chains
 * (SURVEY App. A.3), i.e. one all-reduce of (sum, count) per burn-in transition.  It is fused into the
 * dual-averaging kernel: every rank owns a small mailbox in device memory, exported with CUDA IPC and
 * read by its peers over NVLink.  Protocol: pm4_comm_create on every rank -> exchange the 64-byte
 * handles (e.g. torch.distributed.all_gather) -> pm4_comm_connect -> pm4_model_attach_comm.
 * All ranks must then issue the same sequence of pm4_nuts_run / pm4_hmc_run calls.
 * ------------------------------------------------------------------------------------------ */
#define PM4_IPC_HANDLE_BYTES 64
typedef struct pm4_comm pm4_comm_t;
int pm4_comm_create(int device, int rank, int world, pm4_comm_t** out, unsigned char handle_out[PM4_IPC_HANDLE_BYTES]);
int pm4_comm_connect(pm4_comm_t* comm, const unsigned char* handles /* [world][PM4_IPC_HANDLE_BYTES] */);
int pm4_comm_destroy(pm4_comm_t* comm);
int pm4_model_attach_comm(pm4_model_t* m, pm4_comm_t* comm /* NULL detaches */);

/* ------------------------------------------------------------------------------------------
 * Dense Bernoulli-logit likelihood (PM4_T_BERNOULLI_LOGIT_GLM) for C chains at once, the building block
 * pm4_logp_grad and pm4_hmc_run call for such terms (csrc/pm4_glm.cu): two GEMMs on the tcgen05
 * tensor cores in 3xTF32, eta = X beta with the Bernoulli epilogue and g = X^T (y - sigmoid(eta)).
 *   X [n, P] row-major, Xt [P, ldxt] its transpose (ldxt >= n, multiple of 4), y [n] as float,
 *   theta [C, ldt] with beta at columns base .. base+P-1; adds into lp [C] and grad [C, ldg] (NULL: skip).
 *   P at most 104 (multiples of 4 take the 16-byte load path).  work: pm4_glm_work_bytes(n, P, C) bytes of device memory.
 *   *error_dev (device int, zeroed by the caller) is set when a tensor-core step timed out.
 * pm4_glm_gemm_tile: one 128 x 128 tile C = A B^T of the same MMA path (kernel-level test). */
int64_t pm4_glm_work_bytes(int n, int P, int C);

/* The marginal GP likelihood (PM4_T_MVN_EXPQUAD; /root/reference/pymc4/gp/cov.py:227-272,399-475, gp/util.py:15-21,
 * distributions/multivariate.py:159-199) for C chains at once, the building block pm4_logp_grad / pm4_hmc_run /
 * pm4_nuts_run / pm4_log_likelihood call for such terms (csrc/pm4_gp.cu): per chain the n x n kernel matrix, a
 * blocked Cholesky factorisation, the inverse factor and the contraction of (alpha alpha^T - K^-1) with dK, all in
 * float64.  X [n, d] and y [n] (y - loc) are DEVICE doubles; par is the HOST 8-double record described at
 * PM4_T_MVN_EXPQUAD; tr / shift / scale (HOST, 3 each, may be NULL = identity) are the transforms of the free
 * variables behind (l, a, noise).  dtype is the type of theta [C, ldt], lp and grad [C, ldg].  lp[c * lp_stride] is
 * overwritten (accumulate = 0) or added to; grad (may be NULL) is added to.  work: pm4_gp_work_bytes(n, C) bytes.
 * Returns 0, PM4_EINVAL or a cudaError_t. */
int64_t pm4_gp_work_bytes(int n, int C);
int pm4_gp_logp_grad(const double* X, const double* y, int n, int d, const double* par, const int* tr,
                     const double* shift, const double* scale, int dtype, const void* theta, int ldt, int C, void* lp,
                     int lp_stride, int accumulate, void* grad, int ldg, void* work, void* stream);
int pm4_gp_launches(int n, int want_grad);
/* Forward sampling of the GP term's observed vector: out[c, :] = loc + chol(K_c) z_c, z_c from the Philox stream of
 * (state0 + c, elem0 + i, seed).  loc [n]: DEVICE doubles; out [C, ldo] of `dtype`; other arguments as above. */
int pm4_gp_sample(const double* X, const double* loc, int n, int d, const double* par, const int* tr, const double* shift,
                  const double* scale, int dtype, const void* theta, int ldt, int C, int64_t state0, uint32_t elem0,
                  uint64_t seed, void* out, int ldo, void* work, void* stream);

/* Model-level wrapper: one draw of the observed vector of GP term `which` (0-based among the model's GP terms) for each
 * of the M states theta_dev [M, D] (unconstrained); out_dev [M, n_points].  Replaces, for MvNormal observations, the
 * resampling step of sample_posterior_predictive / sample_prior_predictive (/root/reference/pymc4/forward_sampling.py:
 * 26-427).  pm4_gp_points returns n_points of that term (or PM4_EINVAL). */
int pm4_gp_predictive(pm4_model_t* m, int dtype, int which, int64_t M, const void* theta_dev, uint64_t seed, void* out_dev,
                      void* stream);
int pm4_gp_points(const pm4_model_t* m, int which);
int pm4_glm_logp_grad(const float* X, const float* Xt, int ldxt, const float* y, int n, int P, const float* theta,
                      int ldt, int base, int C, float* lp, float* grad, int ldg, void* work, int* error_dev,
                      void* stream);
int pm4_glm_gemm_tile(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int K, int passes,
                      void* stream);

/* Fused evaluation of the same term (value + gradient in ONE kernel, csrc/pm4_glm.cu: glm_fused_kernel): the constant
 * operand is pre-split once into its two TF32 terms, tile by tile, in the shared-memory layout of both GEMMs
 * (pm4_glm_image_bytes / pm4_glm_build_image); per evaluation a warp-specialised kernel streams the tiles in with TMA
 * bulk copies, runs eta = beta X^T, the Bernoulli epilogue and G += R Xt^T back to back on the tensor cores -- the
 * residual matrix stays in shared memory.  pm4_logp_grad / pm4_hmc_run / pm4_nuts_run use it whenever the image fits in
 * device memory (PM4_GLM_FUSED=0 keeps the two-kernel path).  Same contract as pm4_glm_logp_grad; grad must not be NULL;
 * work: pm4_glm_fused_work_bytes(n, P, C) bytes.  gate_dev (device int or NULL): when it reads 0 at run time the call
 * does nothing -- lock-step NUTS enqueues several leapfrogs ahead and lets the unused ones fall through. */
int64_t pm4_glm_image_bytes(int n, int P);
int pm4_glm_build_image(const float* X, int n, int P, void* image, void* stream);
int64_t pm4_glm_fused_work_bytes(int n, int P, int C);
int pm4_glm_fused_logp_grad(const void* image, const float* y, int n, int P, const float* theta, int ldt, int base,
                            int C, float* lp, float* grad, int ldg, void* work, int* error_dev, const int* gate_dev,
                            void* stream);

/* Bytes of device scratch pm4_nuts_run / pm4_hmc_run allocate internally for C chains. */
int64_t pm4_scratch_bytes(const pm4_model_t* m, int dtype, int C, int max_tree_depth);

/* Number of kernels launched by this library on the calling thread since the last reset. */
int64_t pm4_launch_count(int reset);

const char* pm4_last_error(void);
int pm4_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PM4B200_H */
