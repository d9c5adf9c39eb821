// pm4_core.cu -- libpm4b200.so: the C ABI of include/pm4b200.h.
//
// Owns model handles (device copies of the constant pools + the dlopen'ed specialised kernel
// module) and drives the sampler launches: bootstrap, per-transition lock-step launches while a
// pooled step size is adapting (tfp DualAveragingStepSizeAdaptation with a scalar step size
// couples all chains every burn-in transition), otherwise a few persistent launches
// ("phases"); between phases the chains are re-ordered longest-first by the leapfrogs they took in
// the previous phase, so the persistent teams of the next launch start the expensive chains first.
// Replaces the hand-off _BaseSampler._run_chains -> tfp.mcmc.sample_chain
// (/root/reference/pymc4/mcmc/samplers.py:172-189).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/pm4b200.h"
#include "pm4_lockstep.cuh"
#include "pm4_module.h"

namespace {

// NVTX range around a C-ABI call (header-only NVTX3: a no-op unless a profiler injects itself)
struct Nvtx {
  explicit Nvtx(const char* name) { nvtxRangePushA(name); }
  ~Nvtx() { nvtxRangePop(); }
};

thread_local std::string g_err;
thread_local int64_t g_launches = 0;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(expr)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess) return fail(PM4_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

#define MOD(expr, what)                                                                                \
  do {                                                                                                 \
    int rc_ = (expr);                                                                                  \
    ++g_launches;                                                                                      \
    if (rc_ == -38) return fail(PM4_ENOSYS, "%s: dtype not compiled into this module", what);          \
    if (rc_ != 0) return fail(PM4_ECUDA, "%s: launch failed: %s", what, cudaGetErrorString((cudaError_t)rc_)); \
  } while (0)

}  // namespace

struct pm4_comm {
  int device = 0, rank = 0, world = 1;
  void* box[PM4_MAX_RANKS] = {nullptr};  // box[rank] is ours (cudaMalloc), the others are IPC mappings
  int32_t* d_error = nullptr;
  uint32_t epoch = 0;                    // one per sampling call: stale mailbox entries never match
  bool connected = false;
};

// a dense Bernoulli-logit term (PM4_T_BERNOULLI_LOGIT_GLM): pool offsets + the transposed covariates
struct GlmTerm {
  int n = 0, P = 0, base = 0, ldxt = 0;
  int64_t x_off = 0, y_off = 0;
  float* d_xt = nullptr;  // [P, ldxt]
  void* d_img = nullptr;   // pre-split tile image of X for the fused kernel (pm4_glm_build_image), or NULL
};

// a marginal GP likelihood term (PM4_T_MVN_EXPQUAD): pool offsets + how its three scalars are read from theta
struct GpTerm {
  int n = 0, d = 0, ll_offset = 0;
  int64_t x_off = 0, y_off = 0, loc_off = 0;
  double par[8] = {0};
  int tr[3] = {0, 0, 0};
  double shift[3] = {0, 0, 0}, scale[3] = {1, 1, 1};
};

struct pm4_model {
  int device = 0;
  pm4_comm* comm = nullptr;
  std::vector<GlmTerm> glm;
  std::vector<GpTerm> gp;
  void* gp_work = nullptr;
  size_t gp_work_bytes = 0;
  bool lockstep() const { return !glm.empty() || !gp.empty(); }
  void* glm_work = nullptr;
  size_t glm_work_bytes = 0;
  int* d_glm_err = nullptr;
  int* h_flag = nullptr;  // pinned: chains still growing their tree (lock-step NUTS)
  int D = 0, n_terms = 0, n_ops = 0, n_dets = 0, det_row = 0;
  int64_t n_dataf = 0, n_datai = 0;
  std::vector<double> dataf;
  void* module = nullptr;
  pm4_mod_info_t info{};
  pm4m_logp_grad_fn f_logp = nullptr;
  pm4m_dets_fn f_dets = nullptr;
  pm4m_loglik_fn f_loglik = nullptr;
  pm4m_predict_fn f_predict = nullptr;
  pm4m_prior_fn f_prior = nullptr;
  pm4m_init_fn f_init = nullptr;
  pm4m_run_fn f_nuts = nullptr, f_hmc = nullptr;
  pm4m_da_pooled_fn f_da = nullptr;
  pm4m_export_eps_fn f_eps = nullptr;
  pm4m_scratch_bytes_fn f_scratch = nullptr;
  pm4m_lpt_order_fn f_order = nullptr;
  int elem_off = 0, spart_off = 0, phdr_off = 0;
  int32_t owned_hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // one-shot trace sink of the next NUTS run (pm4_model_set_trace_sink): the draws leave the device in slabs
  // on a copy stream while the persistent kernel keeps sampling
  void* mass_scale = nullptr;  // [D] reals: step scale the last run with mass windows ended with
  int mass_scale_dtype = -1;
  void* sink_host = nullptr;
  int sink_slabs = 0;
  cudaStream_t sink_stream = nullptr;
  cudaEvent_t sink_event = nullptr;
  float* d_f32 = nullptr;
  double* d_f64 = nullptr;
  int32_t *d_datai = nullptr, *d_term_n = nullptr, *d_op_blob = nullptr;
  // execution plans (pymc4_b200/plan.py)
  int n_plans = 0, part_reals = 0;
  int64_t n_planf = 0, n_plani = 0;
  float* d_pf32 = nullptr;
  double* d_pf64 = nullptr;
  int32_t *d_plani = nullptr, *d_plan_off = nullptr;

  pm4_mod_args_t args(int dtype) const {
    pm4_mod_args_t a;
    a.dataf = dtype == PM4_F64 ? (const void*)d_f64 : (const void*)d_f32;
    a.datai = d_datai;
    a.term_n = d_term_n;
    a.op_blob = d_op_blob;
    a.planf = dtype == PM4_F64 ? (const void*)d_pf64 : (const void*)d_pf32;
    a.plani = d_plani;
    a.plan_off = d_plan_off;
    a.n_planf = (int32_t)n_planf;
    a.n_plani = (int32_t)n_plani;
    a.part_reals = part_reals;
    a.elem_off = elem_off;
    a.spart_off = spart_off;
    a.phdr_off = phdr_off;
    // stage the plan pools in shared memory when they leave room for the per-chain state
    const size_t pool_bytes = (size_t)n_planf * (dtype == PM4_F64 ? 8 : 4) + (size_t)n_plani * 4;
    a.stage = (pool_bytes <= 96 * 1024) ? 1 : 0;
    memcpy(a.owned, owned_hdr, sizeof a.owned);
    return a;
  }
};

static void fill_comm(pm4_model* m, pm4_mod_run_t& r);
static int check_comm(pm4_model* m, cudaStream_t st);

static int upload_plans(pm4_model* m, const pm4_ir_t* ir) {
  const int64_t nf = ir->n_planf, ni = ir->n_plani;
  if (nf % 4 || ni % 4) return fail(PM4_EINVAL, "plan pools must be padded to multiples of 4 elements");
  // validate everything before the first byte is copied: a rejected IR leaves the model untouched
  for (int k = 0; k < ir->n_plans; ++k) {
    const pm4_plan_t& pl = ir->plans[k];
    if (pl.meta < 0 || pl.meta >= ni || pl.recs < 0 || pl.recs > nf) return fail(PM4_EINVAL, "plan %d points outside the plan pools", k);
  }
  if (ir->team_warps != m->info.T) return fail(PM4_EINVAL, "IR is laid out for teams of %d warps, the kernel module for %d", ir->team_warps, m->info.T);
  if (ir->elem_off < 0 || ir->elem_off + m->info.DP > ni || ir->spart_off < 0 || ir->spart_off >= ni ||
      ir->phdr_off < 0 || ir->phdr_off + 4 * ir->n_plans > ni || ir->part_reals <= 0)
    return fail(PM4_EINVAL, "part-array tables point outside the plan metadata pool");
  {
    const int64_t oh = ir->phdr_off + 4 * (ir->n_plans > 0 ? ir->n_plans : 1);
    if (oh + 8 <= ni && ir->plani[oh] == 1) {
      const int32_t* h = ir->plani + oh;
      if (h[1] < 0 || h[2] < 0 || h[1] + (int64_t)h[2] > ni || h[3] < 0 || h[4] < 0 || h[3] + (int64_t)h[4] > nf || (h[2] % 4) || (h[4] % 4) ||
          (h[1] % 4) || (h[3] % 4))
        return fail(PM4_EINVAL, "owned plan sections point outside the plan pools");
    }
  }
  std::vector<float> f32((size_t)nf);
  for (int64_t k = 0; k < nf; ++k) f32[(size_t)k] = (float)ir->planf[k];
  if (nf) {
    CU(cudaMemcpy(m->d_pf32, f32.data(), sizeof(float) * (size_t)nf, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->d_pf64, ir->planf, sizeof(double) * (size_t)nf, cudaMemcpyHostToDevice));
  }
  if (ni) CU(cudaMemcpy(m->d_plani, ir->plani, sizeof(int32_t) * (size_t)ni, cudaMemcpyHostToDevice));
  std::vector<int32_t> off((size_t)(2 * ir->n_plans) + 2);
  for (int k = 0; k < ir->n_plans; ++k) {
    const pm4_plan_t& pl = ir->plans[k];
    if (pl.meta < 0 || pl.meta >= ni || pl.recs < 0 || pl.recs > nf) return fail(PM4_EINVAL, "plan %d points outside the plan pools", k);
    off[(size_t)(2 * k)] = (int32_t)pl.meta;
    off[(size_t)(2 * k + 1)] = (int32_t)pl.recs;
  }
  CU(cudaMemcpy(m->d_plan_off, off.data(), sizeof(int32_t) * off.size(), cudaMemcpyHostToDevice));
  if (ir->team_warps != m->info.T) return fail(PM4_EINVAL, "IR is laid out for teams of %d warps, the kernel module for %d", ir->team_warps, m->info.T);
  if (ir->elem_off < 0 || ir->elem_off + m->info.DP > ni || ir->spart_off < 0 || ir->spart_off >= ni ||
      ir->phdr_off < 0 || ir->phdr_off + 4 * ir->n_plans > ni || ir->part_reals <= 0)
    return fail(PM4_EINVAL, "part-array tables point outside the plan metadata pool");
  m->part_reals = ir->part_reals;
  m->elem_off = (int)ir->elem_off;
  m->spart_off = (int)ir->spart_off;
  m->phdr_off = (int)ir->phdr_off;
  // owner-computes plan of the warp-per-chain path (pymc4_b200/owned.py): 8 ints behind the per-plan headers
  memset(m->owned_hdr, 0, sizeof m->owned_hdr);
  const int64_t oh = ir->phdr_off + 4 * (ir->n_plans > 0 ? ir->n_plans : 1);
  if (oh + 8 <= ni) {
    const int32_t* h = ir->plani + oh;
    if (h[0] == 1) {
      if (h[1] < 0 || h[2] < 0 || h[1] + (int64_t)h[2] > ni || h[3] < 0 || h[4] < 0 || h[3] + (int64_t)h[4] > nf || (h[2] % 4) || (h[4] % 4) ||
          (h[1] % 4) || (h[3] % 4))
        return fail(PM4_EINVAL, "owned plan sections point outside the plan pools");
      memcpy(m->owned_hdr, h, sizeof m->owned_hdr);
    }
  }
  return PM4_OK;
}

// Xt[q, i] = X[i, q] through a 32 x 33 shared-memory tile (both sides coalesced); columns i >= n stay zero
__global__ void transpose_kernel(const float* __restrict__ X, int n, int P, float* __restrict__ Xt, int ldxt) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, q0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int i = i0 + r, q = q0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && q < P) ? X[(size_t)i * P + q] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int q = q0 + r, i = i0 + threadIdx.x;
    if (q < P && i < ldxt) Xt[(size_t)q * ldxt + i] = tile[threadIdx.x][r];
  }
}

// transposed single-precision copy of every GLM term's covariates (the gradient GEMM contracts over observations),
// made on the device from the pool upload_pools has just written
static int upload_glm(pm4_model* m, const double*) {
  for (GlmTerm& g : m->glm) {
    if (!g.d_xt) CU(cudaMalloc(&g.d_xt, sizeof(float) * (size_t)g.P * g.ldxt));
    transpose_kernel<<<dim3((unsigned)((g.ldxt + 31) / 32), (unsigned)((g.P + 31) / 32)), dim3(32, 8)>>>(
        m->d_f32 + g.x_off, g.n, g.P, g.d_xt, g.ldxt);
    CU(cudaGetLastError());
    // pre-split image for the fused kernel (2 x the bytes of X and of X^T): built when it fits comfortably
    static const bool fused_off = getenv("PM4_GLM_FUSED") && atoi(getenv("PM4_GLM_FUSED")) == 0;
    const int64_t img_bytes = fused_off ? -1 : pm4_glm_image_bytes(g.n, g.P);
    if (img_bytes > 0 && !g.d_img) {
      size_t free_b = 0, total_b = 0;
      CU(cudaMemGetInfo(&free_b, &total_b));
      if ((size_t)img_bytes < free_b / 2 && cudaMalloc(&g.d_img, (size_t)img_bytes) != cudaSuccess) {
        g.d_img = nullptr;
        cudaGetLastError();
      }
    }
    if (g.d_img) {
      int rc = pm4_glm_build_image(m->d_f32 + g.x_off, g.n, g.P, g.d_img, nullptr);
      if (rc != 0) return fail(PM4_ECUDA, "GLM image kernel: %s", cudaGetErrorString((cudaError_t)rc));
    }
  }
  CU(cudaDeviceSynchronize());
  return PM4_OK;
}

// grow the GP workspace to hold all C chains when it can (60 % of what is free, or PM4_GP_WORK_MB); otherwise the
// chains go through it in chunks
static int gp_workspace(pm4_model* m, const GpTerm& g, int C, cudaStream_t st) {
  const size_t per_chain = (size_t)pm4_gp_work_bytes(g.n, 1);
  if (m->gp_work_bytes < per_chain * (size_t)C) {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    size_t cap = (free_b + m->gp_work_bytes) / 10 * 6;
    if (const char* e = getenv("PM4_GP_WORK_MB")) cap = (size_t)atoll(e) << 20;
    size_t want = std::min(per_chain * (size_t)C, cap / per_chain * per_chain);
    if (want > m->gp_work_bytes) {
      CU(cudaStreamSynchronize(st));
      if (m->gp_work) CU(cudaFree(m->gp_work));
      m->gp_work = nullptr; m->gp_work_bytes = 0;
      if (want < per_chain) return fail(PM4_ENOMEM, "GP term with %d points needs %zu bytes of workspace per chain, %zu are free", g.n, per_chain, free_b);
      cudaError_t e = cudaMalloc(&m->gp_work, want);
      if (e != cudaSuccess) return fail(PM4_ENOMEM, "cudaMalloc(%zu) for the GP factorisations: %s", want, cudaGetErrorString(e));
      m->gp_work_bytes = want;
    }
  }
  if (m->gp_work_bytes < per_chain) return fail(PM4_ENOMEM, "no workspace for a GP term with %d points", g.n);
  return PM4_OK;
}

// add the GP terms for C chains: batched float64 Cholesky kernels (csrc/pm4_gp.cu), chains in chunks that fit
// the workspace (2 n^2 doubles per chain)
static int add_gp(pm4_model* m, int dtype, int C, const void* theta, void* lp, int lp_stride, bool accumulate, void* grad,
                  cudaStream_t st) {
  const size_t rs = dtype == PM4_F64 ? 8 : 4;
  for (const GpTerm& g : m->gp) {
    const size_t per_chain = (size_t)pm4_gp_work_bytes(g.n, 1);
    if (int rcw = gp_workspace(m, g, C, st)) return rcw;
    const int chunk = (int)std::min<size_t>((size_t)C, m->gp_work_bytes / per_chain);
    for (int c0 = 0; c0 < C; c0 += chunk) {
      const int cc = std::min(chunk, C - c0);
      int rc = pm4_gp_logp_grad(m->d_f64 + g.x_off, m->d_f64 + g.y_off, g.n, g.d, g.par, g.tr, g.shift, g.scale, dtype,
                                (const char*)theta + (size_t)c0 * m->D * rs, m->D, cc,
                                (char*)lp + ((size_t)c0 * lp_stride + (accumulate ? 0 : (size_t)g.ll_offset)) * rs, lp_stride,
                                accumulate ? 1 : 0,
                                grad ? (char*)grad + (size_t)c0 * m->D * rs : nullptr, m->D, m->gp_work, st);
      g_launches += pm4_gp_launches(g.n, grad ? 1 : 0);
      if (rc != 0) return fail(rc < 0 ? PM4_EINVAL : PM4_ECUDA, "GP kernels: %s", rc < 0 ? "bad argument" : cudaGetErrorString((cudaError_t)rc));
    }
  }
  return PM4_OK;
}

// add the dense terms' log-likelihood and gradient for C chains (theta [C, D] -> lp [C], grad [C, D] or NULL)
static int add_glm(pm4_model* m, int dtype, int C, const void* theta, void* lp, void* grad, cudaStream_t st,
                   const int* gate = nullptr) {
  if (!m->gp.empty()) {
    int rc = add_gp(m, dtype, C, theta, lp, 1, true, grad, st);
    if (rc) return rc;
  }
  if (m->glm.empty()) return PM4_OK;
  if (dtype != PM4_F32) return fail(PM4_ENOSYS, "dense GLM terms are evaluated in 3xTF32 on the tensor cores: float32 only");
  for (const GlmTerm& g : m->glm) {
    const bool fused = g.d_img && grad;
    const size_t need = (size_t)(fused ? pm4_glm_fused_work_bytes(g.n, g.P, C) : pm4_glm_work_bytes(g.n, g.P, C));
    if (need > m->glm_work_bytes) {
      CU(cudaStreamSynchronize(st));
      if (m->glm_work) CU(cudaFree(m->glm_work));
      m->glm_work = nullptr; m->glm_work_bytes = 0;
      cudaError_t e = cudaMalloc(&m->glm_work, need);
      if (e != cudaSuccess) return fail(PM4_ENOMEM, "cudaMalloc(%zu) for the GLM residuals: %s", need, cudaGetErrorString(e));
      m->glm_work_bytes = need;
    }
    if (gate && !fused) return fail(PM4_EINVAL, "gated evaluation needs the fused GLM path");
    int rc = fused ? pm4_glm_fused_logp_grad(g.d_img, m->d_f32 + g.y_off, g.n, g.P, (const float*)theta, m->D, g.base, C,
                                             (float*)lp, (float*)grad, m->D, m->glm_work, m->d_glm_err, gate, st)
                   : pm4_glm_logp_grad(m->d_f32 + g.x_off, g.d_xt, g.ldxt, m->d_f32 + g.y_off, g.n, g.P, (const float*)theta, m->D,
                                       g.base, C, (float*)lp, (float*)grad, m->D, m->glm_work, m->d_glm_err, st);
    g_launches += fused ? 2 : (grad ? 4 : 3);
    if (rc != 0) return fail(rc < 0 ? PM4_EINVAL : PM4_ECUDA, "GLM kernels: %s", rc < 0 ? "unsupported shape (at most 104 covariates)" : cudaGetErrorString((cudaError_t)rc));
  }
  return PM4_OK;
}

// did a tensor-core step time out?  (synchronises the stream)
static int check_glm(pm4_model* m, cudaStream_t st) {
  if (m->glm.empty()) return PM4_OK;
  int err = 0;
  CU(cudaMemcpyAsync(&err, m->d_glm_err, sizeof err, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (err) {
    cudaMemsetAsync(m->d_glm_err, 0, sizeof(int), st);
    return fail(PM4_ECUDA, "a tcgen05 step of the GLM kernels timed out (code %d)", err);
  }
  return PM4_OK;
}

static int upload_pools(pm4_model* m, const double* dataf, int64_t nf, const int32_t* datai, int64_t ni) {
  m->dataf.assign(dataf, dataf + nf);
  std::vector<float> f32((size_t)nf);
  for (int64_t k = 0; k < nf; ++k) f32[(size_t)k] = (float)dataf[k];
  if (nf) {
    CU(cudaMemcpy(m->d_f32, f32.data(), sizeof(float) * (size_t)nf, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->d_f64, dataf, sizeof(double) * (size_t)nf, cudaMemcpyHostToDevice));
  }
  if (ni) CU(cudaMemcpy(m->d_datai, datai, sizeof(int32_t) * (size_t)ni, cudaMemcpyHostToDevice));
  return PM4_OK;
}

extern "C" int pm4_model_create(const pm4_ir_t* ir, const char* module_path, int device, pm4_model_t** out) {
  Nvtx nvtx_range("pm4_model_create");
  if (!ir || !module_path || !out) return fail(PM4_EINVAL, "pm4_model_create: null argument");
  if (ir->abi_version != PM4_ABI_VERSION) return fail(PM4_EINVAL, "IR ABI version %d != %d", ir->abi_version, PM4_ABI_VERSION);
  if (ir->D <= 0 || ir->n_terms <= 0) return fail(PM4_EINVAL, "IR has no free variables or no terms");
  for (int t = 0; t < ir->n_terms; ++t) {
    const pm4_term_t& tm = ir->terms[t];
    if (tm.op_begin < 0 || tm.op_end > ir->n_ops || tm.op_begin > tm.op_end || tm.n < 0)
      return fail(PM4_EINVAL, "term %d has a malformed tape range", t);
  }
  for (int k = 0; k < ir->n_ops; ++k) {
    const pm4_op_t& o = ir->ops[k];
    const bool is_ld = o.opcode == PM4_OP_LD_X || o.opcode == PM4_OP_LD_Z;
    if (is_ld && (o.base < 0 || o.base >= ir->D)) return fail(PM4_EINVAL, "op %d reads outside theta", k);
    if (o.opcode == PM4_OP_LD_DATA && (o.blob < 0 || o.blob >= ir->n_dataf)) return fail(PM4_EINVAL, "op %d reads outside the real pool", k);
    if (is_ld && o.mode == PM4_MODE_GATHER && (o.blob < 0 || o.blob >= ir->n_datai)) return fail(PM4_EINVAL, "op %d reads outside the index pool", k);
  }
  pm4_model* m = new pm4_model();
  m->device = device;
  m->D = ir->D; m->n_terms = ir->n_terms; m->n_ops = ir->n_ops; m->n_dets = ir->n_dets; m->det_row = ir->det_row;
  m->n_dataf = ir->n_dataf; m->n_datai = ir->n_datai;

  m->module = dlopen(module_path, RTLD_NOW | RTLD_LOCAL);
  if (!m->module) {
    int rc = fail(PM4_ENOMOD, "cannot load kernel module %s: %s", module_path, dlerror());
    delete m;
    return rc;
  }
  pm4m_info_fn f_info = (pm4m_info_fn)dlsym(m->module, "pm4m_info");
  m->f_logp = (pm4m_logp_grad_fn)dlsym(m->module, "pm4m_logp_grad");
  m->f_dets = (pm4m_dets_fn)dlsym(m->module, "pm4m_dets");
  m->f_loglik = (pm4m_loglik_fn)dlsym(m->module, "pm4m_loglik");
  m->f_predict = (pm4m_predict_fn)dlsym(m->module, "pm4m_predict");
  m->f_prior = (pm4m_prior_fn)dlsym(m->module, "pm4m_prior");
  m->f_init = (pm4m_init_fn)dlsym(m->module, "pm4m_init");
  m->f_nuts = (pm4m_run_fn)dlsym(m->module, "pm4m_nuts");
  m->f_hmc = (pm4m_run_fn)dlsym(m->module, "pm4m_hmc");
  m->f_da = (pm4m_da_pooled_fn)dlsym(m->module, "pm4m_da_pooled");
  m->f_eps = (pm4m_export_eps_fn)dlsym(m->module, "pm4m_export_eps");
  m->f_scratch = (pm4m_scratch_bytes_fn)dlsym(m->module, "pm4m_scratch_bytes");
  m->f_order = (pm4m_lpt_order_fn)dlsym(m->module, "pm4m_lpt_order");
  if (!f_info || !m->f_logp || !m->f_dets || !m->f_init || !m->f_nuts || !m->f_hmc || !m->f_da || !m->f_eps ||
      !m->f_scratch || !m->f_order || !m->f_loglik || !m->f_predict || !m->f_prior) {
    int rc = fail(PM4_ENOMOD, "kernel module %s lacks a required entry point", module_path);
    pm4_model_destroy(m);
    return rc;
  }
  f_info(&m->info);
  int64_t ir_ll_row = 0;  // the module bakes the pointwise log-likelihood row layout in: it must be the IR's
  for (int k = 0; k < ir->n_terms; ++k)
    if (ir->terms[k].observed) ir_ll_row += ir->terms[k].n;
  if (m->info.abi != PM4_MODULE_ABI || m->info.struct_hash != ir->struct_hash || m->info.D != ir->D ||
      m->info.n_terms != ir->n_terms || m->info.n_ops != ir->n_ops || m->info.det_row != ir->det_row ||
      (int64_t)m->info.ll_row != ir_ll_row) {
    int rc = fail(PM4_ENOMOD, "kernel module %s was built for another model structure (hash %016llx, IR %016llx)",
                  module_path, (unsigned long long)m->info.struct_hash, (unsigned long long)ir->struct_hash);
    pm4_model_destroy(m);
    return rc;
  }

  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    int rc = fail(PM4_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    pm4_model_destroy(m);
    return rc;
  }
  std::vector<int32_t> term_n((size_t)(ir->n_terms + ir->n_dets) + 1), op_blob((size_t)ir->n_ops + 1);
  for (int t = 0; t < ir->n_terms; ++t) term_n[(size_t)t] = ir->terms[t].n;
  for (int d = 0; d < ir->n_dets; ++d) term_n[(size_t)(ir->n_terms + d)] = ir->dets[d].n;
  for (int k = 0; k < ir->n_ops; ++k) op_blob[(size_t)k] = (int32_t)ir->ops[k].blob;
  auto alloc = [&]() -> int {
    CU(cudaMalloc(&m->d_f32, sizeof(float) * (size_t)(ir->n_dataf + 4)));
    CU(cudaMalloc(&m->d_f64, sizeof(double) * (size_t)(ir->n_dataf + 4)));
    CU(cudaMalloc(&m->d_datai, sizeof(int32_t) * (size_t)(ir->n_datai + 4)));
    CU(cudaMalloc(&m->d_term_n, sizeof(int32_t) * term_n.size()));
    CU(cudaMalloc(&m->d_op_blob, sizeof(int32_t) * op_blob.size()));
    CU(cudaMemcpy(m->d_term_n, term_n.data(), sizeof(int32_t) * term_n.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m->d_op_blob, op_blob.data(), sizeof(int32_t) * op_blob.size(), cudaMemcpyHostToDevice));
    m->n_plans = ir->n_plans; m->n_planf = ir->n_planf; m->n_plani = ir->n_plani;
    CU(cudaMalloc(&m->d_pf32, sizeof(float) * (size_t)(ir->n_planf + 4)));
    CU(cudaMalloc(&m->d_pf64, sizeof(double) * (size_t)(ir->n_planf + 4)));
    CU(cudaMalloc(&m->d_plani, sizeof(int32_t) * (size_t)(ir->n_plani + 4)));
    CU(cudaMalloc(&m->d_plan_off, sizeof(int32_t) * (size_t)(2 * ir->n_plans + 2)));
    int rcp = upload_plans(m, ir);
    if (rcp != PM4_OK) return rcp;
    for (int t = 0; t < ir->n_terms; ++t) {
      const pm4_term_t& tm = ir->terms[t];
      if (tm.kind != PM4_T_BERNOULLI_LOGIT_GLM) continue;
      if (tm.glm_p <= 0 || tm.glm_p > 104) return fail(PM4_ENOSYS, "GLM term %d: at most 104 covariates fit the shared-memory operand tiles (got %d)", t, tm.glm_p);
      if (tm.glm_base < 0 || tm.glm_base + tm.glm_p > ir->D || tm.glm_x < 0 || tm.glm_x + (int64_t)tm.n * tm.glm_p > ir->n_dataf ||
          tm.glm_y < 0 || tm.glm_y + tm.n > ir->n_dataf)
        return fail(PM4_EINVAL, "GLM term %d points outside theta or the data pool", t);
      GlmTerm g;
      g.n = tm.n; g.P = tm.glm_p; g.base = tm.glm_base; g.x_off = tm.glm_x; g.y_off = tm.glm_y;
      g.ldxt = (tm.n + 3) / 4 * 4;
      m->glm.push_back(g);
    }
    for (int t = 0; t < ir->n_terms; ++t) {
      const pm4_term_t& tm = ir->terms[t];
      if (tm.kind != PM4_T_MVN_EXPQUAD) continue;
      const int n = tm.glm_base, d = tm.glm_p;
      if (n <= 0 || d <= 0 || tm.glm_x < 0 || tm.glm_x + (int64_t)n * d > ir->n_dataf || tm.glm_y < 0 ||
          tm.glm_y + n > ir->n_dataf || tm.gp_par < 0 || tm.gp_par + 9 > ir->n_dataf)
        return fail(PM4_EINVAL, "GP term %d points outside the data pool", t);
      const int64_t loc_off = (int64_t)ir->dataf[tm.gp_par + 8];
      if (loc_off < 0 || loc_off + n > ir->n_dataf) return fail(PM4_EINVAL, "GP term %d: loc points outside the data pool", t);
      GpTerm g;
      g.n = n; g.d = d; g.ll_offset = tm.ll_offset; g.x_off = tm.glm_x; g.y_off = tm.glm_y; g.loc_off = loc_off;
      for (int k = 0; k < 8; ++k) g.par[k] = ir->dataf[tm.gp_par + k];
      for (int k = 0; k < 3; ++k) {
        const int e = (int)g.par[k];
        if (e < 0) continue;
        bool found = false;
        for (int r = 0; r < ir->n_rv && !found; ++r) {
          const pm4_rv_t& rv = ir->rvs[r];
          if (e >= rv.offset && e < rv.offset + rv.size) {
            g.tr[k] = rv.transform; g.shift[k] = rv.shift; g.scale[k] = rv.scale;
            found = true;
          }
        }
        if (!found) return fail(PM4_EINVAL, "GP term %d: parameter %d is not a position of theta", t, k);
      }
      m->gp.push_back(g);
    }
    if (!m->glm.empty()) {
      CU(cudaMalloc((void**)&m->d_glm_err, sizeof(int)));
      CU(cudaMemset(m->d_glm_err, 0, sizeof(int)));
    }
    rcp = upload_pools(m, ir->dataf, ir->n_dataf, ir->datai, ir->n_datai);
    if (rcp != PM4_OK) return rcp;
    return upload_glm(m, ir->dataf);
  };
  int rc = alloc();
  if (rc != PM4_OK) {
    std::string keep = g_err;
    pm4_model_destroy(m);
    g_err = keep;
    return rc;
  }
  *out = m;
  return PM4_OK;
}

extern "C" int pm4_model_destroy(pm4_model_t* m) {
  if (!m) return PM4_OK;
  cudaFree(m->d_f32); cudaFree(m->d_f64); cudaFree(m->d_datai); cudaFree(m->d_term_n); cudaFree(m->d_op_blob);
  cudaFree(m->d_pf32); cudaFree(m->d_pf64); cudaFree(m->d_plani); cudaFree(m->d_plan_off);
  for (GlmTerm& g : m->glm) { cudaFree(g.d_xt); cudaFree(g.d_img); }
  cudaFree(m->glm_work); cudaFree(m->d_glm_err); cudaFree(m->gp_work);
  cudaFree(m->mass_scale);
  if (m->sink_event) cudaEventDestroy(m->sink_event);
  if (m->h_flag) cudaFreeHost(m->h_flag);
  if (m->module) dlclose(m->module);
  delete m;
  return PM4_OK;
}

extern "C" int pm4_model_dim(const pm4_model_t* m) { return m ? m->D : PM4_EINVAL; }

extern "C" int pm4_model_set_data(pm4_model_t* m, const double* dataf, int64_t n_dataf, const int32_t* datai, int64_t n_datai) {
  Nvtx nvtx_range("pm4_model_set_data");
  if (!m) return fail(PM4_EINVAL, "null model");
  if (n_dataf != m->n_dataf || n_datai != m->n_datai)
    return fail(PM4_EINVAL, "pool sizes differ from the model's (%lld/%lld vs %lld/%lld)", (long long)n_dataf,
                (long long)n_datai, (long long)m->n_dataf, (long long)m->n_datai);
  CU(cudaSetDevice(m->device));
  CU(cudaDeviceSynchronize());
  int rc = upload_pools(m, dataf, n_dataf, datai, n_datai);
  return rc != PM4_OK ? rc : upload_glm(m, dataf);
}

extern "C" int pm4_model_get_mass_scale(pm4_model_t* m, int dtype, void* out_dev, void* stream) {
  if (!m || !out_dev) return fail(PM4_EINVAL, "pm4_model_get_mass_scale: null argument");
  if (!m->mass_scale || m->mass_scale_dtype != dtype) return fail(PM4_EINVAL, "no NUTS run with mass_windows > 0 in this dtype yet");
  CU(cudaSetDevice(m->device));
  CU(cudaMemcpyAsync(out_dev, m->mass_scale, (dtype == PM4_F64 ? 8 : 4) * (size_t)m->D, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return PM4_OK;
}

extern "C" int pm4_model_set_trace_sink(pm4_model_t* m, void* states_host, int n_slabs, void* copy_stream) {
  if (!m) return fail(PM4_EINVAL, "null model");
  if (!states_host || n_slabs <= 0) {  // detach
    m->sink_host = nullptr;
    m->sink_slabs = 0;
    return PM4_OK;
  }
  if (!copy_stream) return fail(PM4_EINVAL, "pm4_model_set_trace_sink: the copies need their own stream");
  CU(cudaSetDevice(m->device));
  if (!m->sink_event) CU(cudaEventCreateWithFlags(&m->sink_event, cudaEventDisableTiming));
  m->sink_host = states_host;
  m->sink_slabs = n_slabs > 64 ? 64 : n_slabs;
  m->sink_stream = (cudaStream_t)copy_stream;
  return PM4_OK;
}

extern "C" int pm4_model_set_plans(pm4_model_t* m, const pm4_ir_t* ir) {
  Nvtx nvtx_range("pm4_model_set_plans");
  if (!m || !ir) return fail(PM4_EINVAL, "null argument");
  if (ir->n_plans != m->n_plans || ir->n_planf != m->n_planf || ir->n_plani != m->n_plani)
    return fail(PM4_EINVAL, "plan pool sizes differ from the model's");
  CU(cudaSetDevice(m->device));
  CU(cudaDeviceSynchronize());
  return upload_plans(m, ir);
}

static int check_dtype(const pm4_model* m, int dtype) {
  if (dtype != PM4_F32 && dtype != PM4_F64) return fail(PM4_EINVAL, "dtype must be PM4_F32 or PM4_F64");
  if (dtype == PM4_F64 && !m->info.has_f64) return fail(PM4_ENOSYS, "this kernel module was built without double precision");
  return PM4_OK;
}

extern "C" int pm4_logp_grad(pm4_model_t* m, int dtype, int C, const void* theta_dev, void* lp_dev, void* grad_dev, void* stream) {
  Nvtx nvtx_range("pm4_logp_grad");
  if (!m || !theta_dev || !lp_dev || C < 0) return fail(PM4_EINVAL, "pm4_logp_grad: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (C == 0) return PM4_OK;
  CU(cudaSetDevice(m->device));
  pm4_mod_args_t a = m->args(dtype);
  MOD(m->f_logp(dtype, &a, C, theta_dev, lp_dev, grad_dev, stream), "logp_grad");
  if (m->lockstep()) {
    if ((rc = add_glm(m, dtype, C, theta_dev, lp_dev, grad_dev, (cudaStream_t)stream))) return rc;
    return check_glm(m, (cudaStream_t)stream);
  }
  return PM4_OK;
}

extern "C" int pm4_deterministics(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev, void* out_dev, void* stream) {
  Nvtx nvtx_range("pm4_deterministics");
  if (!m || M < 0) return fail(PM4_EINVAL, "pm4_deterministics: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (M == 0 || m->det_row == 0) return PM4_OK;
  if (!theta_dev || !out_dev) return fail(PM4_EINVAL, "pm4_deterministics: null buffer");
  CU(cudaSetDevice(m->device));
  pm4_mod_args_t a = m->args(dtype);
  MOD(m->f_dets(dtype, &a, M, theta_dev, out_dev, stream), "deterministics");
  return PM4_OK;
}

extern "C" int pm4_model_ll_row(const pm4_model_t* m) { return m ? m->info.ll_row : PM4_EINVAL; }

extern "C" int pm4_log_likelihood(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev, void* out_dev, void* stream) {
  Nvtx nvtx_range("pm4_log_likelihood");
  if (!m || M < 0) return fail(PM4_EINVAL, "pm4_log_likelihood: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (M == 0 || m->info.ll_row == 0) return PM4_OK;
  if (!theta_dev || !out_dev) return fail(PM4_EINVAL, "pm4_log_likelihood: null buffer");
  CU(cudaSetDevice(m->device));
  pm4_mod_args_t a = m->args(dtype);
  MOD(m->f_loglik(dtype, &a, M, theta_dev, out_dev, stream), "log_likelihood");
  if (!m->gp.empty()) {  // one density value per state, written at the term's position of the row
    if (M > 0x7fffffff) return fail(PM4_EINVAL, "pm4_log_likelihood: too many states for a GP model");
    return add_gp(m, dtype, (int)M, theta_dev, out_dev, m->info.ll_row, false, nullptr, (cudaStream_t)stream);
  }
  return PM4_OK;
}

extern "C" int pm4_posterior_predictive(pm4_model_t* m, int dtype, int64_t M, const void* theta_dev, uint64_t seed,
                                        void* out_dev, void* stream) {
  Nvtx nvtx_range("pm4_posterior_predictive");
  if (!m || M < 0) return fail(PM4_EINVAL, "pm4_posterior_predictive: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (M == 0 || m->info.ll_row == 0) return PM4_OK;
  if (!theta_dev || !out_dev) return fail(PM4_EINVAL, "pm4_posterior_predictive: null buffer");
  CU(cudaSetDevice(m->device));
  pm4_mod_args_t a = m->args(dtype);
  MOD(m->f_predict(dtype, &a, M, theta_dev, out_dev, seed, stream), "posterior_predictive");
  return PM4_OK;
}

extern "C" int pm4_gp_points(const pm4_model_t* m, int which) {
  if (!m || which < 0 || which >= (int)m->gp.size()) return PM4_EINVAL;
  return m->gp[(size_t)which].n;
}

extern "C" int pm4_gp_predictive(pm4_model_t* m, int dtype, int which, int64_t M, const void* theta_dev, uint64_t seed,
                                 void* out_dev, void* stream) {
  Nvtx nvtx_range("pm4_gp_predictive");
  if (!m || M < 0 || which < 0 || which >= (int)m->gp.size()) return fail(PM4_EINVAL, "pm4_gp_predictive: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (M == 0) return PM4_OK;
  if (!theta_dev || !out_dev) return fail(PM4_EINVAL, "pm4_gp_predictive: null buffer");
  if (M > 0x7fffffff) return fail(PM4_EINVAL, "pm4_gp_predictive: too many states");
  CU(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const GpTerm& g = m->gp[(size_t)which];
  if ((rc = gp_workspace(m, g, (int)M, st))) return rc;
  const size_t rs = dtype == PM4_F64 ? 8 : 4, per_chain = (size_t)pm4_gp_work_bytes(g.n, 1);
  const int chunk = (int)std::min<size_t>((size_t)M, m->gp_work_bytes / per_chain);
  for (int c0 = 0; c0 < (int)M; c0 += chunk) {
    const int cc = std::min(chunk, (int)M - c0);
    // element ids of the draws live in their own range: they never coincide with those of another observed term
    rc = pm4_gp_sample(m->d_f64 + g.x_off, m->d_f64 + g.loc_off, g.n, g.d, g.par, g.tr, g.shift, g.scale, dtype,
                       (const char*)theta_dev + (size_t)c0 * m->D * rs, m->D, cc, (int64_t)c0,
                       0x40000000u + ((uint32_t)which << 24), seed, (char*)out_dev + (size_t)c0 * g.n * rs, g.n, m->gp_work, st);
    g_launches += pm4_gp_launches(g.n, 0);
    if (rc != 0) return fail(rc < 0 ? PM4_EINVAL : PM4_ECUDA, "GP kernels: %s", rc < 0 ? "bad argument" : cudaGetErrorString((cudaError_t)rc));
  }
  return PM4_OK;
}

extern "C" int pm4_prior_predictive(pm4_model_t* m, int dtype, int64_t M, uint64_t seed, void* x_dev, void* obs_dev,
                                    void* stream) {
  Nvtx nvtx_range("pm4_prior_predictive");
  if (!m || M < 0) return fail(PM4_EINVAL, "pm4_prior_predictive: bad argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (M == 0) return PM4_OK;
  if (!x_dev) return fail(PM4_EINVAL, "pm4_prior_predictive: null buffer");
  CU(cudaSetDevice(m->device));
  pm4_mod_args_t a = m->args(dtype);
  MOD(m->f_prior(dtype, &a, M, x_dev, obs_dev, seed, stream), "prior_predictive");
  return PM4_OK;
}

// ------------------------------------------------------------------------------------------
// diagonal mass-matrix windows (pm4_adapt_cfg_t.mass_windows): cross-chain standard deviation of the
// unconstrained states -> per-dimension step scale, dual averaging restarted from the current epsilon
// ------------------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) mass_logsd_kernel(const real* __restrict__ th, int C, int DP, double* __restrict__ logsd,
                                                          int* __restrict__ bad) {
  __shared__ double red[256];
  const int d = blockIdx.x;
  double s = 0;
  for (int c = threadIdx.x; c < C; c += 256) s += (double)th[(size_t)c * DP + d];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double mean = red[0] / C;
  __syncthreads();
  double v = 0;
  for (int c = threadIdx.x; c < C; c += 256) { const double e = (double)th[(size_t)c * DP + d] - mean; v += e * e; }
  red[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double var = red[0] / C;
    if (!(var > 0) || isinf(var) || isnan(var)) atomicExch(bad, 1);
    else logsd[d] = 0.5 * log(var);
  }
}

// one block: geometric-mean normalisation, new scale, restart of every dual-averaging state, window start word
template <typename real>
__global__ void __launch_bounds__(256) mass_finish_kernel(const double* __restrict__ logsd, int* __restrict__ bad, int D,
                                                           real* __restrict__ scale, real* __restrict__ da, int n_da,
                                                           int32_t* __restrict__ da_t0, int t0) {
  __shared__ double red[256];
  if (*bad) {  // some dimension has no finite positive spread yet: keep the scale and the recursion
    __syncthreads();
    if (threadIdx.x == 0) *bad = 0;
    return;
  }
  double s = 0;
  for (int d = threadIdx.x; d < D; d += 256) s += logsd[d];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const double mean = red[0] / D;
  for (int d = threadIdx.x; d < D; d += 256) scale[d] = (real)exp(logsd[d] - mean);
  for (int c = threadIdx.x; c < n_da; c += 256) {
    real* p = da + (size_t)c * 4;
    p[1] = 0;
    p[2] = 0;
    p[3] = (real)log(10.0 * (double)p[0]);
  }
  if (threadIdx.x == 0) *da_t0 = t0;
}

static size_t real_size(int dtype) { return dtype == PM4_F64 ? 8 : 4; }

extern "C" int64_t pm4_scratch_bytes(const pm4_model_t* m, int dtype, int C, int max_tree_depth) {
  if (!m) return PM4_EINVAL;
  const size_t rs = real_size(dtype), DP = (size_t)m->info.DP;
  pm4_mod_args_t a = m->args(dtype);
  const int64_t tree = m->f_scratch(dtype, &a, C, 0, max_tree_depth);  // 0 when the tree state fits on chip
  if (tree < 0) return tree;
  return (int64_t)(((size_t)C * DP * 2 + (size_t)C * 6) * rs + (size_t)C * 8 + 16) + tree;
}

namespace {
// device buffers of one run; freed on scope exit (stream-ordered)
// The default stream-ordered pool hands freed memory back to the driver at the next synchronisation (release
// threshold 0): every run would map its scratch again and every host synchronisation after a run would pay for the
// unmapping.  Keep up to 256 MiB of freed scratch in the pool (set once per device).
static void retain_run_scratch() {
  static std::atomic<unsigned> done{0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 32) { cudaGetLastError(); return; }
  if (done.fetch_or(1u << dev) & (1u << dev)) return;
  cudaMemPool_t pool;
  uint64_t keep = 256ull << 20;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) != cudaSuccess ||
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess)
    cudaGetLastError();
}

struct RunBuffers {
  cudaStream_t st;
  std::vector<void*> ptrs;
  explicit RunBuffers(cudaStream_t s) : st(s) { retain_run_scratch(); }
  ~RunBuffers() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
  }
  int get(void** out, size_t bytes) {
    cudaError_t e = cudaMallocAsync(out, bytes ? bytes : 16, st);
    if (e != cudaSuccess) return fail(PM4_ENOMEM, "cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
    ptrs.push_back(*out);
    // run scratch starts zeroed: the padded tails of the state rows (columns D .. DP of th / gr) are read by the team
    // kernels and never written, and with the pool retaining memory a buffer may come back with a previous run's
    // contents (a freshly mapped one reads as zeros)
    e = cudaMemsetAsync(*out, 0, bytes ? bytes : 16, st);
    if (e != cudaSuccess) return fail(PM4_ECUDA, "cudaMemsetAsync(%zu): %s", bytes, cudaGetErrorString(e));
    return PM4_OK;
  }
};

// Rewrites the initial dual-averaging state the module's bootstrap left: per-chain initial step sizes (`eps_c`, or
// NULL) and a shrinkage target other than 10 x the initial step size (`log_target`, or NaN).  Rows: [eps, error_sum,
// log_avg, log_shrink_target].
template <typename real>
__global__ void da_override_kernel(real* __restrict__ da, int rows, const real* __restrict__ eps_c, double log_target) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= rows) return;
  real* d = da + (size_t)c * 4;
  if (eps_c) {
    d[0] = eps_c[c];
    d[3] = (real)log(10.0 * (double)eps_c[c]);
  }
  if (log_target == log_target) d[3] = (real)log_target;
}

static int da_override(int dtype, void* da, int rows, const void* eps_c, double shrinkage_target, cudaStream_t st) {
  if (!eps_c && !(shrinkage_target > 0)) return PM4_OK;
  const double lt = shrinkage_target > 0 ? log(shrinkage_target) : (double)NAN;
  const unsigned grid = (unsigned)((rows + 255) / 256);
  if (dtype == PM4_F64) da_override_kernel<double><<<grid, 256, 0, st>>>((double*)da, rows, (const double*)eps_c, lt);
  else da_override_kernel<float><<<grid, 256, 0, st>>>((float*)da, rows, (const float*)eps_c, lt);
  CU(cudaGetLastError());
  return PM4_OK;
}

int fill_adapt(pm4_mod_run_t& r, const pm4_adapt_cfg_t& a) {
  if (a.mode != PM4_EPS_POOLED && a.mode != PM4_EPS_PER_CHAIN) return fail(PM4_EINVAL, "unknown step-size mode %d", a.mode);
  if (a.kind != PM4_ADAPT_DUAL_AVERAGING && a.kind != PM4_ADAPT_SIMPLE) return fail(PM4_EINVAL, "unknown adaptation kind %d", a.kind);
  r.adapt_kind = a.kind;
  r.adapt_rate = a.adaptation_rate;
  r.adapt_enabled = a.enabled ? 1 : 0;
  r.adapt_pooled = a.mode == PM4_EPS_POOLED ? 1 : 0;
  r.num_adapt = a.num_adaptation_steps;
  r.target_accept = a.target_accept_prob;
  r.shrinkage = a.exploration_shrinkage;
  r.smoothing = a.step_count_smoothing;
  r.decay = a.decay_rate;
  return PM4_OK;
}

// run T transitions: lock-step single-transition launches while a pooled step size still moves
// (transitions 0..num_adapt), then a few persistent launches for the rest.  NUTS chains differ a lot
// in cost (per-chain step sizes): after each phase the chains are ranked by the leapfrogs they took
// and the next launch hands them out longest-first (`order_buf`: int32[C], `r.launch_leaps` set).
struct MassCtx {  // device buffers of the mass-matrix windows of one run (null when off)
  int windows = 0, num_adapt = 0;
  double* logsd = nullptr;
  int* bad = nullptr;
  void* scale = nullptr;
};

// window starts in (lo, hi): w_k = num_adapt * k / (K + 1), k = 1 .. K
static std::vector<int> mass_cuts(const MassCtx& mc, int lo, int hi) {
  std::vector<int> out;
  for (int k = 1; k <= mc.windows; ++k) {
    const int w = (int)(((int64_t)mc.num_adapt * k) / (mc.windows + 1));
    if (w > 0 && w < mc.num_adapt && w > lo && w < hi && (out.empty() || out.back() != w)) out.push_back(w);
  }
  return out;
}

static int mass_window(pm4_model* m, int dtype, const pm4_mod_run_t& r, const MassCtx& mc, int t0, cudaStream_t st) {
  const int n_da = r.adapt_pooled ? 1 : r.C;
  if (dtype == PM4_F64) {
    mass_logsd_kernel<double><<<m->D, 256, 0, st>>>((const double*)r.th, r.C, m->info.DP, mc.logsd, mc.bad);
    mass_finish_kernel<double><<<1, 256, 0, st>>>(mc.logsd, mc.bad, m->D, (double*)mc.scale, (double*)r.da, n_da, (int32_t*)r.da_t0, t0);
  } else {
    mass_logsd_kernel<float><<<m->D, 256, 0, st>>>((const float*)r.th, r.C, m->info.DP, mc.logsd, mc.bad);
    mass_finish_kernel<float><<<1, 256, 0, st>>>(mc.logsd, mc.bad, m->D, (float*)mc.scale, (float*)r.da, n_da, (int32_t*)r.da_t0, t0);
  }
  g_launches += 2;
  CU(cudaGetLastError());
  return PM4_OK;
}

int drive(pm4_model* m, int dtype, pm4_mod_run_t& r, bool nuts, int wpb, int T, int32_t* order_buf, cudaStream_t st,
          const MassCtx& mc = MassCtx()) {
  pm4m_run_fn f = nuts ? m->f_nuts : m->f_hmc;
  int t = 0;
  r.order = nullptr;
  // PM4_PHASE_TIMES=1 (diagnostics): device time of the lock-step phase and of the persistent launches on stderr
  static const bool phase_times = getenv("PM4_PHASE_TIMES") != nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (phase_times) {
    for (auto& e : ev) cudaEventCreate(&e);
    cudaEventRecord(ev[0], st);
  }
  static const bool no_persistent = getenv("PM4_NO_PERSISTENT_LOCKSTEP") != nullptr;
  if (r.adapt_enabled && r.adapt_pooled && nuts && (m->info.caps & PM4_CAP_PERSISTENT_LOCKSTEP) && r.sync && !no_persistent) {
    // the module runs the lock-step transitions in one cooperative launch (grid barrier + pooled update in the kernel)
    const int coupled = r.num_adapt + 1 < T ? r.num_adapt + 1 : T;
    std::vector<int> cuts = mass_cuts(mc, 0, coupled);  // a mass-matrix window starts a new launch
    cuts.push_back(coupled);
    for (int cut : cuts) {
      if (t > 0)
        if (int rc2 = mass_window(m, dtype, r, mc, t, st)) return rc2;
      r.t_begin = t;
      r.t_count = cut - t;
      r.rounds = cut - t;
      r.tb = 0;
      MOD(f(dtype, &r, wpb, st), "nuts (lock-step rounds)");
      t = cut;
    }
    r.rounds = 0;
  } else if (r.adapt_enabled && r.adapt_pooled) {
    const int coupled = r.num_adapt + 1 < T ? r.num_adapt + 1 : T;
    const std::vector<int> cuts = mass_cuts(mc, 0, coupled);
    for (; t < coupled; ++t) {
      if (std::find(cuts.begin(), cuts.end(), t) != cuts.end())
        if (int rc2 = mass_window(m, dtype, r, mc, t, st)) return rc2;
      r.t_begin = t;
      r.t_count = 1;
      MOD(f(dtype, &r, wpb, st), nuts ? "nuts (lock-step)" : "hmc (lock-step)");
      MOD(m->f_da(dtype, &r, t, st), "pooled dual averaging");
    }
  }
  const int t_coupled = t;
  if (phase_times) cudaEventRecord(ev[1], st);
  // phase boundaries: a short first phase (step sizes settle within ~20 transitions), the end of
  // adaptation, the end of the run
  std::vector<int> ends;
  // (a pooled step size is frozen after the lock-step phase: the chains are balanced, one launch is enough)
  const bool phased = nuts && order_buf && r.launch_leaps && !(r.adapt_enabled && r.adapt_pooled);
  if (phased) {
    const int freeze = r.adapt_enabled ? r.num_adapt + 1 : 0;
    if (t + 20 < T - 20) ends.push_back(t + 20);
    if (freeze > (ends.empty() ? t : ends.back()) + 10 && freeze < T - 10) ends.push_back(freeze);
  }
  ends.push_back(T);
  {  // per-chain step sizes: the mass-matrix windows cut the persistent launches
    std::vector<int> cuts = mass_cuts(mc, t, T);
    ends.insert(ends.end(), cuts.begin(), cuts.end());
    std::sort(ends.begin(), ends.end());
    ends.erase(std::unique(ends.begin(), ends.end()), ends.end());
  }
  const std::vector<int> window_starts = mass_cuts(mc, t, T);
  // persistent launches of the warp-per-chain path hand out (chain, block of tb transitions) items
  static const int tb_env = getenv("PM4_TB") ? atoi(getenv("PM4_TB")) : 10;
  r.tb = nuts && r.progress ? tb_env : 0;
  // trace sink: the kept draws of the last segment are produced in `sink_slabs` launches; the states of a finished
  // slab travel to the (pinned) host buffer on the copy stream while the next slab is sampled
  const bool sink = nuts && m->sink_host && m->sink_slabs > 0 && r.states && r.S > 0;
  if (sink) {
    const int last_begin = ends.size() > 1 ? ends[ends.size() - 2] : t;
    const int first = last_begin > r.B ? last_begin : r.B;  // the slabs cover draws only
    ends.pop_back();
    if (first > t && (ends.empty() || ends.back() < first)) ends.push_back(first);
    for (int k = 1; k <= m->sink_slabs; ++k) {
      const int e = first + (int)(((int64_t)(T - first) * k) / m->sink_slabs);
      if (e > (ends.empty() ? t : ends.back())) ends.push_back(e);
    }
    if (ends.empty() || ends.back() != T) ends.push_back(T);
  }
  const size_t row_bytes = (size_t)r.C * (size_t)m->D * (dtype == PM4_F64 ? 8 : 4);  // one draw of all chains
  int copied_to = 0;  // draws already handed to the copy stream
  auto flush_sink = [&](int t_done) -> int {
    const int k1 = t_done - r.B;  // draws [0, k1) are complete
    if (!sink || k1 <= copied_to) return PM4_OK;
    CU(cudaEventRecord(m->sink_event, st));
    CU(cudaStreamWaitEvent(m->sink_stream, m->sink_event, 0));
    CU(cudaMemcpyAsync((char*)m->sink_host + (size_t)copied_to * row_bytes, (const char*)r.states + (size_t)copied_to * row_bytes,
                       (size_t)(k1 - copied_to) * row_bytes, cudaMemcpyDeviceToHost, m->sink_stream));
    copied_to = k1;
    return PM4_OK;
  };
  for (size_t k = 0; k < ends.size(); ++k) {
    if (ends[k] <= t) continue;
    if (std::find(window_starts.begin(), window_starts.end(), t) != window_starts.end())
      if (int rc2 = mass_window(m, dtype, r, mc, t, st)) return rc2;
    r.t_begin = t;
    r.t_count = ends[k] - t;
    MOD(f(dtype, &r, wpb, st), nuts ? "nuts" : "hmc");
    t = ends[k];
    if (int rc2 = flush_sink(t)) return rc2;
    if (phased && t < T && !sink) {
      MOD(m->f_order(r.C, r.launch_leaps, order_buf, st), "longest-first chain order");
      r.order = order_buf;
    }
  }
  if (sink) {  // one-shot
    m->sink_host = nullptr;
    m->sink_slabs = 0;
  }
  if (phase_times) {
    cudaEventRecord(ev[2], st);
    cudaEventSynchronize(ev[2]);
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    fprintf(stderr, "[pm4 phases] lock-step %d transitions %.3f ms (%.1f us each); persistent %d transitions %.3f ms\n",
            t_coupled, a, t_coupled ? 1e3 * a / t_coupled : 0.0, T - t_coupled, b);
    for (auto& e : ev) cudaEventDestroy(e);
  }
  return PM4_OK;
}
}  // namespace

// Lock-step NUTS for models with dense GLM terms (csrc/pm4_lockstep.cuh): one batched gradient per
// lock-step leapfrog for every chain whose tree is still growing -- the reference's execution model
// (tfp NoUTurnSampler under tf.vectorized_map: every chain pays for the deepest tree of the transition).
static int nuts_lockstep(pm4_model* m, int dtype, const pm4_nuts_cfg_t* cfg, const void* theta0_dev,
                         const void* step_scale_dev, const void* momenta_dev, void* states_dev, void* lp_dev,
                         int32_t* leapfrogs_dev, uint8_t* diverging_dev, void* energy_dev, void* log_accept_dev,
                         int32_t* depth_dev, void* step_size_dev, int64_t* total_leapfrogs_dev, cudaStream_t st) {
  if (dtype != PM4_F32) return fail(PM4_ENOSYS, "the lock-step GLM path is float32 only");
  const int C = cfg->C, D = m->D, T = cfg->num_burnin + cfg->num_results;
  const size_t CD = (size_t)C * D;
  int rc;
  pm4_mod_run_t r;  // only what fill_adapt and the pooled dual-averaging kernel read
  memset(&r, 0, sizeof r);
  r.C = C;
  if ((rc = fill_adapt(r, cfg->adapt))) return rc;
  fill_comm(m, r);  // chains sharded over the GPUs of the box: the pooled adaptation kernel exchanges over NVLink
  RunBuffers buf(st);
  pm4ls::NutsParams p;
  memset(&p, 0, sizeof p);
  p.max_depth = cfg->max_tree_depth;
  p.ns = pm4ls::NS_MEM + 2 * (cfg->max_tree_depth + 1);
  p.max_energy_diff = (float)cfg->max_energy_diff;
  void* q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.h.th = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.h.g = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.thb = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.gb = (float*)q;
  if ((rc = buf.get(&q, CD * 4 * (size_t)p.ns))) return rc; p.slots = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * sizeof(pm4ls::NutsChain)))) return rc; p.cs = (pm4ls::NutsChain*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.h.lp = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.lpb = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 16))) return rc; p.h.da = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.h.accept = (float*)q;
  if ((rc = buf.get(&q, 64 * sizeof(int)))) return rc; p.active = (int*)q;
  int* const active_ring = (int*)q;
  p.h.C = C; p.h.D = D; p.h.B = cfg->num_burnin; p.h.S = cfg->num_results;
  p.h.chain_offset = cfg->chain_offset;
  p.h.adapt_enabled = r.adapt_enabled; p.h.adapt_pooled = r.adapt_pooled; p.h.adapt_kind = r.adapt_kind; p.h.num_adapt = r.num_adapt;
  p.h.target_accept = (float)r.target_accept; p.h.shrinkage = (float)r.shrinkage; p.h.smoothing = (float)r.smoothing;
  p.h.decay = (float)r.decay; p.h.adapt_rate = (float)r.adapt_rate;
  p.h.seed_lo = (uint32_t)cfg->seed; p.h.seed_hi = (uint32_t)(cfg->seed >> 32);
  p.h.step_scale = (const float*)step_scale_dev; p.h.momenta = (const float*)momenta_dev;
  p.h.states = (float*)states_dev; p.h.lp_out = (float*)lp_dev; p.h.log_accept = (float*)log_accept_dev;
  p.leapfrogs = leapfrogs_dev; p.diverging = diverging_dev; p.energy = (float*)energy_dev; p.depth_out = depth_dev;
  p.total_leapfrogs = (long long*)total_leapfrogs_dev;
  r.da = p.h.da; r.accept = p.h.accept;
  if (!m->h_flag) CU(cudaMallocHost((void**)&m->h_flag, sizeof(int)));
  pm4_mod_args_t a = m->args(dtype);
  auto full_grad = [&](const float* th, float* lp, float* g, const int* gate) -> int {
    MOD(m->f_logp(dtype, &a, C, th, lp, g, st), "logp_grad (lock-step)");
    return add_glm(m, dtype, C, th, lp, g, st, gate);
  };
  const unsigned warp_grid = (unsigned)((C + 7) / 8);
  CU(cudaMemcpyAsync(p.h.th, theta0_dev, CD * 4, cudaMemcpyDeviceToDevice, st));
  if (total_leapfrogs_dev) CU(cudaMemsetAsync(total_leapfrogs_dev, 0, sizeof(int64_t) * (size_t)C, st));
  if ((rc = full_grad(p.h.th, p.h.lp, p.h.g, nullptr))) return rc;
  pm4ls::init_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(p.h, (float)cfg->step_size);
  if (cfg->step_size_dev && r.adapt_pooled) return fail(PM4_EINVAL, "per-chain initial step sizes need the per-chain adaptation mode");
  if ((rc = da_override(dtype, p.h.da, r.adapt_pooled ? 1 : C, cfg->step_size_dev, cfg->adapt.shrinkage_target, st))) return rc;
  // The leapfrogs of a transition are enqueued in BATCHES without a host round trip in between: every iteration
  // counts the chains that asked for another gradient in its own word of a ring, the dense evaluation of the next
  // iteration is gated on that word (a kernel that finds 0 returns at once), and the host reads one word per batch.
  // Only the fused GLM evaluation can be gated; with GP terms (0.7 s per evaluation) the batch is one iteration.
  bool gated = m->gp.empty() && !m->glm.empty();
  for (const GlmTerm& g : m->glm) gated = gated && g.d_img != nullptr;
  const bool no_batch = getenv("PM4_LOCKSTEP_BATCH") && atoi(getenv("PM4_LOCKSTEP_BATCH")) == 0;  // (read per call: tests toggle it)
  if (no_batch) gated = false;
  constexpr int RING = 64;
  int* const ring = active_ring;
  for (int t = 0; t < T; ++t) {
    p.h.t = t;
    pm4ls::nuts_begin_kernel<<<warp_grid, 256, 0, st>>>(p);
    g_launches += 2;
    int j = 0, batch = gated ? 4 : 1;
    for (;;) {
      CU(cudaMemsetAsync(ring, 0, sizeof(int) * RING, st));  // the words this batch will count into (batch <= RING / 2)
      const int j0 = j;
      for (; j < j0 + batch; ++j) {
        const int* gate = (gated && j > j0) ? ring + (j - 1 - j0) : nullptr;  // the first iteration of a batch is known to be needed
        if ((rc = full_grad(p.thb, p.lpb, p.gb, gate))) return rc;
        p.active = ring + (j - j0);
        pm4ls::nuts_advance_kernel<<<warp_grid, 256, 0, st>>>(p);
        g_launches += 1;
      }
      CU(cudaMemcpyAsync(m->h_flag, ring + (batch - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
      CU(cudaStreamSynchronize(st));
      if (*m->h_flag == 0) break;
      if (gated && batch < RING / 2) batch *= 2;
    }
    pm4ls::nuts_finish_kernel<<<warp_grid, 256, 0, st>>>(p);
    if (p.h.adapt_enabled && p.h.adapt_pooled) MOD(m->f_da(dtype, &r, t, st), "pooled dual averaging");
  }
  CU(cudaGetLastError());
  if (step_size_dev) MOD(m->f_eps(dtype, &r, step_size_dev, st), "export step size");
  if ((rc = check_glm(m, st))) return rc;
  return (r.adapt_enabled && r.adapt_pooled) ? check_comm(m, st) : PM4_OK;
}

extern "C" int pm4_nuts_run(pm4_model_t* m, int dtype, const pm4_nuts_cfg_t* cfg, const void* theta0_dev,
                            const void* step_scale_dev, const void* momenta_dev, void* states_dev, void* lp_dev,
                            int32_t* leapfrogs_dev, uint8_t* diverging_dev, void* energy_dev, void* log_accept_dev,
                            int32_t* depth_dev, void* step_size_dev, int64_t* total_leapfrogs_dev, void* stream) {
  Nvtx nvtx_range("pm4_nuts_run");
  if (!m || !cfg || !theta0_dev) return fail(PM4_EINVAL, "pm4_nuts_run: null argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (cfg->C <= 0 || cfg->num_results < 0 || cfg->num_burnin < 0) return fail(PM4_EINVAL, "pm4_nuts_run: bad chain/draw counts");
  if (cfg->max_tree_depth < 1 || cfg->max_tree_depth > 20) return fail(PM4_EINVAL, "max_tree_depth must be in [1, 20]");
  if (!(cfg->step_size > 0)) return fail(PM4_EINVAL, "step_size must be positive");
  CU(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (m->lockstep())
    return nuts_lockstep(m, dtype, cfg, theta0_dev, step_scale_dev, momenta_dev, states_dev, lp_dev, leapfrogs_dev,
                         diverging_dev, energy_dev, log_accept_dev, depth_dev, step_size_dev, total_leapfrogs_dev, st);
  const size_t rs = real_size(dtype), C = (size_t)cfg->C, DP = (size_t)m->info.DP;
  const int T = cfg->num_burnin + cfg->num_results;

  pm4_mod_run_t r;
  memset(&r, 0, sizeof r);
  r.A = m->args(dtype);
  r.C = cfg->C; r.B = cfg->num_burnin; r.S = cfg->num_results;
  r.max_depth = cfg->max_tree_depth;
  r.chain_offset = cfg->chain_offset;
  r.max_energy_diff = cfg->max_energy_diff;
  r.seed_lo = (uint32_t)cfg->seed; r.seed_hi = (uint32_t)(cfg->seed >> 32);
  if ((rc = fill_adapt(r, cfg->adapt))) return rc;
  RunBuffers buf(st);
  if ((rc = buf.get(&r.th, C * DP * rs))) return rc;
  if ((rc = buf.get(&r.gr, C * DP * rs))) return rc;
  if ((rc = buf.get(&r.lp, C * rs))) return rc;
  if ((rc = buf.get(&r.da, C * 4 * rs))) return rc;
  if ((rc = buf.get(&r.accept, C * rs))) return rc;
  const int64_t tree_bytes = m->f_scratch(dtype, &r.A, cfg->C, cfg->warps_per_block, cfg->max_tree_depth);
  if (tree_bytes < 0) return fail(PM4_EINVAL, "kernel module cannot size its tree scratch");
  if (tree_bytes > 0 && (rc = buf.get(&r.scratch, (size_t)tree_bytes))) return rc;
  void *queue = nullptr, *order = nullptr, *leaps = nullptr, *progress = nullptr, *sync = nullptr;
  if ((rc = buf.get(&queue, sizeof(int32_t) * (size_t)(T + 4)))) return rc;  // one counter per lock-step round
  if ((rc = buf.get(&sync, 16))) return rc;
  r.sync = (int32_t*)sync;
  if ((rc = buf.get(&order, C * sizeof(int32_t)))) return rc;
  if ((rc = buf.get(&leaps, C * sizeof(int32_t)))) return rc;
  if ((rc = buf.get(&progress, C * sizeof(int32_t)))) return rc;
  r.queue = (int32_t*)queue;
  r.launch_leaps = (int32_t*)leaps;
  r.progress = (int32_t*)progress;
  r.step_scale = step_scale_dev; r.momenta = momenta_dev;
  r.states = states_dev; r.lp_out = lp_dev; r.leapfrogs = leapfrogs_dev; r.diverging = diverging_dev;
  r.energy = energy_dev; r.log_accept = log_accept_dev; r.depth = depth_dev;
  r.total_leapfrogs = total_leapfrogs_dev;

  fill_comm(m, r);
  void* t0w = nullptr;  // first transition of the current dual-averaging window (mass-matrix windows move it)
  if ((rc = buf.get(&t0w, 16))) return rc;
  CU(cudaMemsetAsync(t0w, 0, 16, st));
  r.da_t0 = (const int32_t*)t0w;
  MassCtx mc;
  if (r.adapt_enabled && cfg->adapt.mass_windows > 0) {
    mc.windows = cfg->adapt.mass_windows;
    mc.num_adapt = r.num_adapt;
    void *ls = nullptr, *bad = nullptr;
    if ((rc = buf.get(&ls, sizeof(double) * (size_t)m->D))) return rc;
    if ((rc = buf.get(&bad, 16))) return rc;
    CU(cudaMemsetAsync(bad, 0, 16, st));
    mc.logsd = (double*)ls;
    mc.bad = (int*)bad;
    // the scale outlives the run (pm4_model_get_mass_scale): owned by the model
    if (m->mass_scale && m->mass_scale_dtype != dtype) { cudaFree(m->mass_scale); m->mass_scale = nullptr; }
    if (!m->mass_scale) CU(cudaMalloc(&m->mass_scale, rs * (size_t)m->D));
    m->mass_scale_dtype = dtype;
    mc.scale = m->mass_scale;
    if (step_scale_dev) CU(cudaMemcpyAsync(mc.scale, step_scale_dev, rs * (size_t)m->D, cudaMemcpyDeviceToDevice, st));
    else {
      std::vector<double> ones64((size_t)m->D, 1.0);
      std::vector<float> ones32((size_t)m->D, 1.0f);
      CU(cudaMemcpyAsync(mc.scale, dtype == PM4_F64 ? (const void*)ones64.data() : (const void*)ones32.data(), rs * (size_t)m->D,
                         cudaMemcpyHostToDevice, st));
      CU(cudaStreamSynchronize(st));  // the host vectors go out of scope
    }
    r.step_scale = mc.scale;
  }
  MOD(m->f_init(dtype, &r, theta0_dev, cfg->step_size, st), "bootstrap");
  if (cfg->step_size_dev && r.adapt_pooled) return fail(PM4_EINVAL, "per-chain initial step sizes need the per-chain adaptation mode");
  if ((rc = da_override(dtype, r.da, r.adapt_pooled ? 1 : r.C, cfg->step_size_dev, cfg->adapt.shrinkage_target, st))) return rc;
  if ((rc = drive(m, dtype, r, true, cfg->warps_per_block, T, (int32_t*)order, st, mc))) return rc;
  if (step_size_dev) MOD(m->f_eps(dtype, &r, step_size_dev, st), "export step size");
  if (r.adapt_enabled && r.adapt_pooled) return check_comm(m, st);
  return PM4_OK;
}

// Lock-step HMC for models with dense GLM terms (csrc/pm4_lockstep.cuh): every leapfrog is one batched
// gradient -- the generated module's per-chain pass over the other terms + the tensor-core GLM pass.
static int hmc_lockstep(pm4_model* m, int dtype, const pm4_hmc_cfg_t* cfg, const void* theta0_dev,
                        const void* step_scale_dev, const void* momenta_dev, const void* uniforms_dev, void* states_dev,
                        void* lp_dev, void* log_accept_dev, uint8_t* accepted_dev, void* step_size_dev, void* traj_dev,
                        cudaStream_t st) {
  if (dtype != PM4_F32) return fail(PM4_ENOSYS, "the lock-step GLM path is float32 only");
  if (cfg->proposal != PM4_PROPOSAL_LEAPFROG) return fail(PM4_ENOSYS, "random-walk proposals are not built for GLM models");
  const int C = cfg->C, D = m->D, T = cfg->num_burnin + cfg->num_results;
  const size_t CD = (size_t)C * D;
  int rc;
  pm4_mod_run_t r;  // only what fill_adapt and the pooled dual-averaging kernel read
  memset(&r, 0, sizeof r);
  r.C = C;
  if ((rc = fill_adapt(r, cfg->adapt))) return rc;
  fill_comm(m, r);  // chains sharded over the GPUs of the box: the pooled adaptation kernel exchanges over NVLink
  RunBuffers buf(st);
  pm4ls::Params p;
  memset(&p, 0, sizeof p);
  void* q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.th = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.g = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.thp = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.gp = (float*)q;
  if ((rc = buf.get(&q, CD * 4))) return rc; p.p = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.lp = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.lpp = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.k0 = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 16))) return rc; p.da = (float*)q;
  if ((rc = buf.get(&q, (size_t)C * 4))) return rc; p.accept = (float*)q;
  p.C = C; p.D = D; p.B = cfg->num_burnin; p.S = cfg->num_results; p.L = cfg->num_leapfrog_steps;
  p.chain_offset = cfg->chain_offset;
  p.adapt_enabled = r.adapt_enabled; p.adapt_pooled = r.adapt_pooled; p.adapt_kind = r.adapt_kind; p.num_adapt = r.num_adapt;
  p.target_accept = (float)r.target_accept; p.shrinkage = (float)r.shrinkage; p.smoothing = (float)r.smoothing;
  p.decay = (float)r.decay; p.adapt_rate = (float)r.adapt_rate;
  p.seed_lo = (uint32_t)cfg->seed; p.seed_hi = (uint32_t)(cfg->seed >> 32);
  p.step_scale = (const float*)step_scale_dev; p.momenta = (const float*)momenta_dev; p.uniforms = (const float*)uniforms_dev;
  p.states = (float*)states_dev; p.lp_out = (float*)lp_dev; p.log_accept = (float*)log_accept_dev;
  p.accepted = accepted_dev; p.traj = (float*)traj_dev;
  r.da = p.da; r.accept = p.accept;
  pm4_mod_args_t a = m->args(dtype);
  auto full_grad = [&](const float* th, float* lp, float* g) -> int {
    MOD(m->f_logp(dtype, &a, C, th, lp, g, st), "logp_grad (lock-step)");
    return add_glm(m, dtype, C, th, lp, g, st);
  };
  const unsigned warp_grid = (unsigned)((C + 7) / 8), elem_grid = (unsigned)((CD + 255) / 256);
  CU(cudaMemcpyAsync(p.th, theta0_dev, CD * 4, cudaMemcpyDeviceToDevice, st));
  if ((rc = full_grad(p.th, p.lp, p.g))) return rc;
  pm4ls::init_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(p, (float)cfg->step_size);
  if (cfg->step_size_dev && r.adapt_pooled) return fail(PM4_EINVAL, "per-chain initial step sizes need the per-chain adaptation mode");
  if ((rc = da_override(dtype, p.da, r.adapt_pooled ? 1 : C, cfg->step_size_dev, cfg->adapt.shrinkage_target, st))) return rc;
  for (int t = 0; t < T; ++t) {
    p.t = t;
    pm4ls::begin_kernel<<<warp_grid, 256, 0, st>>>(p);
    for (int l = 0; l < p.L; ++l) {
      pm4ls::drift_kernel<<<elem_grid, 256, 0, st>>>(p);
      if ((rc = full_grad(p.thp, p.lpp, p.gp))) return rc;
      pm4ls::kick_kernel<<<elem_grid, 256, 0, st>>>(p);
    }
    pm4ls::finish_kernel<<<warp_grid, 256, 0, st>>>(p);
    g_launches += 2 + 2 * p.L;
    if (p.adapt_enabled && p.adapt_pooled) MOD(m->f_da(dtype, &r, t, st), "pooled dual averaging");
  }
  CU(cudaGetLastError());
  if (step_size_dev) {
    r.adapt_pooled = p.adapt_pooled;
    MOD(m->f_eps(dtype, &r, step_size_dev, st), "export step size");
  }
  if ((rc = check_glm(m, st))) return rc;
  return (r.adapt_enabled && r.adapt_pooled) ? check_comm(m, st) : PM4_OK;
}

extern "C" int pm4_hmc_run(pm4_model_t* m, int dtype, const pm4_hmc_cfg_t* cfg, const void* theta0_dev,
                           const void* step_scale_dev, const void* momenta_dev, const void* uniforms_dev,
                           void* states_dev, void* lp_dev, void* log_accept_dev, uint8_t* accepted_dev,
                           void* step_size_dev, void* traj_dev, void* stream) {
  Nvtx nvtx_range("pm4_hmc_run");
  if (!m || !cfg || !theta0_dev) return fail(PM4_EINVAL, "pm4_hmc_run: null argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  const bool rw = cfg->proposal == PM4_PROPOSAL_RANDOM_WALK;
  if (cfg->proposal != PM4_PROPOSAL_LEAPFROG && !rw) return fail(PM4_EINVAL, "unknown proposal %d", cfg->proposal);
  if (cfg->C <= 0 || cfg->num_results < 0 || cfg->num_burnin < 0 || (!rw && cfg->num_leapfrog_steps < 1))
    return fail(PM4_EINVAL, "pm4_hmc_run: bad chain/draw/leapfrog counts");
  if (!(cfg->step_size > 0)) return fail(PM4_EINVAL, "step_size must be positive");
  if (rw && !(cfg->rw_scale > 0)) return fail(PM4_EINVAL, "rw_scale must be positive");
  CU(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (m->lockstep())
    return hmc_lockstep(m, dtype, cfg, theta0_dev, step_scale_dev, momenta_dev, uniforms_dev, states_dev, lp_dev,
                        log_accept_dev, accepted_dev, step_size_dev, traj_dev, st);
  const size_t rs = real_size(dtype), C = (size_t)cfg->C, DP = (size_t)m->info.DP;
  const int T = cfg->num_burnin + cfg->num_results;

  pm4_mod_run_t r;
  memset(&r, 0, sizeof r);
  r.A = m->args(dtype);
  r.C = cfg->C; r.B = cfg->num_burnin; r.S = cfg->num_results;
  r.num_leapfrog = cfg->num_leapfrog_steps;
  r.random_walk = rw ? 1 : 0;
  r.rw_scale = cfg->rw_scale;
  r.chain_offset = cfg->chain_offset;
  r.seed_lo = (uint32_t)cfg->seed; r.seed_hi = (uint32_t)(cfg->seed >> 32);
  if ((rc = fill_adapt(r, cfg->adapt))) return rc;
  RunBuffers buf(st);
  if ((rc = buf.get(&r.th, C * DP * rs))) return rc;
  if ((rc = buf.get(&r.gr, C * DP * rs))) return rc;
  if ((rc = buf.get(&r.lp, C * rs))) return rc;
  if ((rc = buf.get(&r.da, C * 4 * rs))) return rc;
  if ((rc = buf.get(&r.accept, C * rs))) return rc;
  void* queue = nullptr;
  if ((rc = buf.get(&queue, 16))) return rc;
  r.queue = (int32_t*)queue;
  r.step_scale = step_scale_dev; r.momenta = momenta_dev; r.uniforms = uniforms_dev;
  r.states = states_dev; r.lp_out = lp_dev; r.log_accept = log_accept_dev; r.diverging = accepted_dev;
  r.traj = traj_dev;

  fill_comm(m, r);
  MOD(m->f_init(dtype, &r, theta0_dev, cfg->step_size, st), "bootstrap");
  if (cfg->step_size_dev && r.adapt_pooled) return fail(PM4_EINVAL, "per-chain initial step sizes need the per-chain adaptation mode");
  if ((rc = da_override(dtype, r.da, r.adapt_pooled ? 1 : r.C, cfg->step_size_dev, cfg->adapt.shrinkage_target, st))) return rc;
  if ((rc = drive(m, dtype, r, false, cfg->warps_per_block, T, nullptr, st))) return rc;
  if (step_size_dev) MOD(m->f_eps(dtype, &r, step_size_dev, st), "export step size");
  if (r.adapt_enabled && r.adapt_pooled) return check_comm(m, st);
  return PM4_OK;
}

// Compound step (reference pymc4/mcmc/samplers.py:490-726, tf_support.py:52-163): per transition, one transition of
// every block's kernel in order on the SAME chain state (th / gr / lp rows of the run): a block's kernel sees the other
// blocks' dimensions frozen at their current values (step scale 0) -- the reference's
// _target_log_prob_fn_part_compound.  Every block has its own adaptation state and its own Philox streams.
extern "C" int pm4_compound_run(pm4_model_t* m, int dtype, int32_t C_, int32_t num_results, int32_t num_burnin, uint64_t seed,
                                int32_t chain_offset, int32_t warps_per_block, const pm4_compound_block_t* blocks,
                                int32_t n_blocks, const void* theta0_dev, void* states_dev, void* lp_dev,
                                void* step_size_dev, void* log_accept_dev, void* stream) {
  Nvtx nvtx_range("pm4_compound_run");
  if (!m || !blocks || !theta0_dev || n_blocks < 1) return fail(PM4_EINVAL, "pm4_compound_run: null argument");
  int rc = check_dtype(m, dtype);
  if (rc) return rc;
  if (C_ <= 0 || num_results < 0 || num_burnin < 0) return fail(PM4_EINVAL, "pm4_compound_run: bad chain/draw counts");
  if (m->lockstep()) return fail(PM4_ENOSYS, "the compound step is not built for models with dense (GLM / GP) terms");
  for (int b = 0; b < n_blocks; ++b) {
    const pm4_compound_block_t& k = blocks[b];
    if (k.kind != PM4_BLOCK_NUTS && k.kind != PM4_BLOCK_HMC && k.kind != PM4_BLOCK_RWM) return fail(PM4_EINVAL, "block %d: unknown kind %d", b, k.kind);
    if (!k.dim_scale_dev) return fail(PM4_EINVAL, "block %d: dim_scale_dev is required", b);
    if (k.kind == PM4_BLOCK_NUTS && (k.max_tree_depth < 1 || k.max_tree_depth > 20)) return fail(PM4_EINVAL, "block %d: max_tree_depth must be in [1, 20]", b);
    if (k.kind == PM4_BLOCK_HMC && k.num_leapfrog_steps < 1) return fail(PM4_EINVAL, "block %d: num_leapfrog_steps must be positive", b);
    if (k.kind != PM4_BLOCK_RWM && !(k.step_size > 0)) return fail(PM4_EINVAL, "block %d: step_size must be positive", b);
    if (k.kind == PM4_BLOCK_RWM && !(k.rw_scale > 0)) return fail(PM4_EINVAL, "block %d: rw_scale must be positive", b);
    if (k.kind != PM4_BLOCK_RWM && k.adapt.mass_windows > 0) return fail(PM4_ENOSYS, "block %d: mass-matrix windows are not built for the compound step", b);
  }
  CU(cudaSetDevice(m->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rs = real_size(dtype), C = (size_t)C_, DP = (size_t)m->info.DP;
  const int T = num_burnin + num_results, S = num_results;

  pm4_mod_run_t base;
  memset(&base, 0, sizeof base);
  base.A = m->args(dtype);
  base.C = C_; base.B = num_burnin; base.S = num_results;
  base.chain_offset = chain_offset;
  RunBuffers buf(st);
  if ((rc = buf.get(&base.th, C * DP * rs))) return rc;
  if ((rc = buf.get(&base.gr, C * DP * rs))) return rc;
  if ((rc = buf.get(&base.lp, C * rs))) return rc;
  if ((rc = buf.get(&base.accept, C * rs))) return rc;
  int max_depth = 1;
  for (int b = 0; b < n_blocks; ++b)
    if (blocks[b].kind == PM4_BLOCK_NUTS && blocks[b].max_tree_depth > max_depth) max_depth = blocks[b].max_tree_depth;
  const int64_t tree_bytes = m->f_scratch(dtype, &base.A, C_, warps_per_block, max_depth);
  if (tree_bytes < 0) return fail(PM4_EINVAL, "kernel module cannot size its tree scratch");
  if (tree_bytes > 0 && (rc = buf.get(&base.scratch, (size_t)tree_bytes))) return rc;
  void *queue = nullptr, *sync = nullptr, *leaps = nullptr, *progress = nullptr, *t0w = nullptr;
  if ((rc = buf.get(&queue, 64))) return rc;
  if ((rc = buf.get(&sync, 16))) return rc;
  if ((rc = buf.get(&leaps, C * sizeof(int32_t)))) return rc;
  if ((rc = buf.get(&progress, C * sizeof(int32_t)))) return rc;
  if ((rc = buf.get(&t0w, 16))) return rc;
  CU(cudaMemsetAsync(t0w, 0, 16, st));
  base.queue = (int32_t*)queue; base.sync = (int32_t*)sync; base.launch_leaps = (int32_t*)leaps;
  base.progress = (int32_t*)progress; base.da_t0 = (const int32_t*)t0w;
  base.comm_world = 1;  // the blocks' step sizes are not coupled across GPUs: every rank adapts on its own chains

  // per-block run records: adaptation state, streams, outputs
  std::vector<pm4_mod_run_t> runs((size_t)n_blocks, base);
  std::vector<unsigned char> host_da;
  for (int b = 0; b < n_blocks; ++b) {
    const pm4_compound_block_t& k = blocks[b];
    pm4_mod_run_t& r = runs[(size_t)b];
    if (k.kind == PM4_BLOCK_RWM) {
      pm4_adapt_cfg_t off;
      memset(&off, 0, sizeof off);
      off.mode = PM4_EPS_PER_CHAIN;
      if ((rc = fill_adapt(r, off))) return rc;
      r.random_walk = 1;
      r.rw_scale = k.rw_scale;
      r.rw_kind = k.rw_kind_dev;
    } else {
      if ((rc = fill_adapt(r, k.adapt))) return rc;
      r.max_depth = k.max_tree_depth;
      r.num_leapfrog = k.num_leapfrog_steps;
      r.max_energy_diff = k.max_energy_diff;
    }
    r.step_scale = k.dim_scale_dev;
    // independent streams per block: two Metropolis tests of one transition must not share their uniform
    const uint64_t sb = seed + 0x9E3779B97F4A7C15ull * (uint64_t)b;
    r.seed_lo = (uint32_t)sb; r.seed_hi = (uint32_t)(sb >> 32);
    if ((rc = buf.get(&r.da, C * 4 * rs))) return rc;
    // dual-averaging state [eps, error_sum, log_avg, log(10 eps)] of every chain (the pooled mode reads row 0)
    const double eps0 = k.kind == PM4_BLOCK_RWM ? 1.0 : k.step_size;
    host_da.resize(C * 4 * rs);
    for (size_t c = 0; c < C; ++c) {
      const double target = (k.kind != PM4_BLOCK_RWM && k.adapt.shrinkage_target > 0) ? k.adapt.shrinkage_target : 10.0 * eps0;
      const double row[4] = {eps0, 0.0, 0.0, log(target)};
      for (int q = 0; q < 4; ++q) {
        if (dtype == PM4_F64) ((double*)host_da.data())[c * 4 + q] = row[q];
        else ((float*)host_da.data())[c * 4 + q] = (float)row[q];
      }
    }
    CU(cudaMemcpyAsync(r.da, host_da.data(), host_da.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));  // host_da is reused
    // the last block of a transition writes the trace
    const bool last = b == n_blocks - 1;
    r.states = last ? states_dev : nullptr;
    r.lp_out = last ? lp_dev : nullptr;
    r.log_accept = log_accept_dev ? (void*)((char*)log_accept_dev + (size_t)b * (size_t)S * C * rs) : nullptr;
  }
  {  // bootstrap: log-density and gradient of the initial point (block 0's record; its da row is rewritten identically)
    pm4_mod_run_t& r0 = runs[0];
    MOD(m->f_init(dtype, &r0, theta0_dev, blocks[0].kind == PM4_BLOCK_RWM ? 1.0 : blocks[0].step_size, st), "bootstrap");
  }
  const bool persistent_cap = (m->info.caps & PM4_CAP_PERSISTENT_LOCKSTEP) != 0;
  for (int t = 0; t < T; ++t) {
    for (int b = 0; b < n_blocks; ++b) {
      pm4_mod_run_t& r = runs[(size_t)b];
      const bool nuts = blocks[b].kind == PM4_BLOCK_NUTS;
      r.t_begin = t; r.t_count = 1; r.tb = 0; r.order = nullptr;
      const bool coupled = r.adapt_enabled && r.adapt_pooled && t <= r.num_adapt;
      if (coupled && nuts && persistent_cap) {
        r.rounds = 1;
        MOD(m->f_nuts(dtype, &r, warps_per_block, st), "compound: nuts (lock-step round)");
        r.rounds = 0;
      } else {
        r.rounds = 0;
        MOD((nuts ? m->f_nuts : m->f_hmc)(dtype, &r, warps_per_block, st), nuts ? "compound: nuts" : "compound: hmc / rwm");
        if (coupled) MOD(m->f_da(dtype, &r, t, st), "compound: pooled dual averaging");
      }
    }
  }
  if (step_size_dev)
    for (int b = 0; b < n_blocks; ++b)
      MOD(m->f_eps(dtype, &runs[(size_t)b], (char*)step_size_dev + (size_t)b * C * rs, st), "export step size");
  return PM4_OK;
}

extern "C" int pm4_comm_create(int device, int rank, int world, pm4_comm_t** out, unsigned char* handle_out) {
  if (!out || !handle_out || world < 1 || world > PM4_MAX_RANKS || rank < 0 || rank >= world)
    return fail(PM4_EINVAL, "pm4_comm_create: bad argument (at most %d ranks)", PM4_MAX_RANKS);
  static_assert(sizeof(cudaIpcMemHandle_t) == PM4_IPC_HANDLE_BYTES, "IPC handle size");
  CU(cudaSetDevice(device));
  pm4_comm* c = new pm4_comm();
  c->device = device; c->rank = rank; c->world = world;
  const size_t bytes = 32 * PM4_COMM_SLOTS + 64;
  cudaError_t e = cudaMalloc(&c->box[rank], bytes);
  if (e == cudaSuccess) e = cudaMemset(c->box[rank], 0, bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_error, sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(c->d_error, 0, sizeof(int32_t));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->box[rank]);
  if (e != cudaSuccess) {
    int rc = fail(PM4_ECUDA, "pm4_comm_create: %s", cudaGetErrorString(e));
    pm4_comm_destroy(c);
    return rc;
  }
  memcpy(handle_out, &h, sizeof h);
  *out = c;
  return PM4_OK;
}

extern "C" int pm4_comm_connect(pm4_comm_t* c, const unsigned char* handles) {
  if (!c || !handles) return fail(PM4_EINVAL, "pm4_comm_connect: null argument");
  CU(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || c->box[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * PM4_IPC_HANDLE_BYTES, sizeof h);
    cudaError_t e = cudaIpcOpenMemHandle(&c->box[r], h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(PM4_ECUDA, "pm4_comm_connect: cannot map the mailbox of rank %d: %s", r, cudaGetErrorString(e));
  }
  c->connected = true;
  return PM4_OK;
}

extern "C" int pm4_comm_destroy(pm4_comm_t* c) {
  if (!c) return PM4_OK;
  cudaSetDevice(c->device);
  for (int r = 0; r < c->world; ++r) {
    if (!c->box[r]) continue;
    if (r == c->rank) cudaFree(c->box[r]);
    else cudaIpcCloseMemHandle(c->box[r]);
  }
  cudaFree(c->d_error);
  delete c;
  return PM4_OK;
}

extern "C" int pm4_model_attach_comm(pm4_model_t* m, pm4_comm_t* c) {
  if (!m) return fail(PM4_EINVAL, "null model");
  if (c && (!c->connected || c->device != m->device)) return fail(PM4_EINVAL, "communicator is not connected or lives on another device");
  m->comm = c;
  return PM4_OK;
}

// fill the coupling fields of a run; a new epoch per sampling call
static void fill_comm(pm4_model* m, pm4_mod_run_t& r) {
  for (int k = 0; k < PM4_MAX_RANKS; ++k) r.comm_box[k] = nullptr;
  r.comm_rank = 0; r.comm_world = 1; r.comm_epoch = 0; r.comm_error = nullptr;
  if (!m->comm || m->comm->world < 2) return;
  pm4_comm* c = m->comm;
  for (int k = 0; k < c->world; ++k) r.comm_box[k] = c->box[k];
  r.comm_rank = c->rank; r.comm_world = c->world; r.comm_epoch = ++c->epoch; r.comm_error = c->d_error;
}

// after a run with a communicator: did a peer time out?  (synchronises the stream)
static int check_comm(pm4_model* m, cudaStream_t st) {
  if (!m->comm || m->comm->world < 2) return PM4_OK;
  int32_t err = 0;
  CU(cudaMemcpyAsync(&err, m->comm->d_error, sizeof err, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (err) {
    cudaMemsetAsync(m->comm->d_error, 0, sizeof(int32_t), st);
    return fail(PM4_ECUDA, "pooled step-size exchange timed out waiting for rank %d", err - 1);
  }
  return PM4_OK;
}

extern "C" int64_t pm4_launch_count(int reset) {
  const int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

extern "C" const char* pm4_last_error(void) { return g_err.c_str(); }
extern "C" int pm4_abi_version(void) { return PM4_ABI_VERSION; }
