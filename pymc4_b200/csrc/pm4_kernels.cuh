// pm4_kernels.cuh -- the sampler kernels, templated on a generated `Model` and on the arithmetic
// type.  One WARP owns one chain: every D-vector of the chain (position, momentum, gradient,
// trajectory momentum sums) is spread over the 32 lanes, lane l holding elements l, l+32, ...
// in registers (M::R per vector).  Reductions over dimensions / observations are xor-butterfly
// warp shuffles, so every lane sees bit-identical sums and all sampler decisions are
// warp-uniform.  Observed-data records and execution plans of gather terms are copied once per
// CTA into shared memory with a TMA bulk copy (cp.async.bulk + mbarrier).
//
// What this replaces in the reference: the body of tfp.mcmc.sample_chain that
// _BaseSampler._run_chains hands control to (/root/reference/pymc4/mcmc/samplers.py:172-189):
// NoUTurnSampler.one_step / HamiltonianMonteCarlo.one_step wrapped in
// DualAveragingStepSizeAdaptation, calling parallel_logpfn + its gradient per leapfrog
// (samplers.py:117-119).  Algorithm restated in SURVEY App. A; the CPU restatement is
// oracle/pm4_oracle_impl.h.
#pragma once
#include "pm4_common.cuh"

namespace pm4 {

// ------------------------------------------------------------------------------------------
// shared-memory layout of one CTA:
//   [0, 16)                  mbarrier of the bulk copy
//   [16, 16 + pool bytes)    plan records (reals) then plan metadata (ints), when staged
//   then per warp            xs[DP] | gx[DP] | part[part_reals]
// ------------------------------------------------------------------------------------------
template <typename real> __host__ __device__ inline size_t pool_bytes(const ModelArgs<real>& A) {
  return A.stage ? 16 + (size_t)A.n_planf * sizeof(real) + (size_t)A.n_plani * sizeof(int32_t) : 0;
}
template <typename real> __host__ __device__ inline size_t warp_reals(const ModelArgs<real>& A, int DP) {
  return 2 * (size_t)DP + (((size_t)A.part_reals + 3) & ~(size_t)3);
}
template <typename real> __host__ inline size_t cta_smem_bytes(const ModelArgs<real>& A, int DP, int wpb) {
  return pool_bytes(A) + (size_t)wpb * warp_reals(A, DP) * sizeof(real);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Build the per-warp context; when STAGED, thread 0 issues the TMA bulk copies of the plan pools
// and every thread of the CTA waits on the mbarrier.  Must be called by ALL threads of the CTA.
template <typename real, bool STAGED>
__device__ __forceinline__ Ctx<real> make_ctx(const ModelArgs<real>& A, unsigned char* smem_raw, int DP) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Ctx<real> c;
  c.lane = lane;
  c.A = A;
  size_t base = 0;
  if (STAGED) {
    const uint32_t fbytes = (uint32_t)A.n_planf * (uint32_t)sizeof(real), ibytes = (uint32_t)A.n_plani * 4u;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    unsigned char* sf = smem_raw + 16;
    unsigned char* si = sf + fbytes;
    const uint32_t bar = smem_u32(mbar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(fbytes + ibytes) : "memory");
      const unsigned char* gf = reinterpret_cast<const unsigned char*>(A.planf);
      const unsigned char* gi = reinterpret_cast<const unsigned char*>(A.plani);
      for (uint32_t off = 0; off < fbytes; off += 32768u) {
        const uint32_t nb = fbytes - off < 32768u ? fbytes - off : 32768u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sf + off)), "l"(gf + off), "r"(nb), "r"(bar) : "memory");
      }
      for (uint32_t off = 0; off < ibytes; off += 32768u) {
        const uint32_t nb = ibytes - off < 32768u ? ibytes - off : 32768u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(si + off)), "l"(gi + off), "r"(nb), "r"(bar) : "memory");
      }
    }
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(0u) : "memory");
    }
    c.pf = reinterpret_cast<const real*>(sf);
    c.pi = reinterpret_cast<const int32_t*>(si);
    base = 16 + (size_t)fbytes + ibytes;
  } else {
    c.pf = A.planf;
    c.pi = A.plani;
  }
  real* w = reinterpret_cast<real*>(smem_raw + base) + (size_t)warp * warp_reals(A, DP);
  c.xs = w;
  c.gx = w + DP;
  c.part = w + 2 * DP;
  return c;
}

// ------------------------------------------------------------------------------------------
// joint log-density + gradient in unconstrained space for the chain owned by this warp.
// `th`/`g` are the lane-distributed theta and gradient; returns the (warp-uniform) log-density.
// ------------------------------------------------------------------------------------------
template <class M, typename real>
__device__ __forceinline__ real logp_grad(const real (&th)[M::R], real (&g)[M::R], Ctx<real>& c) {
  constexpr int R = M::R, D = M::D;
  real x[R], dxdz[R], gr[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = c.lane + 32 * r;
    M::template constrain<real>(r, c.lane, th[r], x[r], dxdz[r]);
    gr[r] = 0;
    if (j < D) {
      c.xs[j] = x[r];
      c.gx[j] = 0;
    }
  }
  __syncwarp();
  const real lp = M::template terms<real>(x, gr, c);  // already summed over the warp
  __syncwarp();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = c.lane + 32 * r;
    g[r] = (j < D) ? (gr[r] + c.gx[j]) * dxdz[r] : (real)0;
  }
  return lp;
}

// ------------------------------------------------------------------------------------------
// parity kernel (a): log-density and gradient for C independent states
// ------------------------------------------------------------------------------------------
template <class M, typename real, bool STAGED>
__global__ void __launch_bounds__(128) logp_grad_kernel(ModelArgs<real> A, int C, const real* __restrict__ theta,
                                                        real* __restrict__ lp_out, real* __restrict__ grad_out) {
  constexpr int R = M::R, D = M::D, DP = M::DP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<real, STAGED>(A, smem_raw, DP);
  const int lane = c.lane;
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= C) return;
  real th[R], g[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + 32 * r;
    th[r] = (j < D) ? theta[(size_t)chain * D + j] : (real)0;
  }
  const real lp = logp_grad<M, real>(th, g, c);
  if (lane == 0) lp_out[chain] = lp;
  if (grad_out) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + 32 * r;
      if (j < D) grad_out[(size_t)chain * D + j] = g[r];
    }
  }
}

// traced outputs (deterministics and back-transformed variables): one thread per state
template <class M, typename real>
__global__ void dets_kernel(ModelArgs<real> A, int64_t n_states, const real* __restrict__ theta,
                            real* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_states) return;
  M::template dets<real>(theta + m * M::D, out + m * M::DET_ROW, A);
}

// ------------------------------------------------------------------------------------------
// run-time parameters shared by the HMC and NUTS kernels
// ------------------------------------------------------------------------------------------
template <typename real> struct RunParams {
  ModelArgs<real> A;
  int C, t_begin, t_count, B, S, max_depth, num_leapfrog, chain_offset;
  int adapt_enabled, adapt_pooled, num_adapt;
  real max_energy_diff, target_accept, shrinkage, smoothing, decay;
  uint32_t seed_lo, seed_hi;
  real* th;
  real* gr;
  real* lp;
  real* da;
  real* accept;
  const real* step_scale;
  const real* momenta;
  const real* uniforms;
  real* scratch;
  real* states;
  real* lp_out;
  int32_t* leapfrogs;
  uint8_t* diverging;
  real* energy;
  real* log_accept;
  int32_t* depth;
  int64_t* total_leapfrogs;
  real* traj;
};

template <typename real> __host__ inline RunParams<real> make_run_params(const pm4_mod_run_t& r) {
  RunParams<real> p;
  p.A = make_model_args<real>(r.A);
  p.C = r.C; p.t_begin = r.t_begin; p.t_count = r.t_count; p.B = r.B; p.S = r.S;
  p.max_depth = r.max_depth; p.num_leapfrog = r.num_leapfrog; p.chain_offset = r.chain_offset;
  p.adapt_enabled = r.adapt_enabled; p.adapt_pooled = r.adapt_pooled; p.num_adapt = r.num_adapt;
  p.max_energy_diff = (real)r.max_energy_diff; p.target_accept = (real)r.target_accept;
  p.shrinkage = (real)r.shrinkage; p.smoothing = (real)r.smoothing; p.decay = (real)r.decay;
  p.seed_lo = r.seed_lo; p.seed_hi = r.seed_hi;
  p.th = (real*)r.th; p.gr = (real*)r.gr; p.lp = (real*)r.lp; p.da = (real*)r.da;
  p.accept = (real*)r.accept;
  p.step_scale = (const real*)r.step_scale; p.momenta = (const real*)r.momenta;
  p.uniforms = (const real*)r.uniforms; p.scratch = (real*)r.scratch;
  p.states = (real*)r.states; p.lp_out = (real*)r.lp_out; p.leapfrogs = r.leapfrogs;
  p.diverging = r.diverging; p.energy = (real*)r.energy; p.log_accept = (real*)r.log_accept;
  p.depth = r.depth; p.total_leapfrogs = r.total_leapfrogs; p.traj = (real*)r.traj;
  return p;
}

// dual-averaging state of one chain (or of the pool): tfp DualAveragingStepSizeAdaptation
// (PyMC4 wiring /root/reference/pymc4/mcmc/samplers.py:400-405)
template <typename real> struct DualAvg {
  real eps, error_sum, log_avg, log_shrink_target;
  __device__ __forceinline__ void load(const real* p) {
    eps = p[0]; error_sum = p[1]; log_avg = p[2]; log_shrink_target = p[3];
  }
  __device__ __forceinline__ void store(real* p) const {
    p[0] = eps; p[1] = error_sum; p[2] = log_avg; p[3] = log_shrink_target;
  }
  // called after transition t (0-based) with the mean acceptance probability
  template <class P> __device__ __forceinline__ void update(const P& a, int t, real accept) {
    const real step = (real)(t + 1);
    error_sum += a.target_accept - accept;
    const real log_eps = log_shrink_target - error_sum * m_sqrt(step) / ((a.smoothing + step) * a.shrinkage);
    const real eta = m_pow(step, -a.decay);
    log_avg = eta * log_eps + ((real)1 - eta) * log_avg;
    if (t < a.num_adapt) eps = m_exp(log_eps);
    else if (t == a.num_adapt) eps = m_exp(log_avg);
  }
};

// one-block kernel: pooled (cross-chain) dual-averaging update after a lock-step transition.
// reduce_logmeanexp over chains followed by exp = arithmetic mean of the acceptance probabilities.
template <typename real>
__global__ void __launch_bounds__(1024) da_pooled_kernel(RunParams<real> a, int t) {
  __shared__ double part[32];
  double s = 0;  // fixed-order, deterministic reduction
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) s += (double)a.accept[c];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    if (threadIdx.x == 0) {
      DualAvg<real> d;
      d.load(a.da);
      d.update(a, t, (real)(v / (double)a.C));
      d.store(a.da);
    }
  }
}

template <typename real> __global__ void export_eps_kernel(RunParams<real> a, real* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.C) out[c] = a.da[a.adapt_pooled ? 0 : (size_t)c * 4];
}

// initial log-density / gradient of every chain (tfp bootstrap_results) + dual-averaging init
template <class M, typename real, bool STAGED>
__global__ void __launch_bounds__(128) init_kernel(RunParams<real> a, const real* __restrict__ theta0, real eps0) {
  constexpr int R = M::R, D = M::D, DP = M::DP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<real, STAGED>(a.A, smem_raw, DP);
  const int lane = c.lane;
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= a.C) return;
  real th[R], g[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + 32 * r;
    th[r] = (j < D) ? theta0[(size_t)chain * D + j] : (real)0;
  }
  const real lp = logp_grad<M, real>(th, g, c);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    a.th[(size_t)chain * DP + lane + 32 * r] = th[r];
    a.gr[(size_t)chain * DP + lane + 32 * r] = g[r];
  }
  if (lane == 0) {
    a.lp[chain] = lp;
    if (a.total_leapfrogs) a.total_leapfrogs[chain] = 0;
    if (!a.adapt_pooled || chain == 0) {
      real* d = a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4);
      d[0] = eps0; d[1] = 0; d[2] = 0; d[3] = m_log((real)10 * eps0);
    }
  }
}

// momentum of this chain at transition t: Philox block (q*32 + lane) feeds registers 4q..4q+3
template <class M, typename real>
__device__ __forceinline__ void draw_momentum(const RunParams<real>& a, int chain, int t, int lane, real (&p)[M::R]) {
  constexpr int R = M::R, D = M::D;
  if (a.momenta) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + 32 * r;
      p[r] = (j < D) ? a.momenta[((size_t)t * a.C + chain) * D + j] : (real)0;
    }
    return;
  }
#pragma unroll
  for (int q = 0; q < (R + 3) / 4; ++q) {
    const U4 u = philox4x32_10((uint32_t)(a.chain_offset + chain), (uint32_t)t, (uint32_t)(q * 32 + lane),
                               RNG_MOMENTUM, a.seed_lo, a.seed_hi);
    real n[4];
    box_muller<real>(u.x, u.y, n[0], n[1]);
    box_muller<real>(u.z, u.w, n[2], n[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = 4 * q + e;
      if (r < R) p[r] = (lane + 32 * r < D) ? n[e] : (real)0;
    }
  }
}

// one leapfrog step: tfp SimpleLeapfrogIntegrator with num_steps = 1:
// half kick, drift, gradient, full kick, half un-kick.  es[r] = eps * per-dim scale, sgn = +-1.
template <class M, typename real>
__device__ __forceinline__ real leapfrog(real (&th)[M::R], real (&p)[M::R], real (&g)[M::R],
                                         const real (&es)[M::R], real sgn, Ctx<real>& c) {
#pragma unroll
  for (int r = 0; r < M::R; ++r) {
    const real e = sgn * es[r];
    p[r] = p[r] + ((real)0.5 * e) * g[r];
    th[r] = th[r] + e * p[r];
  }
  const real lp = logp_grad<M, real>(th, g, c);
#pragma unroll
  for (int r = 0; r < M::R; ++r) {
    const real e = sgn * es[r];
    p[r] = p[r] + e * g[r];
    p[r] = p[r] - ((real)0.5 * e) * g[r];
  }
  return lp;
}

template <int R, typename real> __device__ __forceinline__ real dot(const real (&a)[R], const real (&b)[R]) {
  real s = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) s += a[r] * b[r];
  return warp_sum(s);
}

// tree scratch lives in global memory (L2-resident: every access is one coalesced line per register)
template <typename real> __device__ __forceinline__ real ldq(const real* p) { return __ldcg(p); }
template <typename real> __device__ __forceinline__ void stq(real* p, real v) { __stcg(p, v); }

// ------------------------------------------------------------------------------------------
// NUTS: tfp NoUTurnSampler.one_step -> _loop_tree_doubling -> _build_sub_tree ->
// _loop_build_sub_tree in per-chain form, t_count transitions per launch.
// ------------------------------------------------------------------------------------------
template <class M, typename real, int MAXT, bool STAGED>
__global__ void __launch_bounds__(MAXT, 1) nuts_kernel(RunParams<real> a) {
  constexpr int R = M::R, D = M::D, DP = M::DP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<real, STAGED>(a.A, smem_raw, DP);
  const int lane = c.lane;
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= a.C) return;
  const real NEG_INF = -m_inf<real>();
  const size_t C = (size_t)a.C;
  const uint32_t rng_chain = (uint32_t)(a.chain_offset + chain);
  auto slot = [&](int s) -> real* { return a.scratch + ((size_t)s * C + chain) * DP + lane; };
  const int MEM_P = PM4_SLOT_MEM, MEM_RHO = PM4_SLOT_MEM + a.max_depth + 1;

  // chain state carried across transitions
  real cth[R], cg[R], scale[R];
  real clp = a.lp[chain];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + 32 * r;
    cth[r] = a.th[(size_t)chain * DP + j];
    cg[r] = a.gr[(size_t)chain * DP + j];
    scale[r] = (a.step_scale && j < D) ? a.step_scale[j] : (real)1;
  }
  DualAvg<real> da;
  da.load(a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4));
  long long total_leaps = 0;

  for (int t = a.t_begin; t < a.t_begin + a.t_count; ++t) {
    real es[R], th[R], p[R], g[R], rho[R], rho_sub[R];
#pragma unroll
    for (int r = 0; r < R; ++r) es[r] = da.eps * scale[r];
    draw_momentum<M, real>(a, chain, t, lane, p);
    const real H0 = clp - (real)0.5 * dot<R, real>(p, p);

    // merge uniforms (one per doubling, lane d holds depth d) and the direction bits
    const U4 um = philox4x32_10(rng_chain, (uint32_t)t, (uint32_t)lane, RNG_MERGE, a.seed_lo, a.seed_hi);
    const U4 ud = philox4x32_10(rng_chain, (uint32_t)t, 0u, RNG_DIRECTION, a.seed_lo, a.seed_hi);

    // both trajectory ends and the candidate start at the current point
#pragma unroll
    for (int r = 0; r < R; ++r) {
      th[r] = cth[r]; g[r] = cg[r]; rho[r] = p[r];
      stq(slot(PM4_SLOT_OTH_TH) + 32 * r, cth[r]);
      stq(slot(PM4_SLOT_OTH_P) + 32 * r, p[r]);
      stq(slot(PM4_SLOT_OTH_G) + 32 * r, cg[r]);
      stq(slot(PM4_SLOT_CAND_TH) + 32 * r, cth[r]);
      stq(slot(PM4_SLOT_CAND_G) + 32 * r, cg[r]);
    }
    real cur_lp = clp, oth_lp = clp;
    int cur_is_right = 1;
    real cand_lp = clp, cand_energy = H0, w = 0, ediff_sum = 0;
    int nleap = 0, depth = 0;
    bool cont = true, not_div = true;

    while (depth < a.max_depth && cont) {
      const int dir = (int)((ud.x >> depth) & 1u);
      if (dir != cur_is_right) {  // extend the other end: swap the register-resident end with the stored one
#pragma unroll
        for (int r = 0; r < R; ++r) {
          real v;
          v = ldq(slot(PM4_SLOT_OTH_TH) + 32 * r); stq(slot(PM4_SLOT_OTH_TH) + 32 * r, th[r]); th[r] = v;
          v = ldq(slot(PM4_SLOT_OTH_P) + 32 * r); stq(slot(PM4_SLOT_OTH_P) + 32 * r, p[r]); p[r] = v;
          v = ldq(slot(PM4_SLOT_OTH_G) + 32 * r); stq(slot(PM4_SLOT_OTH_G) + 32 * r, g[r]); g[r] = v;
        }
        const real tl = cur_lp; cur_lp = oth_lp; oth_lp = tl;
        cur_is_right = dir;
      }
      const real sgn = dir ? (real)1 : (real)-1;
      const int nsteps = 1 << depth;
#pragma unroll
      for (int r = 0; r < R; ++r) rho_sub[r] = 0;
      real w_sub = NEG_INF, sub_lp = cur_lp, sub_energy = cur_lp, sub_ediff = 0;
      bool cont_sub = true;
      int sub_leaps = 0;
      U4 ul{0u, 0u, 0u, 0u};

      for (int i = 0; i < nsteps && cont_sub; ++i) {
        // leaf uniforms: two leaves share a Philox block, lane l holds pair-block (base + l)
        if ((i & 63) == 0)
          ul = philox4x32_10(rng_chain, (uint32_t)t, ((uint32_t)depth << 20) | (uint32_t)((i >> 1) + lane),
                             RNG_LEAF, a.seed_lo, a.seed_hi);
        cur_lp = leapfrog<M, real>(th, p, g, es, sgn, c);
        sub_leaps += 1;

        bool no_uturn = true;
        if (i & 1) {
          // U-turn checks against checkpoints [popcount(i) - trailing_ones(i), popcount(i))
          const int pc = __popc((unsigned)i), t1 = __ffs(~(unsigned)i) - 1;
#pragma unroll
          for (int r = 0; r < R; ++r) rho_sub[r] += p[r];
          for (int s = pc - t1; s < pc && no_uturn; ++s) {
            real d1 = 0, d2 = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const real mp = ldq(slot(MEM_P + s) + 32 * r);
              const real df = rho_sub[r] - ldq(slot(MEM_RHO + s) + 32 * r);
              d1 += df * mp;
              d2 += df * p[r];
            }
            d1 = warp_sum(d1);
            d2 = warp_sum(d2);
            no_uturn = (d1 > 0) && (d2 > 0);
          }
        } else {
          // even leaves are checkpointed into slot popcount(i): (p_i, rho before this leaf)
          const int s = __popc((unsigned)i);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            stq(slot(MEM_P + s) + 32 * r, p[r]);
            stq(slot(MEM_RHO + s) + 32 * r, rho_sub[r]);
            rho_sub[r] += p[r];
          }
        }

        real energy = cur_lp - (real)0.5 * dot<R, real>(p, p);
        if (m_isnan(energy)) energy = NEG_INF;
        const real dH = energy - H0;
        const bool not_divergent = (-dH < a.max_energy_diff);
        const real w_new = m_logaddexp(w_sub, dH);
        const real thresh = dH - w_new;
        const int src = (i >> 1) & 31;
        const uint32_t wa = __shfl_sync(FULL, (i & 1) ? ul.z : ul.x, src);
        const uint32_t wb = __shfl_sync(FULL, (i & 1) ? ul.w : ul.y, src);
        const real u = m_log1p(-u01<real>(wa, wb));
        if (u <= thresh) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            stq(slot(PM4_SLOT_SUB_TH) + 32 * r, th[r]);
            stq(slot(PM4_SLOT_SUB_G) + 32 * r, g[r]);
          }
          sub_lp = cur_lp;
          sub_energy = energy;
        }
        w_sub = w_new;
        not_div = not_div && not_divergent;
        if (not_divergent) sub_ediff += m_exp(dH < 0 ? dH : (real)0);
        cont_sub = no_uturn && not_divergent;
      }

      // merge the subtree into the trajectory (biased progressive sampling)
      ediff_sum += sub_ediff;
      nleap += sub_leaps;
      const real tree_w = cont_sub ? w_sub : NEG_INF;
      const real thresh = tree_w - w;
      const uint32_t ma = __shfl_sync(FULL, um.x, depth), mb = __shfl_sync(FULL, um.y, depth);
      const real umerge = m_log1p(-u01<real>(ma, mb));
      if ((umerge <= thresh) && cont_sub) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          stq(slot(PM4_SLOT_CAND_TH) + 32 * r, ldq(slot(PM4_SLOT_SUB_TH) + 32 * r));
          stq(slot(PM4_SLOT_CAND_G) + 32 * r, ldq(slot(PM4_SLOT_SUB_G) + 32 * r));
        }
        cand_lp = sub_lp;
        cand_energy = sub_energy;
      }
      w = m_logaddexp(tree_w, w);
      real d_cur = 0, d_oth = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        rho[r] += rho_sub[r];
        d_cur += rho[r] * p[r];
        d_oth += rho[r] * ldq(slot(PM4_SLOT_OTH_P) + 32 * r);
      }
      d_cur = warp_sum(d_cur);
      d_oth = warp_sum(d_oth);
      cont = cont_sub && (d_cur > 0) && (d_oth > 0);
      depth += 1;
    }

    // the transition's result is the candidate
#pragma unroll
    for (int r = 0; r < R; ++r) {
      cth[r] = ldq(slot(PM4_SLOT_CAND_TH) + 32 * r);
      cg[r] = ldq(slot(PM4_SLOT_CAND_G) + 32 * r);
    }
    clp = cand_lp;
    total_leaps += nleap;
    const real accept = ediff_sum / (real)nleap;
    if (a.adapt_enabled) {
      if (a.adapt_pooled) { if (lane == 0) a.accept[chain] = accept; }
      else da.update(a, t, accept);
    }
    if (t >= a.B) {
      const size_t k = (size_t)(t - a.B) * C + chain;
      if (a.states) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = lane + 32 * r;
          if (j < D) __stcs(a.states + k * D + j, cth[r]);
        }
      }
      if (lane == 0) {
        if (a.lp_out) a.lp_out[k] = cand_lp;
        if (a.leapfrogs) a.leapfrogs[k] = nleap;
        if (a.diverging) a.diverging[k] = not_div ? 0 : 1;
        if (a.energy) a.energy[k] = cand_energy;
        if (a.log_accept) a.log_accept[k] = m_log(accept);
        if (a.depth) a.depth[k] = depth;
      }
    }
  }

  // persist the chain for the next launch
#pragma unroll
  for (int r = 0; r < R; ++r) {
    a.th[(size_t)chain * DP + lane + 32 * r] = cth[r];
    a.gr[(size_t)chain * DP + lane + 32 * r] = cg[r];
  }
  if (lane == 0) {
    a.lp[chain] = clp;
    if (!a.adapt_pooled) da.store(a.da + (size_t)chain * 4);
    if (a.total_leapfrogs) a.total_leapfrogs[chain] += total_leaps;
  }
}

// ------------------------------------------------------------------------------------------
// fixed-L HMC: tfp HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC)
// ------------------------------------------------------------------------------------------
template <class M, typename real, int MAXT, bool STAGED>
__global__ void __launch_bounds__(MAXT, 1) hmc_kernel(RunParams<real> a) {
  constexpr int R = M::R, D = M::D, DP = M::DP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<real, STAGED>(a.A, smem_raw, DP);
  const int lane = c.lane;
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= a.C) return;
  const size_t C = (size_t)a.C;
  real cth[R], cg[R], scale[R];
  real clp = a.lp[chain];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = lane + 32 * r;
    cth[r] = a.th[(size_t)chain * DP + j];
    cg[r] = a.gr[(size_t)chain * DP + j];
    scale[r] = (a.step_scale && j < D) ? a.step_scale[j] : (real)1;
  }
  DualAvg<real> da;
  da.load(a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4));

  for (int t = a.t_begin; t < a.t_begin + a.t_count; ++t) {
    real es[R], th[R], p[R], g[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { es[r] = da.eps * scale[r]; th[r] = cth[r]; g[r] = cg[r]; }
    draw_momentum<M, real>(a, chain, t, lane, p);
    const real k0 = (real)0.5 * dot<R, real>(p, p);
    // half kick, L x (drift, gradient, full kick), half un-kick
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = p[r] + ((real)0.5 * es[r]) * g[r];
    real plp = clp;
    for (int l = 0; l < a.num_leapfrog; ++l) {
#pragma unroll
      for (int r = 0; r < R; ++r) th[r] = th[r] + es[r] * p[r];
      plp = logp_grad<M, real>(th, g, c);
#pragma unroll
      for (int r = 0; r < R; ++r) p[r] = p[r] + es[r] * g[r];
    }
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = p[r] - ((real)0.5 * es[r]) * g[r];
    const real k1 = (real)0.5 * dot<R, real>(p, p);
    real lar = (plp - clp) + (k0 - k1);
    if (!(lar == lar) || lar == m_inf<real>() || lar == -m_inf<real>()) lar = -m_inf<real>();
    real u;
    if (a.uniforms) u = a.uniforms[(size_t)t * C + chain];
    else {
      const U4 uu = philox4x32_10((uint32_t)(a.chain_offset + chain), (uint32_t)t, 0u, RNG_HMC, a.seed_lo, a.seed_hi);
      u = u01<real>(uu.x, uu.y);
    }
    const bool ok = m_log(u) < lar;
    if (a.traj) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = lane + 32 * r;
        if (j < D) a.traj[((size_t)t * C + chain) * D + j] = th[r];
      }
    }
    if (ok) {
#pragma unroll
      for (int r = 0; r < R; ++r) { cth[r] = th[r]; cg[r] = g[r]; }
      clp = plp;
    }
    const real accept = m_exp(lar < 0 ? lar : (real)0);
    if (a.adapt_enabled) {
      if (a.adapt_pooled) { if (lane == 0) a.accept[chain] = accept; }
      else da.update(a, t, accept);
    }
    if (t >= a.B) {
      const size_t k = (size_t)(t - a.B) * C + chain;
      if (a.states) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int j = lane + 32 * r;
          if (j < D) a.states[k * D + j] = cth[r];
        }
      }
      if (lane == 0) {
        if (a.lp_out) a.lp_out[k] = clp;
        if (a.log_accept) a.log_accept[k] = lar;
        if (a.diverging) a.diverging[k] = ok ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    a.th[(size_t)chain * DP + lane + 32 * r] = cth[r];
    a.gr[(size_t)chain * DP + lane + 32 * r] = cg[r];
  }
  if (lane == 0) {
    a.lp[chain] = clp;
    if (!a.adapt_pooled) da.store(a.da + (size_t)chain * 4);
  }
}

// ------------------------------------------------------------------------------------------
// host-side launch glue instantiated once per generated module
// ------------------------------------------------------------------------------------------
#define PM4_CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return (int)e_; } while (0)

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <class M, typename real, class Kernel>
inline int launch_chain_kernel(Kernel kernel, const RunParams<real>& p, int wpb, cudaStream_t st) {
  const size_t smem = cta_smem_bytes(p.A, M::DP, wpb);
  if (smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
  PM4_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (p.C + wpb - 1) / wpb;
  kernel<<<grid, wpb * 32, smem, st>>>(p);
  return (int)cudaGetLastError();
}

// warps per block -> the smallest compiled launch bound that fits (more registers per thread)
template <class M, typename real, bool NUTS> inline int dispatch_run(const pm4_mod_run_t& r, int wpb, cudaStream_t st) {
  RunParams<real> p = make_run_params<real>(r);
  if (wpb <= 0) {  // library default: one block per SM, up to 28 warps (= 28 chains) per block
    const int per_sm = (r.C + 147) / 148;
    wpb = per_sm <= 8 ? 8 : 28;
  }
  if (wpb > 28) wpb = 28;
  if (sizeof(real) == 8 && wpb > 8) wpb = 8;  // double precision is only instantiated for 256 threads
  // shrink the block until the per-chain state (and the staged pools) fit in shared memory
  while (wpb > 1 && cta_smem_bytes(p.A, M::DP, wpb) > SMEM_LIMIT) {
    if (p.A.stage && pool_bytes(p.A) > SMEM_LIMIT / 2) p.A.stage = 0;
    else wpb -= 1;
  }
#define PM4_LAUNCH(T)                                                                                       \
  do {                                                                                                      \
    if (p.A.stage)                                                                                          \
      return NUTS ? launch_chain_kernel<M, real>(nuts_kernel<M, real, T, true>, p, wpb, st)                 \
                  : launch_chain_kernel<M, real>(hmc_kernel<M, real, T, true>, p, wpb, st);                 \
    return NUTS ? launch_chain_kernel<M, real>(nuts_kernel<M, real, T, false>, p, wpb, st)                  \
                : launch_chain_kernel<M, real>(hmc_kernel<M, real, T, false>, p, wpb, st);                  \
  } while (0)
  if (wpb <= 8) PM4_LAUNCH(256);
  if constexpr (sizeof(real) == 4) PM4_LAUNCH(896);
#undef PM4_LAUNCH
  return (int)cudaErrorInvalidConfiguration;
}

}  // namespace pm4

// ------------------------------------------------------------------------------------------
// module entry points (csrc/pm4_module.h).  PM4_MODEL is the generated model struct.
// ------------------------------------------------------------------------------------------
#ifdef PM4_MODEL
#ifndef PM4_HAS_F64
#define PM4_HAS_F64 0
#endif

extern "C" int pm4m_info(pm4_mod_info_t* out) {
  out->abi = PM4_MODULE_ABI;
  out->D = PM4_MODEL::D;
  out->R = PM4_MODEL::R;
  out->DP = PM4_MODEL::DP;
  out->det_row = PM4_MODEL::DET_ROW;
  out->has_f64 = PM4_HAS_F64;
  out->n_terms = PM4_MODEL::N_TERMS;
  out->n_ops = PM4_MODEL::N_OPS;
  out->struct_hash = PM4_MODEL::HASH;
  return 0;
}

namespace pm4 {
template <typename real>
int mod_logp_grad(const pm4_mod_args_t* A, int C, const void* theta, void* lp, void* grad, cudaStream_t st) {
  using M = PM4_MODEL;
  ModelArgs<real> ma = make_model_args<real>(*A);
  const int wpb = 4;
  if (cta_smem_bytes(ma, M::DP, wpb) > SMEM_LIMIT) ma.stage = 0;
  const size_t smem = cta_smem_bytes(ma, M::DP, wpb);
  const unsigned grid = (unsigned)((C + wpb - 1) / wpb);
  if (ma.stage) {
    PM4_CUDA_OK(cudaFuncSetAttribute(logp_grad_kernel<M, real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logp_grad_kernel<M, real, true><<<grid, wpb * 32, smem, st>>>(ma, C, (const real*)theta, (real*)lp, (real*)grad);
  } else {
    PM4_CUDA_OK(cudaFuncSetAttribute(logp_grad_kernel<M, real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logp_grad_kernel<M, real, false><<<grid, wpb * 32, smem, st>>>(ma, C, (const real*)theta, (real*)lp, (real*)grad);
  }
  return (int)cudaGetLastError();
}
template <typename real>
int mod_dets(const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, cudaStream_t st) {
  using M = PM4_MODEL;
  if (M::DET_ROW == 0 || n == 0) return 0;
  ModelArgs<real> ma = make_model_args<real>(*A);
  dets_kernel<M, real><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(ma, n, (const real*)theta, (real*)out);
  return (int)cudaGetLastError();
}
template <typename real> int mod_init(const pm4_mod_run_t* r, const void* theta0, double eps0, cudaStream_t st) {
  using M = PM4_MODEL;
  RunParams<real> p = make_run_params<real>(*r);
  const int wpb = 4;
  if (cta_smem_bytes(p.A, M::DP, wpb) > SMEM_LIMIT) p.A.stage = 0;
  const size_t smem = cta_smem_bytes(p.A, M::DP, wpb);
  const unsigned grid = (unsigned)((r->C + wpb - 1) / wpb);
  if (p.A.stage) {
    PM4_CUDA_OK(cudaFuncSetAttribute(init_kernel<M, real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    init_kernel<M, real, true><<<grid, wpb * 32, smem, st>>>(p, (const real*)theta0, (real)eps0);
  } else {
    PM4_CUDA_OK(cudaFuncSetAttribute(init_kernel<M, real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    init_kernel<M, real, false><<<grid, wpb * 32, smem, st>>>(p, (const real*)theta0, (real)eps0);
  }
  return (int)cudaGetLastError();
}
}  // namespace pm4

#define PM4_BY_DTYPE(call32, call64)                 \
  if (dtype == 0) return call32;                     \
  if (dtype == 1) { if (PM4_HAS_F64) return call64; return -38; } \
  return -22

#if PM4_HAS_F64
#define PM4_F64_CALL(x) x
#else
#define PM4_F64_CALL(x) (-38)
#endif

extern "C" int pm4m_logp_grad(int dtype, const pm4_mod_args_t* A, int C, const void* theta, void* lp,
                              void* grad, void* stream) {
  PM4_BY_DTYPE(pm4::mod_logp_grad<float>(A, C, theta, lp, grad, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_logp_grad<double>(A, C, theta, lp, grad, (cudaStream_t)stream)));
}
extern "C" int pm4m_dets(int dtype, const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, void* stream) {
  PM4_BY_DTYPE(pm4::mod_dets<float>(A, n, theta, out, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_dets<double>(A, n, theta, out, (cudaStream_t)stream)));
}
extern "C" int pm4m_init(int dtype, const pm4_mod_run_t* r, const void* theta0, double eps0, void* stream) {
  PM4_BY_DTYPE(pm4::mod_init<float>(r, theta0, eps0, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_init<double>(r, theta0, eps0, (cudaStream_t)stream)));
}
extern "C" int pm4m_nuts(int dtype, const pm4_mod_run_t* r, int wpb, void* stream) {
  PM4_BY_DTYPE((pm4::dispatch_run<PM4_MODEL, float, true>(*r, wpb, (cudaStream_t)stream)),
               PM4_F64_CALL((pm4::dispatch_run<PM4_MODEL, double, true>(*r, wpb, (cudaStream_t)stream))));
}
extern "C" int pm4m_hmc(int dtype, const pm4_mod_run_t* r, int wpb, void* stream) {
  PM4_BY_DTYPE((pm4::dispatch_run<PM4_MODEL, float, false>(*r, wpb, (cudaStream_t)stream)),
               PM4_F64_CALL((pm4::dispatch_run<PM4_MODEL, double, false>(*r, wpb, (cudaStream_t)stream))));
}
extern "C" int pm4m_da_pooled(int dtype, const pm4_mod_run_t* r, int t, void* stream) {
  if (dtype == 0) { pm4::da_pooled_kernel<float><<<1, 1024, 0, (cudaStream_t)stream>>>(pm4::make_run_params<float>(*r), t); return (int)cudaGetLastError(); }
  if (dtype == 1) { pm4::da_pooled_kernel<double><<<1, 1024, 0, (cudaStream_t)stream>>>(pm4::make_run_params<double>(*r), t); return (int)cudaGetLastError(); }
  return -22;
}
extern "C" int pm4m_export_eps(int dtype, const pm4_mod_run_t* r, void* out, void* stream) {
  const unsigned grid = (unsigned)((r->C + 255) / 256);
  if (dtype == 0) { pm4::export_eps_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(pm4::make_run_params<float>(*r), (float*)out); return (int)cudaGetLastError(); }
  if (dtype == 1) { pm4::export_eps_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(pm4::make_run_params<double>(*r), (double*)out); return (int)cudaGetLastError(); }
  return -22;
}
#endif  // PM4_MODEL
