// pm4_kernels.cuh -- the sampler kernels, templated on a generated `Model` and on the arithmetic
// type.  One TEAM of T warps owns one chain (T = Model::T): every D-vector of the chain (position,
// momentum, gradient, trajectory momentum sums) is spread over the TL = 32 T lanes of the team, lane
// tl holding elements tl, tl + TL, ... in registers (M::R per vector).  Reductions over dimensions /
// observations are a packed transposing warp butterfly followed by one exchange through shared
// memory, so every lane of the team sees bit-identical sums and all sampler decisions are
// team-uniform.  The whole NUTS tree state of the chain (far trajectory end, candidates, U-turn
// checkpoints) lives in the team's shared memory; observed-data records and execution plans of
// gather terms are copied once per CTA into shared memory with a TMA bulk copy (cp.async.bulk +
// mbarrier).  Thread blocks are persistent (one per SM): a team that finishes a chain takes the next
// one from a global queue, longest expected chains first.
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
//   then per team            xs[DP] | gx[DP] (NEEDS_GX) | part[part_reals] | red[8 T] | ctl[4] | scratch[n_slots DS]
// ------------------------------------------------------------------------------------------
template <typename real> __host__ __device__ inline size_t pool_bytes(const ModelArgs<real>& A) {
  return A.stage ? 16 + (size_t)A.n_planf * sizeof(real) + (size_t)A.n_plani * sizeof(int32_t) : 0;
}
// reals of one team without tree scratch
template <class M, typename real> __host__ __device__ inline size_t team_base_reals(const ModelArgs<real>& A) {
  return (size_t)M::DP * (M::NEEDS_GX ? 2 : 1) + (((size_t)A.part_reals + 3) & ~(size_t)3) + 8 * (size_t)M::T + 4;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Build the per-lane context; when STAGED, thread 0 issues the TMA bulk copies of the plan pools
// and every thread of the CTA waits on the mbarrier.  Must be called by ALL threads of the CTA.
// `team_reals` = reals per team including tree scratch (the host computed the same number).
template <class M, typename real, bool STAGED>
__device__ __forceinline__ Ctx<real> make_ctx(const ModelArgs<real>& A, unsigned char* smem_raw, int team_reals) {
  constexpr int T = M::T, TL = 32 * T;
  Ctx<real> c;
  c.tl = threadIdx.x % TL;
  c.lane = threadIdx.x & 31;
  c.warp = c.tl >> 5;
  const int tix = threadIdx.x / TL;
  c.bar = tix;  // barrier 0 is used CTA-wide only before the teams start
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
  real* w = reinterpret_cast<real*>(smem_raw + base) + (size_t)tix * (size_t)team_reals;
  c.xs = w;
  w += M::DP;
  c.gx = w;
  if (M::NEEDS_GX) w += M::DP;
  c.part = w;
  w += (A.part_reals + 3) & ~3;
  c.red = w;
  w += 8 * T;
  c.ctl = reinterpret_cast<int32_t*>(w);
  w += 4;
  c.scr = w;
  // the lane that ends up holding packed-reduction value k parks it at spart[k][warp]
  constexpr int per = 32 / M::P;
  const int k = c.lane / per;
  c.my_spart = ((c.lane % per) == 0 && k <= M::N_ACC) ? c.pi[A.spart_off + k * T + c.warp] : -1;
  c.lp_first = c.pi[A.spart_off + M::N_ACC * T];
  // part entries nobody writes must read as zero for the whole launch
  for (int i = c.tl; i < A.part_reals; i += TL) c.part[i] = 0;
  team_sync<T>(c.bar);
  return c;
}

// Models whose state needs more than LEAN_R registers per lane and vector keep only theta, momentum
// and gradient in registers: the trajectory momentum sums live in tree-scratch slots and per-element
// constants (overflow table, step scale) are re-read where they are used.
#ifndef PM4_LEAN_R
#define PM4_LEAN_R 4
#endif
constexpr int LEAN_R = PM4_LEAN_R;

// threads per block: the small variant (double precision, few chains, stateless kernels) and the big
// single-precision variant; both hold at least one whole team
#ifndef PM4_MAXT
#define PM4_MAXT 512
#endif
template <class M> constexpr int small_threads() { return 32 * M::T > 256 ? 32 * M::T : 256; }
template <class M> constexpr int big_threads() { return 32 * M::T > PM4_MAXT ? 32 * M::T : PM4_MAXT; }

// gradient-element ownership of this lane: elem_tab entries of elements tl, tl + TL, ... (overflow
// vectors of the part array; 0 for almost every element).  Small models hold them in registers.
template <class M, typename real> struct Own {
  int reg[M::R > LEAN_R ? 1 : M::R];
  const int32_t* tab;
  __device__ __forceinline__ void load(const Ctx<real>& c) {
    tab = c.pi + c.A.elem_off + c.tl;
    if constexpr (M::R <= LEAN_R) {
#pragma unroll
      for (int r = 0; r < M::R; ++r) reg[r] = tab[32 * M::T * r];
    }
  }
  __device__ __forceinline__ int operator()(int r) const {
    if constexpr (M::R <= LEAN_R) return reg[r];
    else return tab[32 * M::T * r];
  }
};

// ------------------------------------------------------------------------------------------
// joint log-density + gradient in unconstrained space for the chain owned by this team.
// `th`/`g` are the lane-distributed theta and gradient; returns the (team-uniform) log-density.
// Two team barriers: constrained values visible -> terms; partial sums parked -> owners add up.
// ------------------------------------------------------------------------------------------
template <class M, typename real>
__device__ __forceinline__ real logp_grad(const real (&th)[M::R], real (&g)[M::R], Ctx<real>& c,
                                          const Own<M, real>& own) {
  constexpr int R = M::R, T = M::T, TL = 32 * T;
  real x[R], dxdz[R], gr[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = c.tl + TL * r;
    M::template constrain<real>(r, c.tl, th[r], x[r], dxdz[r]);
    gr[r] = 0;
    c.xs[j] = x[r];
    if (M::NEEDS_GX) c.gx[j] = 0;
  }
  team_sync<T>(c.bar);
  M::template terms<real>(x, gr, c);  // parks log-density and adjoint partial sums in c.part
  team_sync<T>(c.bar);
  const real lp = sum_consecutive<T, real>(c.part + c.lp_first);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    // the element's PK fixed entries with one (vector) load, then the (rare) overflow vectors
    real s = gr[r] + sum_consecutive<M::PK, real>(c.part + M::PK * (c.tl + TL * r));
    const int ow = own(r);
    if (ow) {
      const real* ov = c.part + (ow >> ELEM_CNT_BITS);
      const int cnt = ow & ((1 << ELEM_CNT_BITS) - 1);
#pragma unroll 1
      for (int k = 0; k < cnt; ++k) s += sum_consecutive<4, real>(ov + 4 * k);
    }
    if (M::NEEDS_GX) s += c.gx[c.tl + TL * r];
    g[r] = s * dxdz[r];
  }
  return lp;
}

// ------------------------------------------------------------------------------------------
// parity kernel (a): log-density and gradient for C independent states
// ------------------------------------------------------------------------------------------
template <class M, typename real, bool STAGED>
__global__ void __launch_bounds__(small_threads<M>()) logp_grad_kernel(ModelArgs<real> A, int C, int team_reals,
                                                        const real* __restrict__ theta,
                                                        real* __restrict__ lp_out, real* __restrict__ grad_out) {
  constexpr int R = M::R, D = M::D, T = M::T, TL = 32 * T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<M, real, STAGED>(A, smem_raw, team_reals);
  const int tl = c.tl;
  const int chain = blockIdx.x * (blockDim.x / TL) + threadIdx.x / TL;
  if (chain >= C) return;
  Own<M, real> own;
  own.load(c);
  real th[R], g[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = tl + TL * r;
    th[r] = (j < D) ? theta[(size_t)chain * D + j] : (real)0;
  }
  const real lp = logp_grad<M, real>(th, g, c, own);
  if (tl == 0) lp_out[chain] = lp;
  if (grad_out) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = tl + TL * r;
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

// pointwise log-likelihood of the observed variables: one warp per state, lanes stride over the
// observations, coalesced stores; states are taken grid-stride
template <class M, typename real>
__global__ void __launch_bounds__(256) loglik_kernel(ModelArgs<real> A, int64_t n_states, const real* __restrict__ theta,
                                                     real* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < n_states; m += warps)
    M::template loglik<real>(theta + m * M::D, out + m * M::LL_ROW, A, lane);
}

// posterior predictive: one draw of every observed variable per state (same geometry as loglik_kernel)
template <class M, typename real>
__global__ void __launch_bounds__(256) predict_kernel(ModelArgs<real> A, int64_t n_states, const real* __restrict__ theta,
                                                      real* __restrict__ out, uint32_t k0, uint32_t k1) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < n_states; m += warps)
    M::template predict<real>(theta + m * M::D, out + m * M::LL_ROW, A, lane, (uint32_t)m, (uint32_t)((uint64_t)m >> 32), k0, k1);
}

// prior predictive: one ancestral sample of the whole model per warp
template <class M, typename real>
__global__ void __launch_bounds__(256) prior_kernel(ModelArgs<real> A, int64_t n_samples, real* __restrict__ x_out,
                                                    real* __restrict__ obs_out, uint32_t k0, uint32_t k1) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); m < n_samples; m += warps)
    M::template prior<real>(x_out + m * M::D, obs_out ? obs_out + m * M::LL_ROW : nullptr, A, lane, (uint32_t)m,
                            (uint32_t)((uint64_t)m >> 32), k0, k1);
}

// ------------------------------------------------------------------------------------------
// run-time parameters shared by the HMC and NUTS kernels
// ------------------------------------------------------------------------------------------
template <typename real> struct RunParams {
  ModelArgs<real> A;
  int C, t_begin, t_count, B, S, max_depth, num_leapfrog, chain_offset;
  int adapt_enabled, adapt_pooled, num_adapt, adapt_kind, random_walk;
  int team_reals, n_slots, n_fast, tb, rounds;
  int32_t* progress;
  int32_t* sync;
  const int32_t* da_t0;
  const int32_t* rw_kind;
  real max_energy_diff, target_accept, shrinkage, smoothing, decay, adapt_rate, rw_scale;
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
  int32_t* queue;
  const int32_t* order;
  int32_t* launch_leaps;
  void* comm_box[PM4_MAX_RANKS];
  int comm_rank, comm_world;
  uint32_t comm_epoch;
  int32_t* comm_error;
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
  p.adapt_kind = r.adapt_kind; p.random_walk = r.random_walk;
  p.adapt_rate = (real)r.adapt_rate; p.rw_scale = (real)r.rw_scale;
  p.team_reals = 0; p.n_slots = 0; p.n_fast = 0; p.tb = r.tb; p.progress = r.progress; p.rounds = r.rounds; p.sync = r.sync; p.da_t0 = r.da_t0; p.rw_kind = r.rw_kind;
  p.max_energy_diff = (real)r.max_energy_diff; p.target_accept = (real)r.target_accept;
  p.shrinkage = (real)r.shrinkage; p.smoothing = (real)r.smoothing; p.decay = (real)r.decay;
  p.seed_lo = r.seed_lo; p.seed_hi = r.seed_hi;
  p.th = (real*)r.th; p.gr = (real*)r.gr; p.lp = (real*)r.lp; p.da = (real*)r.da;
  p.accept = (real*)r.accept;
  p.step_scale = (const real*)r.step_scale; p.momenta = (const real*)r.momenta;
  p.uniforms = (const real*)r.uniforms; p.scratch = (real*)r.scratch;
  p.queue = r.queue; p.order = r.order; p.launch_leaps = r.launch_leaps;
  for (int k = 0; k < PM4_MAX_RANKS; ++k) p.comm_box[k] = r.comm_box[k];
  p.comm_rank = r.comm_rank; p.comm_world = r.comm_world; p.comm_epoch = r.comm_epoch; p.comm_error = r.comm_error;
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
    if (a.adapt_kind == 1) {  // tfp SimpleStepSizeAdaptation: multiplicative, no averaging
      if (t < a.num_adapt) eps = accept > a.target_accept ? eps * ((real)1 + a.adapt_rate) : eps / ((real)1 + a.adapt_rate);
      return;
    }
    const real step = (real)(t + 1 - (a.da_t0 ? __ldcg(a.da_t0) : 0));  // the recursion restarts at a mass-matrix window
    error_sum += a.target_accept - accept;
    const real log_eps = log_shrink_target - error_sum * m_sqrt(step) / ((a.smoothing + step) * a.shrinkage);
    const real eta = m_pow(step, -a.decay);
    log_avg = eta * log_eps + ((real)1 - eta) * log_avg;
    if (t < a.num_adapt) eps = m_exp(log_eps);
    else if (t == a.num_adapt) eps = m_exp(log_avg);
  }
};

// mailbox slot a rank publishes per lock-step transition (32 bytes; peers read it over NVLink)
struct CommSlot {
  double sum, count;
  unsigned long long seq;
  unsigned long long pad;
};

// one-block kernel: pooled (cross-chain) dual-averaging update after a lock-step transition.
// reduce_logmeanexp over chains followed by exp = arithmetic mean of the acceptance probabilities.
// With a communicator attached (chains of one pm.sample call sharded over the GPUs of the box, scalar
// step_size) the mean is over the chains of ALL ranks: the reduction and the exchange are one kernel --
// every rank writes (sum, count) into its own mailbox, reads its peers' mailboxes through NVLink peer
// memory, and adds them in rank order, so all ranks compute the bit-identical epsilon.  A peer that
// does not show up within ~30 s raises the device error word (once per run) instead of hanging the GPU.
template <typename real>
__device__ __forceinline__ void pooled_update(const RunParams<real>& a, int t, double* part /* shared, 32 doubles */) {
  double s = 0;  // fixed-order, deterministic reduction (the chains' accept words may come from other SMs: L2 loads)
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) s += (double)__ldcg(a.accept + c);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = (threadIdx.x < (blockDim.x >> 5)) ? part[threadIdx.x] : 0.0;
    v = warp_sum(v);
    double total = v, count = (double)a.C;
    // a peer already timed out in this run: keep going with the local chains instead of waiting again
    const bool comm_ok = a.comm_world > 1 && !(a.comm_error && *reinterpret_cast<volatile int32_t*>(a.comm_error) != 0);
    if (comm_ok) {
      const unsigned long long seq = ((unsigned long long)a.comm_epoch << 32) | (unsigned long long)(t + 1);
      const int slot = t % PM4_COMM_SLOTS;
      if (threadIdx.x == 0) {
        volatile CommSlot* mine = reinterpret_cast<volatile CommSlot*>(a.comm_box[a.comm_rank]) + slot;
        mine->sum = v;
        mine->count = (double)a.C;
        __threadfence_system();
        mine->seq = seq;
        __threadfence_system();
      }
      // lane r waits for rank r, then the sums are added in rank order
      double ps = 0, pc = 0;
      if ((int)threadIdx.x < a.comm_world) {
        volatile CommSlot* box = reinterpret_cast<volatile CommSlot*>(a.comm_box[threadIdx.x]) + slot;
        const long long t0 = clock64();
        bool ok = true;
        while (box->seq != seq) {
          if (clock64() - t0 > 60000000000ll) { ok = false; break; }  // ~30 s at 2 GHz
          __nanosleep(200);
        }
        __threadfence_system();
        if (ok) { ps = box->sum; pc = box->count; }
        else if (a.comm_error) atomicExch(a.comm_error, 1 + (int)threadIdx.x);
      }
      total = 0; count = 0;
      for (int r = 0; r < a.comm_world; ++r) {
        total += __shfl_sync(FULL, ps, r);
        count += __shfl_sync(FULL, pc, r);
      }
    }
    if (threadIdx.x == 0) {
      DualAvg<real> d;
      d.eps = __ldcg(a.da); d.error_sum = __ldcg(a.da + 1); d.log_avg = __ldcg(a.da + 2); d.log_shrink_target = __ldcg(a.da + 3);
      d.update(a, t, (real)(total / count));
      d.store(a.da);
    }
  }
  __syncthreads();
}

template <typename real>
__global__ void __launch_bounds__(1024) da_pooled_kernel(RunParams<real> a, int t) {
  __shared__ double part[32];
  pooled_update<real>(a, t, part);
}

template <typename real> __global__ void export_eps_kernel(RunParams<real> a, real* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < a.C) out[c] = a.da[a.adapt_pooled ? 0 : (size_t)c * 4];
}

// Scheduling order for the next launch: chains sorted by the leapfrogs they took during the last one,
// longest first (ties by chain index), so that the persistent teams start the expensive chains first.
// Rank by counting: O(C^2 / threads), deterministic.
static __global__ void __launch_bounds__(256) lpt_order_kernel(int C, const int32_t* __restrict__ work, int32_t* __restrict__ order) {
  __shared__ int32_t tile[1024];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t wi = i < C ? work[i] : 0;
  int rank = 0;
  for (int base = 0; base < C; base += 1024) {
    __syncthreads();
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) tile[k] = base + k < C ? work[base + k] : -1;
    __syncthreads();
    const int n = C - base < 1024 ? C - base : 1024;
    for (int k = 0; k < n; ++k) {
      const int32_t wk = tile[k];
      rank += (wk > wi) || (wk == wi && base + k < i);
    }
  }
  if (i < C) order[rank] = i;
}

// initial log-density / gradient of every chain (tfp bootstrap_results) + dual-averaging init
template <class M, typename real, bool STAGED>
__global__ void __launch_bounds__(small_threads<M>()) init_kernel(RunParams<real> a, const real* __restrict__ theta0, real eps0) {
  constexpr int R = M::R, D = M::D, DP = M::DP, T = M::T, TL = 32 * T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<M, real, STAGED>(a.A, smem_raw, a.team_reals);
  const int tl = c.tl;
  const int chain = blockIdx.x * (blockDim.x / TL) + threadIdx.x / TL;
  if (chain >= a.C) return;
  Own<M, real> own;
  own.load(c);
  real th[R], g[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = tl + TL * r;
    th[r] = (j < D) ? theta0[(size_t)chain * D + j] : (real)0;
  }
  const real lp = logp_grad<M, real>(th, g, c, own);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    a.th[(size_t)chain * DP + tl + TL * r] = th[r];
    a.gr[(size_t)chain * DP + tl + TL * r] = g[r];
  }
  if (tl == 0) {
    a.lp[chain] = lp;
    if (a.total_leapfrogs) a.total_leapfrogs[chain] = 0;
    if (!a.adapt_pooled || chain == 0) {
      real* d = a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4);
      d[0] = eps0; d[1] = 0; d[2] = 0; d[3] = m_log((real)10 * eps0);
    }
  }
}

// momentum of this chain at transition t.  Element j = lane' + 32 r' (lane' = j % 32) is output
// r' % 4 of Philox block ((r' / 4) * 32 + lane'): the stream does not depend on the team size.
template <class M, typename real>
__device__ __forceinline__ void draw_momentum(const RunParams<real>& a, int chain, int t, int tl, real (&p)[M::R]) {
  constexpr int R = M::R, D = M::D, TL = 32 * M::T;
  if (a.momenta) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = tl + TL * r;
      p[r] = (j < D) ? a.momenta[((size_t)t * a.C + chain) * D + j] : (real)0;
    }
    return;
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int j = tl + TL * r;
    const int ro = j >> 5, lo = j & 31;
    const U4 u = philox4x32_10((uint32_t)(a.chain_offset + chain), (uint32_t)t, (uint32_t)((ro >> 2) * 32 + lo),
                               RNG_MOMENTUM, a.seed_lo, a.seed_hi);
    real n0, n1;
    if (ro & 2) box_muller<real>(u.z, u.w, n0, n1);
    else box_muller<real>(u.x, u.y, n0, n1);
    p[r] = (j < D) ? ((ro & 1) ? n1 : n0) : (real)0;
  }
}

// one leapfrog step: tfp SimpleLeapfrogIntegrator with num_steps = 1:
// half kick, drift, gradient, full kick, half un-kick.  es[r] = signed eps * per-dim scale.
template <class M, typename real, class Es>
__device__ __forceinline__ real leapfrog(real (&th)[M::R], real (&p)[M::R], real (&g)[M::R],
                                         const Es& es, Ctx<real>& c, const Own<M, real>& own) {
#pragma unroll
  for (int r = 0; r < M::R; ++r) {
    const real e = es(r);
    p[r] = p[r] + ((real)0.5 * e) * g[r];
    th[r] = th[r] + e * p[r];
  }
  const real lp = logp_grad<M, real>(th, g, c, own);
#pragma unroll
  for (int r = 0; r < M::R; ++r) {
    const real e = es(r);
    p[r] = p[r] + e * g[r];
    p[r] = p[r] - ((real)0.5 * e) * g[r];
  }
  return lp;
}

// per-dimension step scale of this lane's elements: registers for small models, re-read otherwise
template <class M, typename real> struct StepScale {
  real reg[M::R > LEAN_R ? 1 : M::R];
  const real* tab;  // NULL: every scale is 1
  int tl;
  __device__ __forceinline__ void load(const real* scale, int tl_) {
    tab = scale;
    tl = tl_;
    if constexpr (M::R <= LEAN_R) {
#pragma unroll
      for (int r = 0; r < M::R; ++r) reg[r] = (scale && tl + 32 * M::T * r < M::D) ? scale[tl + 32 * M::T * r] : (real)1;
    }
  }
  __device__ __forceinline__ real operator()(int r) const {
    if constexpr (M::R <= LEAN_R) return reg[r];
    else return (tab && tl + 32 * M::T * r < M::D) ? __ldg(tab + tl + 32 * M::T * r) : (real)1;
  }
};

// next chain for this team from the global queue (-1: none left)
template <int T, typename real> __device__ __forceinline__ int next_chain(const RunParams<real>& a, const Ctx<real>& c) {
  int qi = 0;
  if constexpr (T == 1) {
    if (c.lane == 0) qi = atomicAdd(a.queue, 1);
    qi = __shfl_sync(FULL, qi, 0);
  } else {
    if (c.tl == 0) c.ctl[0] = atomicAdd(a.queue, 1);
    team_sync<T>(c.bar);
    qi = c.ctl[0];
    team_sync<T>(c.bar);
  }
  if (qi >= a.C) return -1;
  return a.order ? a.order[qi] : qi;
}

// ------------------------------------------------------------------------------------------
// NUTS: tfp NoUTurnSampler.one_step -> _loop_tree_doubling -> _build_sub_tree ->
// _loop_build_sub_tree in per-chain form, t_count transitions per chain per launch.
// ------------------------------------------------------------------------------------------
// SCR_SMEM: every tree-scratch slot of the team is in shared memory; otherwise the first a.n_fast
// slots are (the hottest ones come first) and the rest live in global memory (L2).
template <class M, typename real, int MAXT, bool STAGED, bool SCR_SMEM>
__global__ void __launch_bounds__(MAXT, 1) nuts_kernel(RunParams<real> a) {
  constexpr int R = M::R, D = M::D, DP = M::DP, DS = M::DS, T = M::T, TL = 32 * T;
  constexpr bool LEAN = R > LEAN_R;
  // slot numbering: lean models keep rho and rho_sub in the two hottest slots
  constexpr int S0 = LEAN ? 2 : 0;
  constexpr int SL_RHO = 0, SL_RHO_SUB = 1;
  constexpr int SL_OTH_TH = S0 + PM4_SLOT_OTH_TH, SL_OTH_P = S0 + PM4_SLOT_OTH_P, SL_OTH_G = S0 + PM4_SLOT_OTH_G;
  constexpr int SL_CAND_TH = S0 + PM4_SLOT_CAND_TH, SL_CAND_G = S0 + PM4_SLOT_CAND_G;
  constexpr int SL_SUB_TH = S0 + PM4_SLOT_SUB_TH, SL_SUB_G = S0 + PM4_SLOT_SUB_G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<M, real, STAGED>(a.A, smem_raw, a.team_reals);
  const int tl = c.tl, lane = c.lane;
  const real NEG_INF = -m_inf<real>();
  const size_t C = (size_t)a.C;
  real* const gscr = SCR_SMEM ? nullptr
                              : a.scratch + ((size_t)blockIdx.x * (blockDim.x / TL) + threadIdx.x / TL) *
                                                (size_t)(a.n_slots - a.n_fast) * DS;
  const int MEM_P = S0 + PM4_SLOT_MEM, MEM_RHO = MEM_P + a.max_depth + 1;
  // element tl + TL r exists for every lane when TL (r + 1) <= D, else only for the first lanes
  auto in = [&](int r) -> bool { return (TL * (r + 1) <= D) || (tl + TL * r < D); };
  auto sptr = [&](int s) -> real* {
    if constexpr (SCR_SMEM) return c.scr + s * DS + tl;
    else return (s < a.n_fast ? c.scr + s * DS : gscr + (size_t)(s - a.n_fast) * DS) + tl;
  };
  auto ldv = [&](int s, int r) -> real { return in(r) ? sptr(s)[TL * r] : (real)0; };
  auto stv = [&](int s, int r, real v) { if (in(r)) sptr(s)[TL * r] = v; };

  Own<M, real> own;
  own.load(c);
  StepScale<M, real> sc;
  sc.load(a.step_scale, tl);
  int rb = 0;  // which half of c.red the next team reduction uses

  for (;;) {
    const int chain = next_chain<T, real>(a, c);
    if (chain < 0) break;
    const uint32_t rng_chain = (uint32_t)(a.chain_offset + chain);
    // the chain's current point lives in the candidate slots between transitions
#pragma unroll
    for (int r = 0; r < R; ++r) {
      stv(SL_CAND_TH, r, a.th[(size_t)chain * DP + tl + TL * r]);
      stv(SL_CAND_G, r, a.gr[(size_t)chain * DP + tl + TL * r]);
    }
    real clp = a.lp[chain];
    DualAvg<real> da;
    da.load(a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4));
    int launch_leaps = 0;

    for (int t = a.t_begin; t < a.t_begin + a.t_count; ++t) {
      real th[R], p[R], g[R];
      real rho_reg[LEAN ? 1 : R], rho_sub_reg[LEAN ? 1 : R];
      auto rho = [&](int r) -> real { if constexpr (LEAN) return ldv(SL_RHO, r); else return rho_reg[r]; };
      auto set_rho = [&](int r, real v) { if constexpr (LEAN) stv(SL_RHO, r, v); else rho_reg[r] = v; };
      auto rho_sub = [&](int r) -> real { if constexpr (LEAN) return ldv(SL_RHO_SUB, r); else return rho_sub_reg[r]; };
      auto set_rho_sub = [&](int r, real v) { if constexpr (LEAN) stv(SL_RHO_SUB, r, v); else rho_sub_reg[r] = v; };
      const real eps = da.eps;
      draw_momentum<M, real>(a, chain, t, tl, p);
      if (a.step_scale) {  // scale 0 = frozen dimension (compound step): no momentum
#pragma unroll
        for (int r = 0; r < M::R; ++r) p[r] = sc(r) == (real)0 ? (real)0 : p[r];
      }
      real q[4];
      q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 0;
      // both trajectory ends and the candidate start at the current point
#pragma unroll
      for (int r = 0; r < R; ++r) {
        th[r] = ldv(SL_CAND_TH, r);
        g[r] = ldv(SL_CAND_G, r);
        set_rho(r, p[r]);
        q[0] += p[r] * p[r];
        stv(SL_OTH_TH, r, th[r]);
        stv(SL_OTH_P, r, p[r]);
        stv(SL_OTH_G, r, g[r]);
      }
      team_sum<T, 4, real>(q, c, c.red + rb * 4 * T);
      rb ^= 1;
      const real H0 = clp - (real)0.5 * q[0];

      // merge uniforms (one per doubling, lane d holds depth d) and the direction bits
      const U4 um = philox4x32_10(rng_chain, (uint32_t)t, (uint32_t)lane, RNG_MERGE, a.seed_lo, a.seed_hi);
      const U4 ud = philox4x32_10(rng_chain, (uint32_t)t, 0u, RNG_DIRECTION, a.seed_lo, a.seed_hi);

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
            v = ldv(SL_OTH_TH, r); stv(SL_OTH_TH, r, th[r]); th[r] = v;
            v = ldv(SL_OTH_P, r); stv(SL_OTH_P, r, p[r]); p[r] = v;
            v = ldv(SL_OTH_G, r); stv(SL_OTH_G, r, g[r]); g[r] = v;
          }
          const real sw = cur_lp; cur_lp = oth_lp; oth_lp = sw;
          cur_is_right = dir;
        }
        const int nsteps = 1 << depth;
#pragma unroll
        for (int r = 0; r < R; ++r) set_rho_sub(r, 0);
        const real eps_s = dir ? eps : -eps;
        auto es = [&](int r) -> real { return eps_s * sc(r); };
        real w_sub = NEG_INF, sub_lp = cur_lp, sub_energy = cur_lp, sub_ediff = 0;
        bool cont_sub = true;
        int sub_leaps = 0;
        U4 ul{0u, 0u, 0u, 0u};

        for (int i = 0; i < nsteps && cont_sub; ++i) {
          // leaf uniforms: two leaves share a Philox block, lane l holds pair-block (base + l)
          if ((i & 63) == 0)
            ul = philox4x32_10(rng_chain, (uint32_t)t, ((uint32_t)depth << 20) | (uint32_t)((i >> 1) + lane),
                               RNG_LEAF, a.seed_lo, a.seed_hi);
          cur_lp = leapfrog<M, real>(th, p, g, es, c, own);
          sub_leaps += 1;

          // one team reduction per leaf: kinetic energy and, on odd leaves, the U-turn test against the
          // checkpoint of the preceding even leaf
          const int pc = __popc((unsigned)i), t1 = __ffs(~(unsigned)i) - 1;
          q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 0;
          if (i & 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const real rs = rho_sub(r) + p[r];
              set_rho_sub(r, rs);
              const real mp = ldv(MEM_P + pc - 1, r);
              const real df = rs - ldv(MEM_RHO + pc - 1, r);
              q[0] += p[r] * p[r];
              q[1] += df * mp;
              q[2] += df * p[r];
            }
          } else {
            // even leaves are checkpointed into slot popcount(i): (p_i, rho before this leaf)
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const real before = rho_sub(r);
              stv(MEM_P + pc, r, p[r]);
              stv(MEM_RHO + pc, r, before);
              set_rho_sub(r, before + p[r]);
              q[0] += p[r] * p[r];
            }
          }
          team_sum<T, 4, real>(q, c, c.red + rb * 4 * T);
          rb ^= 1;
          bool no_uturn = true;
          if (i & 1) {
            no_uturn = (q[1] > 0) && (q[2] > 0);
            // remaining checkpoints [popcount(i) - trailing_ones(i), popcount(i) - 1), two per reduction
            for (int s = pc - 2; s >= pc - t1 && no_uturn; s -= 2) {
              const bool two = s - 1 >= pc - t1;
              real u4[4];
              u4[0] = 0; u4[1] = 0; u4[2] = 0; u4[3] = 0;
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const real rs = rho_sub(r);
                const real df = rs - ldv(MEM_RHO + s, r);
                u4[0] += df * ldv(MEM_P + s, r);
                u4[1] += df * p[r];
                if (two) {
                  const real df2 = rs - ldv(MEM_RHO + s - 1, r);
                  u4[2] += df2 * ldv(MEM_P + s - 1, r);
                  u4[3] += df2 * p[r];
                }
              }
              team_sum<T, 4, real>(u4, c, c.red + rb * 4 * T);
              rb ^= 1;
              no_uturn = (u4[0] > 0) && (u4[1] > 0) && (!two || ((u4[2] > 0) && (u4[3] > 0)));
            }
          }

          real energy = cur_lp - (real)0.5 * q[0];
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
              stv(SL_SUB_TH, r, th[r]);
              stv(SL_SUB_G, r, g[r]);
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
            stv(SL_CAND_TH, r, ldv(SL_SUB_TH, r));
            stv(SL_CAND_G, r, ldv(SL_SUB_G, r));
          }
          cand_lp = sub_lp;
          cand_energy = sub_energy;
        }
        w = m_logaddexp(tree_w, w);
        q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const real rr = rho(r) + rho_sub(r);
          set_rho(r, rr);
          q[0] += rr * p[r];
          q[1] += rr * ldv(SL_OTH_P, r);
        }
        team_sum<T, 4, real>(q, c, c.red + rb * 4 * T);
        rb ^= 1;
        cont = cont_sub && (q[0] > 0) && (q[1] > 0);
        depth += 1;
      }

      // the transition's result is the candidate
      clp = cand_lp;
      launch_leaps += nleap;
      const real accept = ediff_sum / (real)nleap;
      if (a.adapt_enabled) {
        if (a.adapt_pooled) { if (tl == 0) a.accept[chain] = accept; }
        else da.update(a, t, accept);
      }
      if (t >= a.B) {
        const size_t k = (size_t)(t - a.B) * C + chain;
        if (a.states) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int j = tl + TL * r;
            if (j < D) __stcs(a.states + k * D + j, ldv(SL_CAND_TH, r));
          }
        }
        if (tl == 0) {
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
      a.th[(size_t)chain * DP + tl + TL * r] = ldv(SL_CAND_TH, r);
      a.gr[(size_t)chain * DP + tl + TL * r] = ldv(SL_CAND_G, r);
    }
    if (tl == 0) {
      a.lp[chain] = clp;
      if (!a.adapt_pooled) da.store(a.da + (size_t)chain * 4);
      if (a.total_leapfrogs) a.total_leapfrogs[chain] += launch_leaps;
      if (a.launch_leaps) a.launch_leaps[chain] = launch_leaps;
    }
  }
}

// ------------------------------------------------------------------------------------------
// fixed-L HMC: tfp HamiltonianMonteCarlo = MetropolisHastings(UncalibratedHMC)
// ------------------------------------------------------------------------------------------
template <class M, typename real, int MAXT, bool STAGED>
__global__ void __launch_bounds__(MAXT, 1) hmc_kernel(RunParams<real> a) {
  constexpr int R = M::R, D = M::D, DP = M::DP, T = M::T, TL = 32 * T;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Ctx<real> c = make_ctx<M, real, STAGED>(a.A, smem_raw, a.team_reals);
  const int tl = c.tl;
  const size_t C = (size_t)a.C;
  Own<M, real> own;
  own.load(c);
  StepScale<M, real> sc;
  sc.load(a.step_scale, tl);
  int rb = 0;

  for (;;) {
    const int chain = next_chain<T, real>(a, c);
    if (chain < 0) break;
    // the chain's current point stays in global memory (a.th / a.gr): read at the start of a
    // transition, rewritten when a proposal is accepted
    real* const cth = a.th + (size_t)chain * DP + tl;
    real* const cg = a.gr + (size_t)chain * DP + tl;
    real clp = a.lp[chain];
    DualAvg<real> da;
    da.load(a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4));

    for (int t = a.t_begin; t < a.t_begin + a.t_count; ++t) {
      real th[R], p[R], g[R], q[4];
      const real eps = da.eps;
      auto es = [&](int r) -> real { return eps * sc(r); };
      draw_momentum<M, real>(a, chain, t, tl, p);
      if (a.step_scale) {  // scale 0 = frozen dimension (compound step): no momentum
#pragma unroll
        for (int r = 0; r < R; ++r) p[r] = sc(r) == (real)0 ? (real)0 : p[r];
      }
      q[0] = 0; q[1] = 0; q[2] = 0; q[3] = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        th[r] = cth[TL * r]; g[r] = cg[TL * r];
        q[0] += p[r] * p[r];
      }
      real plp = clp, lar;
      if (a.random_walk) {
        // tfp RandomWalkMetropolis: symmetric proposal; random_walk_normal_fn (state + scale * N(0, 1)) unless the
        // dimension carries one of the discrete proposals of the compound step (pm4_module.h: PM4_RW_*)
        if (a.rw_kind) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int j = tl + TL * r;
            if (j >= D || sc(r) == (real)0) continue;
            const int kd = __ldg(a.rw_kind + j), kind = kd & 0xff;
            const real z = p[r];
            if (kind == PM4_RW_BERNOULLI) th[r] = z > 0 ? (real)1 - th[r] : th[r];
            else if (kind == PM4_RW_CATEGORICAL) {
              const int K = kd >> 8;
              int k = (int)((real)K * (real)0.5 * (real)erfc((double)(-z) * 0.70710678118654752440));
              th[r] = (real)(k < 0 ? 0 : (k > K - 1 ? K - 1 : k));
            } else if (kind == PM4_RW_GAUSSIAN_ROUND) th[r] = (real)rint((double)(th[r] + a.rw_scale * z));
            else th[r] = th[r] + a.rw_scale * z;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) th[r] = th[r] + a.rw_scale * p[r];
        }
        plp = logp_grad<M, real>(th, g, c, own);
        lar = plp - clp;
      } else {
        // half kick, L x (drift, gradient, full kick), half un-kick
#pragma unroll
        for (int r = 0; r < R; ++r) p[r] = p[r] + ((real)0.5 * es(r)) * g[r];
        for (int l = 0; l < a.num_leapfrog; ++l) {
#pragma unroll
          for (int r = 0; r < R; ++r) th[r] = th[r] + es(r) * p[r];
          plp = logp_grad<M, real>(th, g, c, own);
#pragma unroll
          for (int r = 0; r < R; ++r) p[r] = p[r] + es(r) * g[r];
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          p[r] = p[r] - ((real)0.5 * es(r)) * g[r];
          q[1] += p[r] * p[r];
        }
        team_sum<T, 4, real>(q, c, c.red + rb * 4 * T);
        rb ^= 1;
        const real k0 = (real)0.5 * q[0], k1 = (real)0.5 * q[1];
        lar = (plp - clp) + (k0 - k1);
      }
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
          const int j = tl + TL * r;
          if (j < D) a.traj[((size_t)t * C + chain) * D + j] = th[r];
        }
      }
      if (ok) {
#pragma unroll
        for (int r = 0; r < R; ++r) { cth[TL * r] = th[r]; cg[TL * r] = g[r]; }
        clp = plp;
      }
      const real accept = m_exp(lar < 0 ? lar : (real)0);
      if (a.adapt_enabled) {
        if (a.adapt_pooled) { if (tl == 0) a.accept[chain] = accept; }
        else da.update(a, t, accept);
      }
      if (t >= a.B) {
        const size_t k = (size_t)(t - a.B) * C + chain;
        if (a.states) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int j = tl + TL * r;
            if (j < D) a.states[k * D + j] = ok ? th[r] : cth[TL * r];
          }
        }
        if (tl == 0) {
          if (a.lp_out) a.lp_out[k] = clp;
          if (a.log_accept) a.log_accept[k] = lar;
          if (a.diverging) a.diverging[k] = ok ? 1 : 0;
        }
      }
    }
    if (tl == 0) {
      a.lp[chain] = clp;
      if (!a.adapt_pooled) da.store(a.da + (size_t)chain * 4);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host-side launch glue instantiated once per generated module
// ------------------------------------------------------------------------------------------
#define PM4_CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return (int)e_; } while (0)

constexpr size_t SMEM_LIMIT = 227 * 1024;

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// Launch geometry of the persistent chain kernels: teams per block, whether the tree scratch lives
// in shared memory, bytes of dynamic shared memory.
struct Geometry {
  int teams, grid, team_reals, n_slots, n_fast, stage, scr_smem;
  size_t smem;
};

template <class M, typename real>
inline Geometry plan_geometry(ModelArgs<real>& A, int C, int wpb, int n_slots, int max_threads) {
  constexpr int T = M::T;
  Geometry g{};
  g.n_slots = n_slots;
  const int sms = sm_count();
  int max_teams = max_threads / (32 * T);
  if (T > 1 && max_teams > 16) max_teams = 16;  // named barriers 0..15
  if (max_teams < 1) max_teams = 1;
  int teams = wpb > 0 ? wpb / T : max_teams;
  if (teams > max_teams) teams = max_teams;
  if (teams < 1) teams = 1;
  // do not pack more teams into a block than the chains need to cover every SM
  const int per_sm = (C + sms - 1) / sms;
  if (wpb <= 0 && teams > per_sm) teams = per_sm;
  const size_t base = team_base_reals<M, real>(A), scr = (size_t)n_slots * M::DS;
  if (A.stage && pool_bytes(A) > SMEM_LIMIT / 2) A.stage = 0;
  // prefer everything on chip; give up teams before giving up the on-chip tree scratch, but keep at
  // least a quarter of the team slots busy
  g.scr_smem = 1;
  int t_scr = teams;
  while (t_scr > 1 && pool_bytes(A) + (size_t)t_scr * (base + scr) * sizeof(real) > SMEM_LIMIT) --t_scr;
  g.n_fast = n_slots;
  if (pool_bytes(A) + (size_t)t_scr * (base + scr) * sizeof(real) > SMEM_LIMIT || t_scr * 4 < teams) {
    // hybrid: as many of the hottest slots on chip as fit beside the per-team arrays, the rest in L2
    g.scr_smem = 0;
    while (teams > 1 && pool_bytes(A) + (size_t)teams * base * sizeof(real) > SMEM_LIMIT) --teams;
    if (pool_bytes(A) + (size_t)teams * base * sizeof(real) > SMEM_LIMIT) A.stage = 0;
    const size_t used = pool_bytes(A) + (size_t)teams * base * sizeof(real);
    const size_t room = used < SMEM_LIMIT ? (SMEM_LIMIT - used) / ((size_t)teams * M::DS * sizeof(real)) : 0;
    g.n_fast = (int)(room < (size_t)n_slots ? room : (size_t)n_slots);
  } else {
    teams = t_scr;
  }
  g.teams = teams;
  g.team_reals = (int)(base + (size_t)g.n_fast * M::DS);
  g.stage = A.stage;
  g.smem = pool_bytes(A) + (size_t)teams * g.team_reals * sizeof(real);
  int grid = (C + teams - 1) / teams;
  if (grid > sms) grid = sms;
  g.grid = grid;
  return g;
}

template <class M, typename real, class Kernel>
inline int launch_chain_kernel(Kernel kernel, const RunParams<real>& p, const Geometry& g, cudaStream_t st) {
  if (g.smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
  PM4_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
  PM4_CUDA_OK(cudaMemsetAsync(p.queue, 0, sizeof(int32_t), st));
  kernel<<<g.grid, g.teams * M::T * 32, g.smem, st>>>(p);
  return (int)cudaGetLastError();
}

template <class M, typename real> constexpr int max_threads_for() {
  return sizeof(real) == 8 ? small_threads<M>() : big_threads<M>();
}

// tree-scratch vector slots of a run (HMC keeps its state in registers)
template <class M> inline int scratch_slots(bool nuts, int max_depth) {
  return nuts ? (M::R > LEAN_R ? 2 : 0) + PM4_SLOT_MEM + 2 * (max_depth + 1) : 0;
}

template <class M, typename real, bool NUTS> inline int dispatch_run(const pm4_mod_run_t& r, int wpb, cudaStream_t st) {
  RunParams<real> p = make_run_params<real>(r);
  const int n_slots = scratch_slots<M>(NUTS, r.max_depth);
  Geometry g = plan_geometry<M, real>(p.A, r.C, wpb, n_slots, max_threads_for<M, real>());
  p.team_reals = g.team_reals;
  p.n_slots = n_slots;
  p.n_fast = g.n_fast;
  if (NUTS && !g.scr_smem && !p.scratch) return (int)cudaErrorInvalidValue;
  const bool small = g.teams * M::T * 32 <= small_threads<M>();
#define PM4_LAUNCH(MAXT)                                                                                              \
  do {                                                                                                                \
    if constexpr (NUTS) {                                                                                             \
      if (p.A.stage) {                                                                                                \
        if (g.scr_smem) return launch_chain_kernel<M, real>(nuts_kernel<M, real, MAXT, true, true>, p, g, st);        \
        return launch_chain_kernel<M, real>(nuts_kernel<M, real, MAXT, true, false>, p, g, st);                       \
      }                                                                                                               \
      if (g.scr_smem) return launch_chain_kernel<M, real>(nuts_kernel<M, real, MAXT, false, true>, p, g, st);         \
      return launch_chain_kernel<M, real>(nuts_kernel<M, real, MAXT, false, false>, p, g, st);                        \
    } else {                                                                                                          \
      if (p.A.stage) return launch_chain_kernel<M, real>(hmc_kernel<M, real, MAXT, true>, p, g, st);                  \
      return launch_chain_kernel<M, real>(hmc_kernel<M, real, MAXT, false>, p, g, st);                                \
    }                                                                                                                 \
  } while (0)
  if (small) PM4_LAUNCH(small_threads<M>());
  if constexpr (sizeof(real) == 4) PM4_LAUNCH(big_threads<M>());
#undef PM4_LAUNCH
  return (int)cudaErrorInvalidConfiguration;
}

}  // namespace pm4

#include "pm4_warp.cuh"

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
  out->T = PM4_MODEL::T;
  out->R = PM4_MODEL::R;
  out->DP = PM4_MODEL::DP;
  out->DS = PM4_MODEL::DS;
  out->det_row = PM4_MODEL::DET_ROW;
  out->ll_row = PM4_MODEL::LL_ROW;
  out->has_f64 = PM4_HAS_F64;
  out->n_terms = PM4_MODEL::N_TERMS;
  out->n_ops = PM4_MODEL::N_OPS;
  out->struct_hash = PM4_MODEL::HASH;
  out->caps = pm4::use_warp_path<PM4_MODEL>() ? PM4_CAP_PERSISTENT_LOCKSTEP : 0;
  return 0;
}

namespace pm4 {
// geometry of the stateless kernels: one team per chain, up to 256 threads per block, no tree scratch
template <class M, typename real> inline Geometry small_geometry(ModelArgs<real>& A, int C) {
  Geometry g = plan_geometry<M, real>(A, C, 0, 0, small_threads<M>());
  g.grid = (C + g.teams - 1) / g.teams;
  return g;
}

template <typename real>
int mod_logp_grad(const pm4_mod_args_t* A, int C, const void* theta, void* lp, void* grad, cudaStream_t st) {
  using M = PM4_MODEL;
  ModelArgs<real> ma = make_model_args<real>(*A);
  if constexpr (M::OWNED) {  // the sampler's own arithmetic: owner-computes layout, one warp per chain
    if (!ma.oh.enabled) return (int)cudaErrorInvalidValue;
    const size_t smem = owned_pool_bytes(ma.oh, sizeof(real));
    if (smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
    PM4_CUDA_OK(cudaFuncSetAttribute(logp_grad_o_kernel<M, real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    logp_grad_o_kernel<M, real><<<(unsigned)((C + 7) / 8), 256, smem, st>>>(ma, C, (const real*)theta, (real*)lp, (real*)grad);
    return (int)cudaGetLastError();
  }
  const Geometry g = small_geometry<M, real>(ma, C);
  if (g.smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
  const unsigned threads = (unsigned)(g.teams * M::T * 32);
  if (ma.stage) {
    PM4_CUDA_OK(cudaFuncSetAttribute(logp_grad_kernel<M, real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    logp_grad_kernel<M, real, true><<<g.grid, threads, g.smem, st>>>(ma, C, g.team_reals, (const real*)theta, (real*)lp, (real*)grad);
  } else {
    PM4_CUDA_OK(cudaFuncSetAttribute(logp_grad_kernel<M, real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    logp_grad_kernel<M, real, false><<<g.grid, threads, g.smem, st>>>(ma, C, g.team_reals, (const real*)theta, (real*)lp, (real*)grad);
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
template <typename real>
int mod_loglik(const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, cudaStream_t st) {
  using M = PM4_MODEL;
  if (M::LL_ROW == 0 || n == 0) return 0;
  ModelArgs<real> ma = make_model_args<real>(*A);
  const int64_t blocks = (n + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  loglik_kernel<M, real><<<grid, 256, 0, st>>>(ma, n, (const real*)theta, (real*)out);
  return (int)cudaGetLastError();
}
template <typename real>
int mod_predict(const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, uint64_t seed, cudaStream_t st) {
  using M = PM4_MODEL;
  if (M::LL_ROW == 0 || n == 0) return 0;
  ModelArgs<real> ma = make_model_args<real>(*A);
  const int64_t blocks = (n + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  predict_kernel<M, real><<<grid, 256, 0, st>>>(ma, n, (const real*)theta, (real*)out, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (int)cudaGetLastError();
}
template <typename real>
int mod_prior(const pm4_mod_args_t* A, int64_t n, void* x_out, void* obs_out, uint64_t seed, cudaStream_t st) {
  using M = PM4_MODEL;
  if (n == 0) return 0;
  ModelArgs<real> ma = make_model_args<real>(*A);
  const int64_t blocks = (n + 7) / 8;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  prior_kernel<M, real><<<grid, 256, 0, st>>>(ma, n, (real*)x_out, (real*)obs_out, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (int)cudaGetLastError();
}
template <typename real> int mod_init(const pm4_mod_run_t* r, const void* theta0, double eps0, cudaStream_t st) {
  using M = PM4_MODEL;
  RunParams<real> p = make_run_params<real>(*r);
  if constexpr (M::OWNED) {
    if (!p.A.oh.enabled) return (int)cudaErrorInvalidValue;
    const size_t smem = owned_pool_bytes(p.A.oh, sizeof(real));
    if (smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
    PM4_CUDA_OK(cudaFuncSetAttribute(init_o_kernel<M, real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    init_o_kernel<M, real><<<(unsigned)((r->C + 7) / 8), 256, smem, st>>>(p, (const real*)theta0, (real)eps0);
    return (int)cudaGetLastError();
  }
  const Geometry g = small_geometry<M, real>(p.A, r->C);
  if (g.smem > SMEM_LIMIT) return (int)cudaErrorInvalidConfiguration;
  p.team_reals = g.team_reals;
  const unsigned threads = (unsigned)(g.teams * M::T * 32);
  if (p.A.stage) {
    PM4_CUDA_OK(cudaFuncSetAttribute(init_kernel<M, real, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    init_kernel<M, real, true><<<g.grid, threads, g.smem, st>>>(p, (const real*)theta0, (real)eps0);
  } else {
    PM4_CUDA_OK(cudaFuncSetAttribute(init_kernel<M, real, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    init_kernel<M, real, false><<<g.grid, threads, g.smem, st>>>(p, (const real*)theta0, (real)eps0);
  }
  return (int)cudaGetLastError();
}
// bytes of global tree scratch a NUTS run of C chains needs (0 when the scratch fits in shared memory)
template <typename real> int64_t mod_scratch_bytes(const pm4_mod_args_t* A, int C, int wpb, int max_depth) {
  using M = PM4_MODEL;
  ModelArgs<real> ma = make_model_args<real>(*A);
  if constexpr (use_warp_path<M>()) return warp_scratch_bytes<M, real>(ma, C, wpb, max_depth);
  const int n_slots = scratch_slots<M>(true, max_depth);
  const Geometry g = plan_geometry<M, real>(ma, C, wpb, n_slots, max_threads_for<M, real>());
  if (g.scr_smem) return 0;
  return (int64_t)g.grid * g.teams * (n_slots - g.n_fast) * M::DS * (int64_t)sizeof(real);
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
extern "C" int pm4m_loglik(int dtype, const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, void* stream) {
  PM4_BY_DTYPE(pm4::mod_loglik<float>(A, n, theta, out, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_loglik<double>(A, n, theta, out, (cudaStream_t)stream)));
}
extern "C" int pm4m_predict(int dtype, const pm4_mod_args_t* A, int64_t n, const void* theta, void* out, uint64_t seed,
                            void* stream) {
  PM4_BY_DTYPE(pm4::mod_predict<float>(A, n, theta, out, seed, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_predict<double>(A, n, theta, out, seed, (cudaStream_t)stream)));
}
extern "C" int pm4m_prior(int dtype, const pm4_mod_args_t* A, int64_t n, void* x_out, void* obs_out, uint64_t seed,
                          void* stream) {
  PM4_BY_DTYPE(pm4::mod_prior<float>(A, n, x_out, obs_out, seed, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_prior<double>(A, n, x_out, obs_out, seed, (cudaStream_t)stream)));
}
extern "C" int pm4m_init(int dtype, const pm4_mod_run_t* r, const void* theta0, double eps0, void* stream) {
  PM4_BY_DTYPE(pm4::mod_init<float>(r, theta0, eps0, (cudaStream_t)stream),
               PM4_F64_CALL(pm4::mod_init<double>(r, theta0, eps0, (cudaStream_t)stream)));
}
extern "C" int64_t pm4m_scratch_bytes(int dtype, const pm4_mod_args_t* A, int C, int wpb, int max_depth) {
  if (dtype == 0) return pm4::mod_scratch_bytes<float>(A, C, wpb, max_depth);
#if PM4_HAS_F64
  if (dtype == 1) return pm4::mod_scratch_bytes<double>(A, C, wpb, max_depth);
#endif
  return -22;
}
namespace pm4 {
template <typename real> int mod_nuts(const pm4_mod_run_t* r, int wpb, cudaStream_t st) {
  if constexpr (use_warp_path<PM4_MODEL>()) return dispatch_warp_nuts<PM4_MODEL, real>(*r, wpb, st);
  else return dispatch_run<PM4_MODEL, real, true>(*r, wpb, st);
}
}  // namespace pm4
extern "C" int pm4m_nuts(int dtype, const pm4_mod_run_t* r, int wpb, void* stream) {
  PM4_BY_DTYPE((pm4::mod_nuts<float>(r, wpb, (cudaStream_t)stream)),
               PM4_F64_CALL((pm4::mod_nuts<double>(r, wpb, (cudaStream_t)stream))));
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
extern "C" int pm4m_lpt_order(int C, const int32_t* work, int32_t* order, void* stream) {
  pm4::lpt_order_kernel<<<(unsigned)((C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(C, work, order);
  return (int)cudaGetLastError();
}
#endif  // PM4_MODEL
