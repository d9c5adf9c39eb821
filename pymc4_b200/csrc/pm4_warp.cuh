// pm4_warp.cuh -- NUTS with ONE WARP per chain (M::T == 1), the path of small models (D <= 256:
// eight schools, radon-919).  Same algorithm, Philox streams and test order as nuts_kernel
// (pm4_kernels.cuh) and as the CPU oracle; what differs is where the chain's data lives and how the
// work is handed out:
//
//   * theta, momentum, gradient and both trajectory momentum sums stay in registers (R <= 8 per vector);
//     sums over dimensions are warp shuffles only -- no named barriers, no exchange through shared memory;
//   * shared memory holds what a leaf touches: the subtree candidate, the far end's momentum and the
//     first NL U-turn checkpoint levels (a subtree of 2^d leaves uses levels 0 .. d-1, so NL = 4 covers
//     every tree of up to 31 leapfrogs).  What a transition touches only per doubling -- far end position and
//     gradient, deeper checkpoint levels -- lives in a per-warp global scratch (L2); the transition's
//     candidate is the chain's own th/gr row in global memory;
//   * work items are (chain, block of `tb` transitions): the persistent warps take them from one atomic
//     queue, block-major, and a per-chain progress word orders the blocks of a chain.  4096 chains on
//     148 x 16 warps would otherwise run as two waves with the second one 73 % full.
//
// Replaces, like nuts_kernel, the body of tfp.mcmc.sample_chain behind _BaseSampler._run_chains
// (/root/reference/pymc4/mcmc/samplers.py:172-189); algorithm restated in SURVEY App. A.1-A.3.
// (included by pm4_kernels.cuh, after the team kernels and their launch glue)
#pragma once

namespace pm4 {

constexpr int WARP_MAX_R = 8;            // registers per lane and vector the warp path accepts
constexpr int WARP_FIXED_SLOTS = 3;      // shared-memory slots besides the checkpoint levels: SUB_TH, SUB_G, OTH_P
constexpr int WARP_GLOBAL_FIXED = 2;     // global slots besides the deep checkpoint levels: OTH_TH, OTH_G

template <class M> constexpr bool use_warp_path() { return M::T == 1 && M::R <= WARP_MAX_R; }

// all lanes receive the warp total (xor butterfly: bit-identical on every lane)
template <typename real> __device__ __forceinline__ real warp_allsum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
// three sums at once: packed transposing butterfly (6 shuffles) + 3 broadcasts
template <typename real> __device__ __forceinline__ void warp_allsum3(real& a, real& b, real& c3, int lane) {
  real v[4] = {a, b, c3, (real)0};
  warp_sum_packed<4, real>(v, lane);
  const real t = v[0];
  a = __shfl_sync(FULL, t, 0);
  b = __shfl_sync(FULL, t, 8);
  c3 = __shfl_sync(FULL, t, 16);
}
template <typename real> __device__ __forceinline__ void warp_allsum4(real (&q)[4], int lane) {
  warp_sum_packed<4, real>(q, lane);
  const real t = q[0];
  q[0] = __shfl_sync(FULL, t, 0);
  q[1] = __shfl_sync(FULL, t, 8);
  q[2] = __shfl_sync(FULL, t, 16);
  q[3] = __shfl_sync(FULL, t, 24);
}

__device__ __forceinline__ int ld_acquire(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------
// owner-computes layout (pymc4_b200/owned.py, Model::OWNED): the chain's state sits in RI registers per
// lane in the plan's internal order; element perm[reg * 32 + lane] of theta (or none).  The plan's pools
// (records as pairs of observations, slice tables) are staged once per CTA with a TMA bulk copy.
// ------------------------------------------------------------------------------------------
template <typename real, int RI> struct OCtx {
  int lane;
  int pj[RI];           // theta element held by register r of this lane, or -1
  const real* rec;      // shared memory: pair records
  const int32_t* oi;    // shared memory: owned int section (perm | own_len | slice bases + sentinel | foreign meta int4)
  int len_off, slb_off, fm_off, NF, Q, n_term;
  ModelArgs<real> A;
};

// bytes of the staged owned pools (+ the mbarrier); the host reads the header from its own copy
__host__ __device__ inline size_t owned_pool_bytes(const OwnedHeader& h, size_t real_size) {
  return 16 + (size_t)h.f_len * real_size + (size_t)h.i_len * sizeof(int32_t);
}

// Must be called by ALL threads of the CTA.  Returns the context; *scratch_base = first byte after the pools.
template <class M, typename real>
__device__ __forceinline__ OCtx<real, M::RI> make_octx(const ModelArgs<real>& A, unsigned char* smem_raw, size_t* pool_bytes_out) {
  OCtx<real, M::RI> c;
  c.lane = threadIdx.x & 31;
  c.A = A;
  const OwnedHeader h = A.oh;
  const uint32_t fbytes = (uint32_t)h.f_len * (uint32_t)sizeof(real), ibytes = (uint32_t)h.i_len * 4u;
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
    const unsigned char* gf = reinterpret_cast<const unsigned char*>(A.planf + h.f_off);
    const unsigned char* gi = reinterpret_cast<const unsigned char*>(A.plani + h.i_off);
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
  c.rec = reinterpret_cast<const real*>(sf);
  c.oi = reinterpret_cast<const int32_t*>(si);
  constexpr int KS1 = M::KS > 0 ? M::KS : 1;
  const int n_sl = M::KS + h.NF + 1;  // slice bases + sentinel
  c.len_off = M::RI * 32;
  c.slb_off = c.len_off + KS1 * 32;
  c.fm_off = c.slb_off + ((n_sl + 3) & ~3);
  c.NF = h.NF; c.Q = h.Q; c.n_term = h.n;
#pragma unroll
  for (int r = 0; r < M::RI; ++r) c.pj[r] = c.oi[r * 32 + c.lane];
  *pool_bytes_out = 16 + (size_t)fbytes + ibytes;
  return c;
}

// joint log-density + gradient in unconstrained space on the internal layout: shuffles only, no barrier
template <class M, typename real, class OC>
__device__ __forceinline__ real logp_grad_o(const real (&th)[M::RI], real (&g)[M::RI], const OC& c) {
  constexpr int RI = M::RI;
  real x[RI], dxdz[RI], gr[RI];
  M::template constrain_o<real>(th, x, dxdz, c.lane);
#pragma unroll
  for (int r = 0; r < RI; ++r) gr[r] = 0;
  const real lp = M::template terms_o<real>(x, gr, c);
#pragma unroll
  for (int r = 0; r < RI; ++r) g[r] = ((M::ID_REGS >> r) & 1u) ? gr[r] : gr[r] * dxdz[r];
  return lp;
}

// ------------------------------------------------------------------------------------------
// what the warp kernel needs from a layout: RV registers per vector, elem(r) = theta element of register r
// of this lane (or -1), the gradient, the momentum draw, this warp's shared-memory scratch
// ------------------------------------------------------------------------------------------
template <class M, typename real, bool STAGED> struct StripedWarp {  // theta order, xs / part hand-over (plan.py)
  static constexpr int RV = M::R;
  Ctx<real> c;
  Own<M, real> own;
  __device__ __forceinline__ void init(const RunParams<real>& a, unsigned char* smem_raw) {
    c = make_ctx<M, real, STAGED>(a.A, smem_raw, a.team_reals);
    own.load(c);
  }
  __device__ __forceinline__ int lane() const { return c.lane; }
  __device__ __forceinline__ real* scratch() const { return c.scr; }
  __device__ __forceinline__ int elem(int r) const { const int j = c.lane + 32 * r; return j < M::D ? j : -1; }
  __device__ __forceinline__ real grad(const real (&th)[RV], real (&g)[RV]) { return logp_grad<M, real>(th, g, c, own); }
  __device__ __forceinline__ void momentum(const RunParams<real>& a, int chain, int t, real (&p)[RV]) {
    draw_momentum<M, real>(a, chain, t, c.lane, p);
  }
};

template <class M, typename real> struct OwnedWarp {  // owner-computes layout (owned.py)
  static constexpr int RV = M::RI;
  OCtx<real, M::RI> c;
  real* scr;
  __device__ __forceinline__ void init(const RunParams<real>& a, unsigned char* smem_raw) {
    size_t pool = 0;
    c = make_octx<M, real>(a.A, smem_raw, &pool);
    scr = reinterpret_cast<real*>(smem_raw + pool) + (size_t)(threadIdx.x >> 5) * (size_t)a.team_reals;
  }
  __device__ __forceinline__ int lane() const { return c.lane; }
  __device__ __forceinline__ real* scratch() const { return scr; }
  __device__ __forceinline__ int elem(int r) const { return c.pj[r]; }
  __device__ __forceinline__ real grad(const real (&th)[RV], real (&g)[RV]) { return logp_grad_o<M, real>(th, g, c); }
  // the Philox stream is keyed by theta element in blocks of four: draw in theta order (two blocks per lane
  // for D <= 256), then permute through this warp's scratch (free at the start of a transition)
  __device__ __forceinline__ void momentum(const RunParams<real>& a, int chain, int t, real (&p)[RV]) {
    if (a.momenta) {
#pragma unroll
      for (int r = 0; r < RV; ++r) p[r] = c.pj[r] >= 0 ? a.momenta[((size_t)t * a.C + chain) * M::D + c.pj[r]] : (real)0;
      return;
    }
    real ps[M::R];
    draw_momentum<M, real>(a, chain, t, c.lane, ps);
#pragma unroll
    for (int r = 0; r < M::R; ++r) scr[c.lane + 32 * r] = ps[r];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < RV; ++r) p[r] = c.pj[r] >= 0 ? scr[c.pj[r]] : (real)0;
    __syncwarp();
  }
};

// SCALED: a per-dimension step scale is attached (step_size given per element / diagonal mass matrix); without
// it the signed step size is one uniform scalar and costs no registers
// ROUNDS: lock-step instantiation (a.rounds transitions, grid barrier + pooled update after each)
template <class M, typename real, int MAXT, bool STAGED, bool SCALED, bool ROUNDS>
__global__ void __launch_bounds__(MAXT, 1) nuts_warp_kernel(RunParams<real> a) {
  static_assert(M::T == 1, "one warp per chain");
  using L = typename std::conditional<M::OWNED, OwnedWarp<M, real>, StripedWarp<M, real, STAGED>>::type;
  constexpr int R = L::RV, D = M::D, DP = M::DP, DS = 32 * R;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  L lay;
  lay.init(a, smem_raw);
  const int lane = lay.lane();
  const real NEG_INF = -m_inf<real>();
  const size_t C = (size_t)a.C;
  const int NL = a.n_fast;  // checkpoint levels on chip
  // shared-memory slots of this warp (lane-private columns: no synchronisation needed)
  real* const s_sub_th = lay.scratch() + lane;
  real* const s_sub_g = s_sub_th + DS;
  real* const s_oth_p = s_sub_g + DS;
  real* const s_mem = s_oth_p + DS;  // level l: momentum at 2 l DS, rho at (2 l + 1) DS
  // global scratch of this warp
  const int n_gslots = WARP_GLOBAL_FIXED + 2 * (a.max_depth + 1 - NL > 0 ? a.max_depth + 1 - NL : 0);
  real* const g_oth_th = a.scratch + ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (size_t)n_gslots * DS + lane;
  real* const g_oth_g = g_oth_th + DS;
  real* const g_mem = g_oth_g + DS;
  // checkpoint level l is in shared memory when l < NL, else in the global scratch: `f` is instantiated
  // once per address space, so the accesses compile to LDS/STS and LDG/STG, never to generic ones
  auto with_level = [&](int l, auto&& f) {
    if (l < NL) f(s_mem + 2 * l * DS);
    else f(g_mem + 2 * (l - NL) * DS);
  };
  // theta element of every register of this lane, and its step scale
  int ej[R];
  real scl[SCALED ? R : 1];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ej[r] = lay.elem(r);
    if constexpr (SCALED) scl[r] = ej[r] >= 0 ? a.step_scale[ej[r]] : (real)1;
  }
  // lock-step rounds (pooled step size still adapting): one transition per round, every chain, then a grid
  // barrier whose last arrival runs the pooled dual-averaging update -- all in this one (cooperative) launch
  const int rounds = ROUNDS ? a.rounds : 1;
  __shared__ double s_part[ROUNDS ? 32 : 1];
  __shared__ int s_last;
  for (int rd = 0; rd < rounds; ++rd) {
  const int rt_begin = ROUNDS ? a.t_begin + rd : a.t_begin;
  const int rt_count = ROUNDS ? 1 : a.t_count;
  int32_t* const queue = ROUNDS ? a.queue + rd : a.queue;
  const int tb = (a.tb > 0 && !ROUNDS) ? a.tb : rt_count;
  const int n_blocks = (rt_count + tb - 1) / tb;
  const int n_items = n_blocks * a.C;

  for (;;) {
    int qi = 0;
    if (lane == 0) qi = atomicAdd(queue, 1);
    qi = __shfl_sync(FULL, qi, 0);
    if (qi >= n_items) break;
    const int blk = qi / a.C, slot = qi - blk * a.C;
    const int chain = a.order ? a.order[slot] : slot;
    const int t0 = rt_begin + blk * tb;
    const int t1 = t0 + tb < rt_begin + rt_count ? t0 + tb : rt_begin + rt_count;
    if (n_blocks > 1 && blk > 0) {  // the previous block of this chain is running on another warp (or done)
      if (lane == 0) {
        while (ld_acquire(a.progress + chain) < blk) __nanosleep(100);
      }
      __syncwarp();
    }
    const uint32_t rng_chain = (uint32_t)(a.chain_offset + chain);
    real* const cth = a.th + (size_t)chain * DP;  // the chain's current point = the transition's candidate (theta order)
    real* const cg = a.gr + (size_t)chain * DP;
    real clp = __ldcg(a.lp + chain);
    DualAvg<real> da;
    {
      const real* dp = a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4);
      da.eps = __ldcg(dp); da.error_sum = __ldcg(dp + 1); da.log_avg = __ldcg(dp + 2); da.log_shrink_target = __ldcg(dp + 3);
    }
    int launch_leaps = 0;

    for (int t = t0; t < t1; ++t) {
      real th[R], p[R], g[R], rho[R], rho_sub[R];
      const real eps = da.eps;
      lay.momentum(a, chain, t, p);
      if constexpr (SCALED) {  // scale 0 = frozen dimension (compound step): no momentum
#pragma unroll
        for (int r = 0; r < R; ++r) p[r] = scl[r] == (real)0 ? (real)0 : p[r];
      }
      real q0 = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        th[r] = ej[r] >= 0 ? __ldcg(cth + ej[r]) : (real)0;
        g[r] = ej[r] >= 0 ? __ldcg(cg + ej[r]) : (real)0;
        rho[r] = p[r];
        q0 += p[r] * p[r];
        g_oth_th[32 * r] = th[r];
        s_oth_p[32 * r] = p[r];
        g_oth_g[32 * r] = g[r];
      }
      q0 = warp_allsum(q0);
      const real H0 = clp - (real)0.5 * q0;

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
            const real vt = g_oth_th[32 * r], vp = s_oth_p[32 * r], vg = g_oth_g[32 * r];
            g_oth_th[32 * r] = th[r]; s_oth_p[32 * r] = p[r]; g_oth_g[32 * r] = g[r];
            th[r] = vt; p[r] = vp; g[r] = vg;
          }
          const real sw = cur_lp; cur_lp = oth_lp; oth_lp = sw;
          cur_is_right = dir;
        }
        const int nsteps = 1 << depth;
#pragma unroll
        for (int r = 0; r < R; ++r) rho_sub[r] = 0;
        const real eps_s = dir ? eps : -eps;
        const real heps_s = (real)0.5 * eps_s;
        auto es = [&](int r) -> real { if constexpr (SCALED) return eps_s * scl[r]; else return eps_s; };
        auto hes = [&](int r) -> real { if constexpr (SCALED) return (real)0.5 * (eps_s * scl[r]); else return heps_s; };
        real w_sub = NEG_INF, sub_lp = cur_lp, sub_energy = cur_lp, sub_ediff = 0;
        bool cont_sub = true;
        int sub_leaps = 0;
        U4 ul{0u, 0u, 0u, 0u};

        for (int i = 0; i < nsteps && cont_sub; ++i) {
          // leaf uniforms: two leaves share a Philox block, lane l holds pair-block (base + l)
          if ((i & 63) == 0)
            ul = philox4x32_10(rng_chain, (uint32_t)t, ((uint32_t)depth << 20) | (uint32_t)((i >> 1) + lane),
                               RNG_LEAF, a.seed_lo, a.seed_hi);
          // one leapfrog step: tfp SimpleLeapfrogIntegrator with num_steps = 1
#pragma unroll
          for (int r = 0; r < R; ++r) {
            p[r] = p[r] + hes(r) * g[r];
            th[r] = th[r] + es(r) * p[r];
          }
          cur_lp = lay.grad(th, g);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            p[r] = p[r] + es(r) * g[r];
            p[r] = p[r] - hes(r) * g[r];
          }
          sub_leaps += 1;

          // one reduction per leaf: kinetic energy and, on odd leaves, the U-turn test against the
          // checkpoint of the preceding even leaf
          const int pc = __popc((unsigned)i), tz = __ffs(~(unsigned)i) - 1;
          real k2 = 0, u1 = 0, u2 = 0;
          bool no_uturn = true;
          if (i & 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              rho_sub[r] += p[r];
              k2 += p[r] * p[r];
            }
            with_level(pc - 1, [&](const real* mp) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                const real df = rho_sub[r] - mp[DS + 32 * r];
                u1 += df * mp[32 * r];
                u2 += df * p[r];
              }
            });
            warp_allsum3(k2, u1, u2, lane);
            no_uturn = (u1 > 0) && (u2 > 0);
            // remaining checkpoints [popcount(i) - trailing_ones(i), popcount(i) - 1), two per reduction
            for (int s = pc - 2; s >= pc - tz && no_uturn; s -= 2) {
              const bool two = s - 1 >= pc - tz;
              real u4[4];
              u4[0] = 0; u4[1] = 0; u4[2] = 0; u4[3] = 0;
              with_level(s, [&](const real* mp) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                  const real df = rho_sub[r] - mp[DS + 32 * r];
                  u4[0] += df * mp[32 * r];
                  u4[1] += df * p[r];
                }
              });
              if (two) {
                with_level(s - 1, [&](const real* mp) {
#pragma unroll
                  for (int r = 0; r < R; ++r) {
                    const real df = rho_sub[r] - mp[DS + 32 * r];
                    u4[2] += df * mp[32 * r];
                    u4[3] += df * p[r];
                  }
                });
              }
              warp_allsum4(u4, lane);
              no_uturn = (u4[0] > 0) && (u4[1] > 0) && (!two || ((u4[2] > 0) && (u4[3] > 0)));
            }
          } else {
            // even leaves are checkpointed into level popcount(i): (p_i, rho before this leaf)
            with_level(pc, [&](real* mp) {
#pragma unroll
              for (int r = 0; r < R; ++r) {
                mp[32 * r] = p[r];
                mp[DS + 32 * r] = rho_sub[r];
              }
            });
#pragma unroll
            for (int r = 0; r < R; ++r) {
              rho_sub[r] += p[r];
              k2 += p[r] * p[r];
            }
            k2 = warp_allsum(k2);
          }

          real energy = cur_lp - (real)0.5 * k2;
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
              s_sub_th[32 * r] = th[r];
              s_sub_g[32 * r] = g[r];
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
            if (ej[r] >= 0) {
              cth[ej[r]] = s_sub_th[32 * r];
              cg[ej[r]] = s_sub_g[32 * r];
            }
          }
          cand_lp = sub_lp;
          cand_energy = sub_energy;
        }
        w = m_logaddexp(tree_w, w);
        real d1 = 0, d2 = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          rho[r] += rho_sub[r];
          d1 += rho[r] * p[r];
          d2 += rho[r] * s_oth_p[32 * r];
        }
        real d3 = 0;
        warp_allsum3(d1, d2, d3, lane);
        cont = cont_sub && (d1 > 0) && (d2 > 0);
        depth += 1;
      }

      // the transition's result is the candidate
      clp = cand_lp;
      launch_leaps += nleap;
      const real accept = ediff_sum / (real)nleap;
      if (a.adapt_enabled) {
        if (a.adapt_pooled) { if (lane == 0) a.accept[chain] = accept; }
        else da.update(a, t, accept);
      }
      if (t >= a.B) {
        const size_t k = (size_t)(t - a.B) * C + chain;
        if (a.states) {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (ej[r] >= 0) __stcs(a.states + k * D + ej[r], __ldcg(cth + ej[r]));
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

    // persist the chain for the next block / launch (th and gr rows are already current)
    if (lane == 0) {
      a.lp[chain] = clp;
      if (!a.adapt_pooled) da.store(a.da + (size_t)chain * 4);
      if (a.total_leapfrogs) a.total_leapfrogs[chain] = __ldcg(a.total_leapfrogs + chain) + launch_leaps;
      if (a.launch_leaps) a.launch_leaps[chain] = (blk == 0 ? 0 : __ldcg(a.launch_leaps + chain)) + launch_leaps;
    }
    if (n_blocks > 1) {  // every lane's stores are fenced before lane 0 publishes the block
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release(a.progress + chain, blk + 1);
    }
  }
  if constexpr (ROUNDS) {
    __threadfence();  // this CTA's accept words and chain rows are visible before it arrives
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(a.sync, 1) == (int)gridDim.x * (rd + 1) - 1;
    __syncthreads();
    if (s_last) {
      __threadfence();
      if (a.adapt_enabled && a.adapt_pooled) pooled_update<real>(a, rt_begin, s_part);
      __threadfence();
      if (threadIdx.x == 0) st_release(a.sync + 1, rd + 1);
    } else if (threadIdx.x == 0) {
      while (ld_acquire(a.sync + 1) < rd + 1) __nanosleep(64);
    }
    __syncthreads();
  }
  }
}

// parity kernel (a) and bootstrap on the owner-computes layout: one warp per chain
template <class M, typename real>
__global__ void __launch_bounds__(256) logp_grad_o_kernel(ModelArgs<real> A, int C, const real* __restrict__ theta,
                                                         real* __restrict__ lp_out, real* __restrict__ grad_out) {
  constexpr int RI = M::RI, D = M::D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  size_t pool = 0;
  const OCtx<real, RI> c = make_octx<M, real>(A, smem_raw, &pool);
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= C) return;
  real th[RI], g[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) th[r] = c.pj[r] >= 0 ? theta[(size_t)chain * D + c.pj[r]] : (real)0;
  const real lp = logp_grad_o<M, real>(th, g, c);
  if (c.lane == 0) lp_out[chain] = lp;
  if (grad_out) {
#pragma unroll
    for (int r = 0; r < RI; ++r)
      if (c.pj[r] >= 0) grad_out[(size_t)chain * D + c.pj[r]] = g[r];
  }
}

template <class M, typename real>
__global__ void __launch_bounds__(256) init_o_kernel(RunParams<real> a, const real* __restrict__ theta0, real eps0) {
  constexpr int RI = M::RI, D = M::D, DP = M::DP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  size_t pool = 0;
  const OCtx<real, RI> c = make_octx<M, real>(a.A, smem_raw, &pool);
  const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (chain >= a.C) return;
  real th[RI], g[RI];
#pragma unroll
  for (int r = 0; r < RI; ++r) th[r] = c.pj[r] >= 0 ? theta0[(size_t)chain * D + c.pj[r]] : (real)0;
  const real lp = logp_grad_o<M, real>(th, g, c);
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    if (c.pj[r] >= 0) {
      a.th[(size_t)chain * DP + c.pj[r]] = th[r];
      a.gr[(size_t)chain * DP + c.pj[r]] = g[r];
    }
  }
  if (c.lane == 0) {
    a.lp[chain] = lp;
    if (a.total_leapfrogs) a.total_leapfrogs[chain] = 0;
    if (!a.adapt_pooled || chain == 0) {
      real* d = a.da + (a.adapt_pooled ? 0 : (size_t)chain * 4);
      d[0] = eps0; d[1] = 0; d[2] = 0; d[3] = m_log((real)10 * eps0);
    }
  }
}

// geometry of the warp path: warps per block and checkpoint levels on chip
struct WarpGeometry {
  int warps, grid, team_reals, n_lvl, n_gslots, stage;
  size_t smem;
};

template <class M, typename real> inline size_t warp_pool_bytes(const ModelArgs<real>& A) {
  if constexpr (M::OWNED) return owned_pool_bytes(A.oh, sizeof(real));
  else return pool_bytes(A);
}

template <class M, typename real>
inline WarpGeometry plan_warp_geometry(ModelArgs<real>& A, int C, int wpb, int max_depth, int max_threads) {
  constexpr int DSV = M::OWNED ? 32 * M::RI : M::DS;
  WarpGeometry g{};
  const int sms = sm_count();
  int warps = max_threads / 32;
  if (wpb > 0 && wpb < warps) warps = wpb;
  const int per_sm = (C + sms - 1) / sms;
  if (wpb <= 0 && warps > per_sm) warps = per_sm;
  if (warps < 1) warps = 1;
  const size_t base = M::OWNED ? 0 : team_base_reals<M, real>(A);
  if (!M::OWNED && A.stage && pool_bytes(A) > SMEM_LIMIT / 2) A.stage = 0;
  const int want = max_depth + 1;
  int lvl = 0;
  for (;;) {
    const size_t fixed = warp_pool_bytes<M, real>(A) + (size_t)warps * (base + (size_t)WARP_FIXED_SLOTS * DSV) * sizeof(real);
    const size_t per_lvl = (size_t)warps * 2 * DSV * sizeof(real);
    lvl = fixed < SMEM_LIMIT ? (int)((SMEM_LIMIT - fixed) / per_lvl) : -1;
    if (lvl > want) lvl = want;
    // four levels serve every tree of up to 31 leapfrogs: give up warps before going below that
    if (lvl >= (want < 4 ? want : 4) || warps == 1) break;
    --warps;
  }
  if (lvl < 0) {
    if (!M::OWNED) A.stage = 0;
    lvl = 0;
  }
  g.warps = warps;
  g.n_lvl = lvl;
  g.n_gslots = WARP_GLOBAL_FIXED + 2 * (want - lvl);
  g.team_reals = (int)(base + (size_t)(WARP_FIXED_SLOTS + 2 * lvl) * DSV);
  g.stage = A.stage;
  g.smem = warp_pool_bytes<M, real>(A) + (size_t)warps * g.team_reals * sizeof(real);
  int grid = (C + warps - 1) / warps;
  if (grid > sms) grid = sms;
  g.grid = grid;
  return g;
}

template <class M, typename real> inline int dispatch_warp_nuts(const pm4_mod_run_t& r, int wpb, cudaStream_t st) {
  RunParams<real> p = make_run_params<real>(r);
  const WarpGeometry g = plan_warp_geometry<M, real>(p.A, r.C, wpb, r.max_depth, max_threads_for<M, real>());
  if (g.smem > SMEM_LIMIT || !p.scratch || !p.progress) return (int)cudaErrorInvalidConfiguration;
  p.team_reals = g.team_reals;
  p.n_fast = g.n_lvl;
  p.n_slots = g.n_gslots;
  PM4_CUDA_OK(cudaMemsetAsync(p.queue, 0, sizeof(int32_t) * (size_t)(p.rounds > 0 ? p.rounds : 1), st));
  if (p.rounds > 0) {
    if (!p.sync) return (int)cudaErrorInvalidValue;
    PM4_CUDA_OK(cudaMemsetAsync(p.sync, 0, 2 * sizeof(int32_t), st));
  }
  if (p.rounds == 0 && p.tb > 0 && p.tb < p.t_count) PM4_CUDA_OK(cudaMemsetAsync(p.progress, 0, sizeof(int32_t) * (size_t)r.C, st));
#define PM4_WLAUNCH(MAXT, STG)                                                                                   \
  do {                                                                                                           \
    auto k = p.rounds > 0 ? (p.step_scale ? nuts_warp_kernel<M, real, MAXT, STG, true, true>                      \
                                          : nuts_warp_kernel<M, real, MAXT, STG, false, true>)                    \
                          : (p.step_scale ? nuts_warp_kernel<M, real, MAXT, STG, true, false>                     \
                                          : nuts_warp_kernel<M, real, MAXT, STG, false, false>);                  \
    PM4_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));             \
    if (p.rounds > 0) { /* the blocks wait on one another: co-residency must be guaranteed */                    \
      void* args[] = {(void*)&p};                                                                                \
      return (int)cudaLaunchCooperativeKernel((const void*)k, dim3(g.grid), dim3(g.warps * 32), args, g.smem, st); \
    }                                                                                                            \
    k<<<g.grid, g.warps * 32, g.smem, st>>>(p);                                                                  \
    return (int)cudaGetLastError();                                                                              \
  } while (0)
  if (g.warps * 32 <= small_threads<M>()) {
    if (M::OWNED || g.stage) PM4_WLAUNCH(small_threads<M>(), true);
    if constexpr (!M::OWNED) PM4_WLAUNCH(small_threads<M>(), false);
  }
  if constexpr (sizeof(real) == 4) {
    if (M::OWNED || g.stage) PM4_WLAUNCH(big_threads<M>(), true);
    if constexpr (!M::OWNED) PM4_WLAUNCH(big_threads<M>(), false);
  }
#undef PM4_WLAUNCH
  return (int)cudaErrorInvalidConfiguration;
}

template <class M, typename real> inline int64_t warp_scratch_bytes(ModelArgs<real>& A, int C, int wpb, int max_depth) {
  const WarpGeometry g = plan_warp_geometry<M, real>(A, C, wpb, max_depth, max_threads_for<M, real>());
  return (int64_t)g.grid * g.warps * g.n_gslots * (M::OWNED ? 32 * M::RI : M::DS) * (int64_t)sizeof(real);
}

}  // namespace pm4
