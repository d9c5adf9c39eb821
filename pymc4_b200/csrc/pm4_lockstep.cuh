// pm4_lockstep.cuh -- lock-step HMC for models with a dense GLM likelihood.
//
// The team kernels (pm4_kernels.cuh) give every chain its own warps and run chains asynchronously; a
// dense term such as Bernoulli(sigmoid(X beta)) with X [10^6, 100] cannot be evaluated per chain -- X
// must be read ONCE per leapfrog for all chains, as a GEMM (csrc/pm4_glm.cu).  So for those models the
// HMC transition is driven in lock-step over the chain axis, the way the reference runs it
// (tf.vectorized_map over chains inside tfp HamiltonianMonteCarlo, /root/reference/pymc4/mcmc/samplers.py:
// 265-315, 820-825): per leapfrog one batched gradient = the per-chain prior pass of the generated module
// + the tensor-core GLM pass, between small element-wise kernels (one warp per chain).  Arithmetic
// follows the oracle: half kick, L x (drift, gradient, full kick), half un-kick, MH.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm4_philox.cuh"

namespace pm4ls {

struct Params {
  int C, D, t, B, S, L, chain_offset;
  int adapt_enabled, adapt_pooled, adapt_kind, num_adapt;
  float target_accept, shrinkage, smoothing, decay, adapt_rate;
  uint32_t seed_lo, seed_hi;
  float* th;        // [C, D] current state
  float* g;         // [C, D] its gradient
  float* lp;        // [C]
  float* thp;       // [C, D] proposal
  float* gp;        // [C, D]
  float* lpp;       // [C]
  float* p;         // [C, D] momentum
  float* k0;        // [C] initial kinetic energy
  float* da;        // [C or 1, 4] eps, error_sum, log_avg, log_shrink_target
  float* accept;    // [C]
  const float* step_scale;  // [D] or NULL
  const float* momenta;     // [T, C, D] or NULL
  const float* uniforms;    // [T, C] or NULL
  float* states; float* lp_out; float* log_accept; uint8_t* accepted; float* traj;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float eps_of(const Params& a, int c) { return a.da[a.adapt_pooled ? 0 : (size_t)c * 4]; }
__device__ __forceinline__ float scale_of(const Params& a, int j) { return a.step_scale ? a.step_scale[j] : 1.0f; }

// dual-averaging state after the bootstrap gradient
__global__ void init_kernel(Params a, float eps0) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < (a.adapt_pooled ? 1 : a.C)) {
    float* d = a.da + (size_t)c * 4;
    d[0] = eps0; d[1] = 0; d[2] = 0; d[3] = logf(10.0f * eps0);
  }
}

// start of a transition: momentum, initial kinetic energy, proposal := current, half kick (one warp per chain)
__global__ void begin_kernel(Params a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.C) return;
  const float eps = eps_of(a, c);
  float kin = 0;
  for (int j = lane; j < a.D; j += 32) {
    float pv;
    if (a.momenta) pv = a.momenta[((size_t)a.t * a.C + c) * a.D + j];
    else {
      const int ro = j >> 5, lo = j & 31;
      const pm4::U4 u = pm4::philox4x32_10((uint32_t)(a.chain_offset + c), (uint32_t)a.t, (uint32_t)((ro >> 2) * 32 + lo),
                                           pm4::RNG_MOMENTUM, a.seed_lo, a.seed_hi);
      float n0, n1;
      if (ro & 2) pm4::box_muller<float>(u.z, u.w, n0, n1);
      else pm4::box_muller<float>(u.x, u.y, n0, n1);
      pv = (ro & 1) ? n1 : n0;
    }
    kin += pv * pv;
    const size_t k = (size_t)c * a.D + j;
    const float gv = a.g[k];
    a.thp[k] = a.th[k];
    a.gp[k] = gv;
    a.p[k] = pv + (0.5f * eps * scale_of(a, j)) * gv;
  }
  kin = wsum(kin);
  if (lane == 0) { a.k0[c] = 0.5f * kin; a.lpp[c] = a.lp[c]; }
}

__global__ void drift_kernel(Params a) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (size_t)a.C * a.D) return;
  const int c = (int)(k / a.D), j = (int)(k - (size_t)c * a.D);
  a.thp[k] = a.thp[k] + (eps_of(a, c) * scale_of(a, j)) * a.p[k];
}

__global__ void kick_kernel(Params a) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (size_t)a.C * a.D) return;
  const int c = (int)(k / a.D), j = (int)(k - (size_t)c * a.D);
  a.p[k] = a.p[k] + (eps_of(a, c) * scale_of(a, j)) * a.gp[k];
}

// end of a transition: half un-kick, Metropolis-Hastings, adaptation, trace (one warp per chain)
__global__ void finish_kernel(Params a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.C) return;
  const float eps = eps_of(a, c);
  float kin = 0;
  for (int j = lane; j < a.D; j += 32) {
    const size_t k = (size_t)c * a.D + j;
    const float pv = a.p[k] - (0.5f * eps * scale_of(a, j)) * a.gp[k];
    kin += pv * pv;
  }
  const float k1 = 0.5f * wsum(kin);
  const float lpc = a.lp[c], lpn = a.lpp[c];
  float lar = (lpn - lpc) + (a.k0[c] - k1);
  if (!(lar == lar) || isinf(lar)) lar = -INFINITY;
  float u;
  if (a.uniforms) u = a.uniforms[(size_t)a.t * a.C + c];
  else {
    const pm4::U4 uu = pm4::philox4x32_10((uint32_t)(a.chain_offset + c), (uint32_t)a.t, 0u, pm4::RNG_HMC, a.seed_lo, a.seed_hi);
    u = pm4::u01<float>(uu.x, uu.y);
  }
  const bool ok = logf(u) < lar;
  for (int j = lane; j < a.D; j += 32) {
    const size_t k = (size_t)c * a.D + j;
    const float tp = a.thp[k];
    if (a.traj) a.traj[((size_t)a.t * a.C + c) * a.D + j] = tp;
    if (ok) { a.th[k] = tp; a.g[k] = a.gp[k]; }
    if (a.t >= a.B && a.states) a.states[((size_t)(a.t - a.B) * a.C + c) * a.D + j] = ok ? tp : a.th[k];
  }
  if (lane == 0) {
    if (ok) a.lp[c] = lpn;
    const float accept = expf(lar < 0 ? lar : 0.0f);
    if (a.adapt_enabled) {
      if (a.adapt_pooled) a.accept[c] = accept;
      else {
        float* d = a.da + (size_t)c * 4;
        if (a.adapt_kind == 1) {
          if (a.t < a.num_adapt) d[0] = accept > a.target_accept ? d[0] * (1.0f + a.adapt_rate) : d[0] / (1.0f + a.adapt_rate);
        } else {
          const float step = (float)(a.t + 1);
          d[1] += a.target_accept - accept;
          const float log_eps = d[3] - d[1] * sqrtf(step) / ((a.smoothing + step) * a.shrinkage);
          const float eta = powf(step, -a.decay);
          d[2] = eta * log_eps + (1.0f - eta) * d[2];
          if (a.t < a.num_adapt) d[0] = expf(log_eps);
          else if (a.t == a.num_adapt) d[0] = expf(d[2]);
        }
      }
    }
    if (a.t >= a.B) {
      const size_t k = (size_t)(a.t - a.B) * a.C + c;
      if (a.lp_out) a.lp_out[k] = ok ? lpn : lpc;
      if (a.log_accept) a.log_accept[k] = lar;
      if (a.accepted) a.accepted[k] = ok ? 1 : 0;
    }
  }
}


// ------------------------------------------------------------------------------------------
// Lock-step NUTS (tfp NoUTurnSampler over the chain axis): every lock-step iteration
// is ONE batched gradient for all chains that are still growing their tree, between two calls of a
// resumable per-chain kernel (one warp per chain) that keeps the whole transition state -- both
// trajectory ends, candidate, subtree candidate, momentum sums and the U-turn checkpoints -- in global
// memory.  The per-chain arithmetic, the order of the tests and the Philox streams are those of the
// team kernel (pm4_kernels.cuh: nuts_kernel) and of the CPU oracle, so a run is replayable on the host.
// ------------------------------------------------------------------------------------------
enum { NS_CUR = 0, NS_OTH = 3, NS_CAND_TH = 6, NS_CAND_G = 7, NS_SUB_TH = 8, NS_SUB_G = 9, NS_RHO = 10, NS_RHO_SUB = 11,
       NS_MEM = 12 };  // then MEM_P[max_depth + 1], MEM_RHO[max_depth + 1]

struct NutsChain {  // 32 words per chain
  float H0, cur_lp, oth_lp, cand_lp, cand_energy, w, ediff_sum, w_sub, sub_lp, sub_energy, sub_ediff, eps;
  int cur_base, cur_is_right, nleap, depth, not_div, i, dir, cont_sub, sub_leaps, building;
  uint32_t ud;
  int pad[9];
};

struct NutsParams {
  Params h;  // C, D, t, B, S, chain_offset, adaptation, seeds, th/g/lp (current point), da, accept, step_scale, momenta, states, lp_out, log_accept
  int max_depth, ns;
  float max_energy_diff;
  float* slots;     // [C, ns, D]
  NutsChain* cs;    // [C]
  int* active;      // [1] chains that asked for another gradient
  float* thb;       // [C, D] positions handed to the batched gradient
  float* gb;        // [C, D] its result
  float* lpb;       // [C]
  int32_t* leapfrogs; uint8_t* diverging; float* energy; int32_t* depth_out; long long* total_leapfrogs;
};

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -INFINITY) return -INFINITY;
  return m + log1pf(expf(-fabsf(a - b)));
}

struct NutsView {
  const NutsParams& a;
  int c, lane;
  float* base;
  __device__ NutsView(const NutsParams& a_, int c_, int lane_) : a(a_), c(c_), lane(lane_) {
    base = a.slots + (size_t)c * a.ns * a.h.D;
  }
  __device__ __forceinline__ float* v(int slot) const { return base + (size_t)slot * a.h.D; }
  __device__ __forceinline__ float es(const NutsChain& s, int j) const { return (s.dir ? s.eps : -s.eps) * scale_of(a.h, j); }

  // a new doubling: pick the end to extend, clear the subtree accumulators
  __device__ void start_subtree(NutsChain& s) const {
    s.dir = (int)((s.ud >> s.depth) & 1u);
    if (s.dir != s.cur_is_right) {
      s.cur_base = NS_OTH - s.cur_base;
      const float sw = s.cur_lp; s.cur_lp = s.oth_lp; s.oth_lp = sw;
      s.cur_is_right = s.dir;
    }
    float* rs = v(NS_RHO_SUB);
    for (int j = lane; j < a.h.D; j += 32) rs[j] = 0.0f;
    s.w_sub = -INFINITY; s.sub_lp = s.cur_lp; s.sub_energy = s.cur_lp; s.sub_ediff = 0.0f;
    s.cont_sub = 1; s.sub_leaps = 0; s.i = 0;
  }

  // first half of a leapfrog: half kick, drift, hand the position to the batched gradient
  __device__ void pre_leapfrog(const NutsChain& s) const {
    float *th = v(s.cur_base), *p = v(s.cur_base + 1), *g = v(s.cur_base + 2);
    for (int j = lane; j < a.h.D; j += 32) {
      const float e = es(s, j);
      const float pv = p[j] + (0.5f * e) * g[j];
      const float tv = th[j] + e * pv;
      p[j] = pv; th[j] = tv;
      a.thb[(size_t)c * a.h.D + j] = tv;
    }
  }

  // second half (the gradient is in gb/lpb) + the leaf's bookkeeping; returns with s.cont_sub / s.i advanced
  __device__ void post_leapfrog(NutsChain& s) const {
    const int D = a.h.D, i = s.i;
    float *p = v(s.cur_base + 1), *g = v(s.cur_base + 2), *rsub = v(NS_RHO_SUB);
    const float* gn = a.gb + (size_t)c * D;
    s.cur_lp = a.lpb[c];
    s.sub_leaps += 1;
    const int pc = __popc((unsigned)i), t1 = __ffs(~(unsigned)i) - 1;
    float *memp = v(NS_MEM), *memr = v(NS_MEM + a.max_depth + 1);
    float q0 = 0, q1 = 0, q2 = 0;
    for (int j = lane; j < D; j += 32) {
      const float e = es(s, j), gv = gn[j];
      float pv = p[j] + e * gv;
      pv = pv - (0.5f * e) * gv;
      p[j] = pv; g[j] = gv;
      q0 += pv * pv;
      if (i & 1) {
        const float rs = rsub[j] + pv;
        rsub[j] = rs;
        const float df = rs - memr[(size_t)(pc - 1) * D + j];
        q1 += df * memp[(size_t)(pc - 1) * D + j];
        q2 += df * pv;
      } else {
        const float before = rsub[j];
        memp[(size_t)pc * D + j] = pv;
        memr[(size_t)pc * D + j] = before;
        rsub[j] = before + pv;
      }
    }
    q0 = wsum(q0);
    bool no_uturn = true;
    if (i & 1) {
      q1 = wsum(q1); q2 = wsum(q2);
      no_uturn = (q1 > 0) && (q2 > 0);
      for (int k = pc - 2; k >= pc - t1 && no_uturn; --k) {
        float u0 = 0, u1 = 0;
        for (int j = lane; j < D; j += 32) {
          const float df = rsub[j] - memr[(size_t)k * D + j];
          u0 += df * memp[(size_t)k * D + j];
          u1 += df * p[j];
        }
        u0 = wsum(u0); u1 = wsum(u1);
        no_uturn = (u0 > 0) && (u1 > 0);
      }
    }
    float energy = s.cur_lp - 0.5f * q0;
    if (energy != energy) energy = -INFINITY;
    const float dH = energy - s.H0;
    const bool not_divergent = (-dH < a.max_energy_diff);
    const float w_new = logaddexp_f(s.w_sub, dH);
    const float thresh = dH - w_new;
    const pm4::U4 ul = pm4::philox4x32_10((uint32_t)(a.h.chain_offset + c), (uint32_t)a.h.t,
                                          ((uint32_t)s.depth << 20) | (uint32_t)(i >> 1), pm4::RNG_LEAF, a.h.seed_lo, a.h.seed_hi);
    const float u = log1pf(-pm4::u01<float>((i & 1) ? ul.z : ul.x, (i & 1) ? ul.w : ul.y));
    if (u <= thresh) {
      const float* th = v(s.cur_base);
      float *sth = v(NS_SUB_TH), *sg = v(NS_SUB_G);
      for (int j = lane; j < D; j += 32) { sth[j] = th[j]; sg[j] = g[j]; }
      s.sub_lp = s.cur_lp;
      s.sub_energy = energy;
    }
    s.w_sub = w_new;
    s.not_div = s.not_div && not_divergent;
    if (not_divergent) s.sub_ediff += expf(dH < 0 ? dH : 0.0f);
    s.cont_sub = (no_uturn && not_divergent) ? 1 : 0;
    s.i = i + 1;
  }

  // the subtree is complete (or stopped): merge it into the trajectory; returns whether the tree goes on
  __device__ bool merge_subtree(NutsChain& s) const {
    const int D = a.h.D;
    s.ediff_sum += s.sub_ediff;
    s.nleap += s.sub_leaps;
    const float tree_w = s.cont_sub ? s.w_sub : -INFINITY;
    const float thresh = tree_w - s.w;
    const pm4::U4 um = pm4::philox4x32_10((uint32_t)(a.h.chain_offset + c), (uint32_t)a.h.t, (uint32_t)s.depth, pm4::RNG_MERGE,
                                          a.h.seed_lo, a.h.seed_hi);
    const float umerge = log1pf(-pm4::u01<float>(um.x, um.y));
    if ((umerge <= thresh) && s.cont_sub) {
      const float *sth = v(NS_SUB_TH), *sg = v(NS_SUB_G);
      float *cth = v(NS_CAND_TH), *cg = v(NS_CAND_G);
      for (int j = lane; j < D; j += 32) { cth[j] = sth[j]; cg[j] = sg[j]; }
      s.cand_lp = s.sub_lp;
      s.cand_energy = s.sub_energy;
    }
    s.w = logaddexp_f(tree_w, s.w);
    float *rho = v(NS_RHO);
    const float *rsub = v(NS_RHO_SUB), *p = v(s.cur_base + 1), *po = v(NS_OTH - s.cur_base + 1);
    float q0 = 0, q1 = 0;
    for (int j = lane; j < D; j += 32) {
      const float rr = rho[j] + rsub[j];
      rho[j] = rr;
      q0 += rr * p[j];
      q1 += rr * po[j];
    }
    q0 = wsum(q0); q1 = wsum(q1);
    const bool cont = s.cont_sub && (q0 > 0) && (q1 > 0);
    s.depth += 1;
    return cont && s.depth < a.max_depth;
  }
};

// start of a NUTS transition: momentum, both ends and the candidate at the current point, first half leapfrog
__global__ void nuts_begin_kernel(NutsParams a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.h.C) return;
  const int D = a.h.D;
  NutsView V(a, c, lane);
  NutsChain s;
  s.eps = eps_of(a.h, c);
  float *th = V.v(NS_CUR), *p = V.v(NS_CUR + 1), *g = V.v(NS_CUR + 2);
  float *oth = V.v(NS_OTH), *op = V.v(NS_OTH + 1), *og = V.v(NS_OTH + 2);
  float *cth = V.v(NS_CAND_TH), *cg = V.v(NS_CAND_G), *rho = V.v(NS_RHO);
  float kin = 0;
  for (int j = lane; j < D; j += 32) {
    float pv;
    if (a.h.momenta) pv = a.h.momenta[((size_t)a.h.t * a.h.C + c) * D + j];
    else {
      const int ro = j >> 5, lo = j & 31;
      const pm4::U4 u = pm4::philox4x32_10((uint32_t)(a.h.chain_offset + c), (uint32_t)a.h.t, (uint32_t)((ro >> 2) * 32 + lo),
                                           pm4::RNG_MOMENTUM, a.h.seed_lo, a.h.seed_hi);
      float n0, n1;
      if (ro & 2) pm4::box_muller<float>(u.z, u.w, n0, n1);
      else pm4::box_muller<float>(u.x, u.y, n0, n1);
      pv = (ro & 1) ? n1 : n0;
    }
    kin += pv * pv;
    const size_t k = (size_t)c * D + j;
    const float tv = a.h.th[k], gv = a.h.g[k];
    th[j] = tv; p[j] = pv; g[j] = gv;
    oth[j] = tv; op[j] = pv; og[j] = gv;
    cth[j] = tv; cg[j] = gv; rho[j] = pv;
  }
  kin = wsum(kin);
  const float clp = a.h.lp[c];
  s.H0 = clp - 0.5f * kin;
  s.cur_lp = clp; s.oth_lp = clp; s.cand_lp = clp; s.cand_energy = s.H0;
  s.w = 0.0f; s.ediff_sum = 0.0f;
  s.cur_base = NS_CUR; s.cur_is_right = 1; s.nleap = 0; s.depth = 0; s.not_div = 1; s.building = 1;
  s.ud = pm4::philox4x32_10((uint32_t)(a.h.chain_offset + c), (uint32_t)a.h.t, 0u, pm4::RNG_DIRECTION, a.h.seed_lo, a.h.seed_hi).x;
  __syncwarp();
  V.start_subtree(s);
  __syncwarp();
  V.pre_leapfrog(s);
  if (lane == 0) a.cs[c] = s;
}

// between two batched gradients: finish the leaf, close subtrees as they complete, start the next leapfrog
__global__ void nuts_advance_kernel(NutsParams a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.h.C) return;
  NutsChain s = a.cs[c];
  if (!s.building) return;
  NutsView V(a, c, lane);
  V.post_leapfrog(s);
  __syncwarp();
  if (!s.cont_sub || s.i >= (1 << s.depth)) {
    const bool go = V.merge_subtree(s);
    __syncwarp();
    if (go) V.start_subtree(s);
    else s.building = 0;
    __syncwarp();
  }
  if (s.building) {
    V.pre_leapfrog(s);
    if (lane == 0) atomicAdd(a.active, 1);
  }
  if (lane == 0) a.cs[c] = s;
}

// end of a transition: the candidate becomes the current point; adaptation and the trace
__global__ void nuts_finish_kernel(NutsParams a) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= a.h.C) return;
  const int D = a.h.D, t = a.h.t;
  const NutsChain s = a.cs[c];
  NutsView V(a, c, lane);
  const float *cth = V.v(NS_CAND_TH), *cg = V.v(NS_CAND_G);
  for (int j = lane; j < D; j += 32) {
    const size_t k = (size_t)c * D + j;
    const float tv = cth[j];
    a.h.th[k] = tv;
    a.h.g[k] = cg[j];
    if (t >= a.h.B && a.h.states) a.h.states[((size_t)(t - a.h.B) * a.h.C + c) * D + j] = tv;
  }
  if (lane != 0) return;
  a.h.lp[c] = s.cand_lp;
  const float accept = s.ediff_sum / (float)s.nleap;
  if (a.h.adapt_enabled) {
    if (a.h.adapt_pooled) a.h.accept[c] = accept;
    else {
      float* d = a.h.da + (size_t)c * 4;
      if (a.h.adapt_kind == 1) {
        if (t < a.h.num_adapt) d[0] = accept > a.h.target_accept ? d[0] * (1.0f + a.h.adapt_rate) : d[0] / (1.0f + a.h.adapt_rate);
      } else {
        const float step = (float)(t + 1);
        d[1] += a.h.target_accept - accept;
        const float log_eps = d[3] - d[1] * sqrtf(step) / ((a.h.smoothing + step) * a.h.shrinkage);
        const float eta = powf(step, -a.h.decay);
        d[2] = eta * log_eps + (1.0f - eta) * d[2];
        if (t < a.h.num_adapt) d[0] = expf(log_eps);
        else if (t == a.h.num_adapt) d[0] = expf(d[2]);
      }
    }
  }
  if (a.total_leapfrogs) a.total_leapfrogs[c] += s.nleap;
  if (t >= a.h.B) {
    const size_t k = (size_t)(t - a.h.B) * a.h.C + c;
    if (a.h.lp_out) a.h.lp_out[k] = s.cand_lp;
    if (a.leapfrogs) a.leapfrogs[k] = s.nleap;
    if (a.diverging) a.diverging[k] = s.not_div ? 0 : 1;
    if (a.energy) a.energy[k] = s.cand_energy;
    if (a.h.log_accept) a.h.log_accept[k] = logf(accept);
    if (a.depth_out) a.depth_out[k] = s.depth;
  }
}

}  // namespace pm4ls
