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

}  // namespace pm4ls
