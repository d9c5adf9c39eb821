// pm4_philox.cuh -- counter-based random numbers shared bit-for-bit with the CPU oracle
// (oracle/pm4_oracle.c).  Philox4x32-10 (Salmon et al., SC'11); counter = (chain, transition,
// slot, purpose), key = 64-bit seed.  The reference seeds TF's stateless RNG
// (/root/reference/pymc4/mcmc/samplers.py:186) and TFP consumes it positionally;
// the purposes below reproduce that structure with a stream the host can replay.
#pragma once
#include <stdint.h>

namespace pm4 {

enum : uint32_t { RNG_MOMENTUM = 0u, RNG_DIRECTION = 1u, RNG_LEAF = 2u, RNG_MERGE = 3u, RNG_HMC = 4u, RNG_PREDICTIVE = 5u };

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

template <typename real> __device__ __forceinline__ real u01(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ float u01<float>(uint32_t a, uint32_t) {
  return (float)(a >> 8) * (1.0f / 16777216.0f);
}
template <> __device__ __forceinline__ double u01<double>(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

template <typename real> __device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, real& n0, real& n1);
template <> __device__ __forceinline__ void box_muller<float>(uint32_t xa, uint32_t xb, float& n0, float& n1) {
  const float u1 = (float)((xa >> 8) + 1u) * (1.0f / 16777216.0f);
  const float u2 = (float)(xb >> 8) * (1.0f / 16777216.0f);
  const float rad = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586476925f * u2, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}
template <> __device__ __forceinline__ void box_muller<double>(uint32_t xa, uint32_t xb, double& n0, double& n1) {
  const double u1 = ((double)xa + 1.0) * (1.0 / 4294967296.0);
  const double u2 = (double)xb * (1.0 / 4294967296.0);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586476925 * u2, &s, &c);
  n0 = rad * c;
  n1 = rad * s;
}

// One draw of a variable from its distribution given the parameters (posterior predictive,
// /root/reference/pymc4/forward_sampling.py:230-427: the observed variables are re-sampled for every
// posterior draw).  Counter = (state lo, state hi, element of the row, purpose | attempt << 8), key =
// seed; the CPU oracle (oracle/pm4_oracle_impl.h: draw_value) runs the identical algorithm on the
// identical stream.  kinds: include/pm4b200.h PM4_T_*.
template <typename real>
__device__ __forceinline__ real draw_value(int kind, real p0, real p1, uint32_t c0, uint32_t c1, uint32_t c2,
                                           uint32_t k0, uint32_t k1) {
  const U4 u = philox4x32_10(c0, c1, c2, RNG_PREDICTIVE, k0, k1);
  if (kind == 1 || kind == 2) {  // Normal(loc, scale), HalfNormal(scale)
    real n0, n1;
    box_muller<real>(u.x, u.y, n0, n1);
    return kind == 1 ? p0 + p1 * n0 : fabs(p0 * n0);
  }
  if (kind == 3) return p0 * tan((real)1.57079632679489661923 * u01<real>(u.x, u.y));  // HalfCauchy(scale)
  if (kind == 4) return u01<real>(u.x, u.y) < p0 ? (real)1 : (real)0;                  // Bernoulli(probs)
  if (kind == 5) {                                                                       // Poisson(rate)
    const real lam = p0;
    if (lam < (real)10) {  // inversion by sequential search
      const real uu = u01<real>(u.x, u.y);
      real p = exp(-lam), F = p;
      int k = 0;
      while (uu > F && k < 1000) { k += 1; p *= lam / (real)k; F += p; }
      return (real)k;
    }
    // PTRS (Hoermann 1993), two uniforms per attempt
    const real slam = sqrt(lam), loglam = log(lam), b = (real)0.931 + (real)2.53 * slam;
    const real a = (real)-0.059 + (real)0.02483 * b, inv_alpha = (real)1.1239 + (real)1.1328 / (b - (real)3.4);
    const real vr = (real)0.9277 - (real)3.6224 / (b - (real)2);
    for (uint32_t it = 0; it < 64u; ++it) {
      const U4 w = it == 0 ? u : philox4x32_10(c0, c1, c2, RNG_PREDICTIVE | (it << 8), k0, k1);
      const real U = u01<real>(w.x, w.y) - (real)0.5, V = u01<real>(w.z, w.w);
      const real us = (real)0.5 - fabs(U);
      const real k = floor(((real)2 * a / us + b) * U + lam + (real)0.43);
      if (us >= (real)0.07 && V <= vr) return k;
      if (k < 0 || (us < (real)0.013 && V > us)) continue;
      if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + (real)1)) return k;
    }
    return floor(lam + (real)0.5);
  }
  return (real)NAN;  // potentials are not distributions
}

}  // namespace pm4
