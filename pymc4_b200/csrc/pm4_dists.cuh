// pm4_dists.cuh -- hand-written log-densities with analytic partial derivatives.
//
// These are the device counterparts of the TFP log_prob closed forms the reference delegates to
// (/root/reference/pymc4/distributions/continuous.py:108-112 Normal, :234-238 HalfNormal,
//  :659-663 HalfCauchy(loc=0); discrete.py:75-78 Bernoulli(probs=), :500-503 Poisson(rate=);
//  SURVEY App. B) and of the element-wise glue ops models use.  Everything is __forceinline__:
//  the code generator composes them into one straight-line body per term, so unused partials are
//  dead-code-eliminated by nvcc.
#pragma once
#include <math.h>

namespace pm4 {

template <typename real> struct K;
template <> struct K<float> {
  static constexpr float HALF_LOG_2PI = 0.91893853320467274178f;
  static constexpr float HALF_LOG_HALF_PI = 0.22579135264472743236f;
  static constexpr float LOG_2_OVER_PI = -0.45158270528945486473f;
};
template <> struct K<double> {
  static constexpr double HALF_LOG_2PI = 0.91893853320467274178;
  static constexpr double HALF_LOG_HALF_PI = 0.22579135264472743236;
  static constexpr double LOG_2_OVER_PI = -0.45158270528945486473;
};

// ---- math shims (float / double overload sets) ---------------------------------------------
// FP32 path: bare MUFU exp2 / log2 / reciprocal (ex2.approx, lg2.approx, rcp.approx with flush-to-zero;
// relative error ~1e-6, two orders below the 1e-5 parity bound) without the denormal-range fix-ups of
// __expf / __logf / __fdividef: arguments here are scales, densities and acceptance ratios, never
// denormal.  The FP64 path keeps the correctly rounded library functions.
__device__ __forceinline__ float mufu_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float mufu_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float mufu_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float m_exp(float v) { return mufu_ex2(v * 1.4426950408889634f); }
__device__ __forceinline__ double m_exp(double v) { return exp(v); }
__device__ __forceinline__ float m_log(float v) { return mufu_lg2(v) * 0.6931471805599453f; }
__device__ __forceinline__ double m_log(double v) { return log(v); }
__device__ __forceinline__ float m_log1p(float v) {
  // lg2.approx(1 + v) has an ABSOLUTE error of ~1e-7, i.e. a relative error of 1e-7 / |log1p v|: below 0.02 the
  // alternating series takes over (four terms: truncation < v^4 / 5 = 3e-8 relative), above it the MUFU form is
  // within 5e-6 relative -- both inside the 1e-5 parity bound, and no slower than the bare MUFU form
  const float s = v * (1.0f + v * (-0.5f + v * (0.33333334f - 0.25f * v)));
  return fabsf(v) < 0.02f ? s : m_log(1.0f + v);
}
__device__ __forceinline__ double m_log1p(double v) { return log1p(v); }
__device__ __forceinline__ float m_sqrt(float v) { return sqrtf(v); }
__device__ __forceinline__ double m_sqrt(double v) { return sqrt(v); }
__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }
__device__ __forceinline__ float m_tanh(float v) { return tanhf(v); }
__device__ __forceinline__ double m_tanh(double v) { return tanh(v); }
__device__ __forceinline__ float m_abs(float v) { return fabsf(v); }
__device__ __forceinline__ double m_abs(double v) { return fabs(v); }
__device__ __forceinline__ float m_lgamma(float v) { return lgammaf(v); }
__device__ __forceinline__ double m_lgamma(double v) { return lgamma(v); }
__device__ __forceinline__ float m_rcp(float v) { return mufu_rcp(v); }
__device__ __forceinline__ double m_rcp(double v) { return 1.0 / v; }
__device__ __forceinline__ bool m_isnan(float v) { return isnan(v); }
__device__ __forceinline__ bool m_isnan(double v) { return isnan(v); }
template <typename real> __device__ __forceinline__ real m_inf() { return (real)INFINITY; }

template <typename real> __device__ __forceinline__ real m_sigmoid(real v) { return m_rcp((real)1 + m_exp(-v)); }
template <typename real> __device__ __forceinline__ real m_softplus(real v) {
  return v > 0 ? v + m_log1p(m_exp(-v)) : m_log1p(m_exp(v));
}
template <typename real> __device__ __forceinline__ real m_digamma(real x) {
  // d/dx lgamma on user expressions: only positive finite arguments are meaningful here; anything else is a bad
  // proposal and gets NaN (a divergent transition) instead of a recurrence that never terminates
  if (!(x > 0) || x == m_inf<real>()) return (real)NAN;
  real r = 0;
#pragma unroll 1
  for (int k = 0; k < 6 && x < 6; ++k) { r -= m_rcp(x); x += 1; }
  const real f = m_rcp(x * x);
  return r + m_log(x) - (real)0.5 / x -
         f * ((real)(1.0 / 12) - f * ((real)(1.0 / 120) - f * ((real)(1.0 / 252) - f * (real)(1.0 / 240))));
}
template <typename real> __device__ __forceinline__ real m_logaddexp(real a, real b) {
  // branch-free: exp(-inf) = 0 covers an operand at -inf, d = 0 covers both at -inf (where a - b would be NaN)
  const real m = a > b ? a : b;
  const real d = (a == b) ? (real)0 : -m_abs(a - b);
  return m + m_log1p(m_exp(d));
}

// ---- log-densities: lp plus d/dvalue and d/dparams --------------------------------------------
// Normal(loc, scale).  `inv_s` = 1/scale and `log_s` = log(scale) are passed in so that a
// chain-uniform scale is inverted once per evaluation, not once per observation.
template <typename real>
__device__ __forceinline__ real normal_lpdf(real x, real mu, real inv_s, real log_s, real& dx, real& dmu, real& ds) {
  const real w = (x - mu) * inv_s;
  dmu = w * inv_s;
  dx = -dmu;
  ds = (w * w - (real)1) * inv_s;
  return (real)-0.5 * w * w - K<real>::HALF_LOG_2PI - log_s;
}

template <typename real>
__device__ __forceinline__ real halfnormal_lpdf(real x, real inv_s, real log_s, real& dx, real& ds) {
  const real w = x * inv_s;
  dx = -w * inv_s;
  ds = (w * w - (real)1) * inv_s;
  const real lp = (real)-0.5 * w * w - K<real>::HALF_LOG_HALF_PI - log_s;
  return x < 0 ? -m_inf<real>() : lp;
}

template <typename real>
__device__ __forceinline__ real halfcauchy_lpdf(real x, real inv_s, real log_s, real& dx, real& ds) {
  const real w = x * inv_s;
  const real w2 = w * w;
  const real iden = m_rcp((real)1 + w2);
  dx = (real)-2 * w * inv_s * iden;
  ds = ((real)2 * w2 * iden - (real)1) * inv_s;
  const real lp = K<real>::LOG_2_OVER_PI - log_s - m_log1p(w2);
  return x < 0 ? -m_inf<real>() : lp;
}

// Bernoulli(probs = p): xlogy(x, p) + xlog1py(1 - x, -p)
template <typename real>
__device__ __forceinline__ real bernoulli_lpdf(real x, real p, real& dx, real& dp) {
  const real lg = m_log(p), l1 = m_log1p(-p);
  const real omx = (real)1 - x;
  dp = (x == 0 ? (real)0 : x / p) - (omx == 0 ? (real)0 : omx / ((real)1 - p));
  dx = 0;  // a free Bernoulli value is a frozen coordinate of the gradient-based blocks (compound step)
  return (x == 0 ? (real)0 : x * lg) + (omx == 0 ? (real)0 : omx * l1);
}

// Poisson(rate = lam): xlogy(x, lam) - lam - lgamma(x + 1)
template <typename real>
__device__ __forceinline__ real poisson_lpdf(real x, real lam, real& dx, real& dlam) {
  dlam = (x == 0 ? (real)0 : x / lam) - (real)1;
  dx = 0;
  return (x == 0 ? (real)0 : x * m_log(lam)) - lam - m_lgamma(x + (real)1);
}

}  // namespace pm4
