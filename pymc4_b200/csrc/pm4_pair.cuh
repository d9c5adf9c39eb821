// pm4_pair.cuh -- two observations per instruction.  Pk<float> is a float2 driven by the packed FP32
// instructions of sm_100 (fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2: SASS FFMA2 / FADD2 / FMUL2), one issue
// slot for two lanes of arithmetic -- the generated row loops of the warp-per-chain path are issue-bound, not
// pipe-bound.  Pk<double> is a plain pair (the FP64 parity build).  Each operation rounds exactly like its
// scalar counterpart, so a pair loop and a scalar loop over the same rows agree bit for bit per row.
#pragma once
#include <cuda_runtime.h>

namespace pm4 {

template <typename real> struct PkT;
template <> struct PkT<float> { using type = float2; };
template <> struct PkT<double> { using type = double2; };
template <typename real> using Pk = typename PkT<real>::type;

template <typename real> __device__ __forceinline__ Pk<real> pk_bcast(real v) { Pk<real> r; r.x = v; r.y = v; return r; }

__device__ __forceinline__ float2 pk_fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 pk_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 pk_neg(float2 a) { return make_float2(-a.x, -a.y); }
// a - b as fma(b, -1, a): one packed instruction, exact product, one rounding -- identical to a - b
__device__ __forceinline__ float2 pk_sub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
__device__ __forceinline__ float2 pk_fnma(float2 a, float2 b, float2 c) { return __ffma2_rn(pk_neg(a), b, c); }   // c - a b
__device__ __forceinline__ float2 pk_fms(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, pk_neg(c)); }    // a b - c

__device__ __forceinline__ double2 pk_fma(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
__device__ __forceinline__ double2 pk_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 pk_mul(double2 a, double2 b) { return make_double2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ double2 pk_neg(double2 a) { return make_double2(-a.x, -a.y); }
__device__ __forceinline__ double2 pk_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 pk_fnma(double2 a, double2 b, double2 c) { return make_double2(fma(-a.x, b.x, c.x), fma(-a.y, b.y, c.y)); }
__device__ __forceinline__ double2 pk_fms(double2 a, double2 b, double2 c) { return make_double2(fma(a.x, b.x, -c.x), fma(a.y, b.y, -c.y)); }

__device__ __forceinline__ float pk_hsum(float2 a) { return a.x + a.y; }
__device__ __forceinline__ double pk_hsum(double2 a) { return a.x + a.y; }

// pair record of W columns: [column][2 observations], one vector load
template <typename real, int W> struct PairRec;
template <> struct PairRec<float, 1> { using type = float2; };
template <> struct PairRec<float, 2> { using type = float4; };
template <> struct PairRec<double, 1> { using type = double2; };
template <> struct PairRec<double, 2> { using type = double4; };

template <typename real, int W> __device__ __forceinline__ Pk<real> rec_col(const typename PairRec<real, W>::type& q, int j) {
  Pk<real> r;
  if constexpr (W == 1) { r.x = q.x; r.y = q.y; }
  else { if (j == 0) { r.x = q.x; r.y = q.y; } else { r.x = q.z; r.y = q.w; } }
  return r;
}

}  // namespace pm4
