// pm4_common.cuh -- types shared by the generated model code and the sampler kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pm4_dists.cuh"
#include "pm4_module.h"
#include "pm4_philox.cuh"

namespace pm4 {

constexpr unsigned FULL = 0xffffffffu;

// constant pools and run-time sizes the generated code reads (device pointers)
template <typename real> struct ModelArgs {
  const real* __restrict__ dataf;
  const int32_t* __restrict__ datai;
  const int32_t* __restrict__ term_n;
  const int32_t* __restrict__ op_blob;
  const real* __restrict__ planf;
  const int32_t* __restrict__ plani;
  const int32_t* __restrict__ plan_off;
  int n_planf, n_plani, part_reals, stage;
};

template <typename real> __host__ __device__ inline ModelArgs<real> make_model_args(const pm4_mod_args_t& a) {
  return ModelArgs<real>{(const real*)a.dataf, a.datai, a.term_n, a.op_blob, (const real*)a.planf, a.plani,
                         a.plan_off, a.n_planf, a.n_plani, a.part_reals, a.stage};
}

// per-warp evaluation context handed to the generated model code
template <typename real> struct Ctx {
  int lane;
  real* xs;           // shared memory [DP]: constrained values of this chain (gather source)
  real* gx;           // shared memory [DP]: adjoint scatter target of this chain
  real* part;         // shared memory [part_reals]: chunk partial sums of planned terms
  const real* pf;     // plan records (shared memory when staged, else global)
  const int32_t* pi;  // plan metadata (same)
  ModelArgs<real> A;
};

// record of a planned term: W consecutive reals read with one (vector) load
template <typename real, int W> struct Rec;
template <> struct Rec<float, 1> { using type = float; };
template <> struct Rec<float, 2> { using type = float2; };
template <> struct Rec<float, 4> { using type = float4; };
template <> struct Rec<double, 1> { using type = double; };
template <> struct Rec<double, 2> { using type = double2; };
template <> struct Rec<double, 4> { using type = double4; };

template <typename real> __device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Sum P (a power of two <= 32) per-lane values over the warp with a transposing butterfly: each
// exchange step halves the number of live values, so P = 8 costs 4+2+1+1+1 = 9 shuffles instead of
// 40.  On return v[0] of lane l is the warp total of value (l * P / 32); all 32/P lanes of a group
// hold bit-identical totals.
template <int P, typename real> __device__ __forceinline__ void warp_sum_packed(real (&v)[P], int lane) {
  int off = 16;
#pragma unroll
  for (int h = P; h > 1; h >>= 1) {
    const int half = h >> 1;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const real send = up ? v[k] : v[k + half];
      const real keep = up ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(FULL, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(FULL, v[0], off);
}

}  // namespace pm4
