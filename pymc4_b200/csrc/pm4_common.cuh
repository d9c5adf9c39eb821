// pm4_common.cuh -- types shared by the generated model code and the sampler kernels.
//
// Execution model: one TEAM of T warps owns one chain (T = Model::T, chosen per model by
// pymc4_b200/plan.py).  TL = 32 T lanes; every D-vector of the chain is spread over the team,
// lane tl holding elements tl, tl + TL, ... in registers.  Teams synchronise with a named
// barrier (bar.sync id, TL) -- a plain __syncwarp() when T = 1.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "pm4_dists.cuh"
#include "pm4_module.h"
#include "pm4_pair.cuh"
#include "pm4_philox.cuh"

namespace pm4 {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ELEM_CNT_BITS = 12;  // elem_tab entry = overflow first << 12 | count (pymc4_b200/plan.py)

// header of the owner-computes plan (8 ints behind the per-plan headers of plani)
struct OwnedHeader { int enabled, i_off, i_len, f_off, f_len, NF, Q, n; };

// constant pools and run-time sizes the generated code reads (device pointers)
template <typename real> struct ModelArgs {
  const real* __restrict__ dataf;
  const int32_t* __restrict__ datai;
  const int32_t* __restrict__ term_n;
  const int32_t* __restrict__ op_blob;
  const real* __restrict__ planf;
  const int32_t* __restrict__ plani;
  const int32_t* __restrict__ plan_off;
  int n_planf, n_plani, part_reals, elem_off, spart_off, phdr_off, stage;
  OwnedHeader oh;  // owner-computes plan of the warp-per-chain path (pymc4_b200/owned.py): host copy of its header
};

template <typename real> __host__ __device__ inline ModelArgs<real> make_model_args(const pm4_mod_args_t& a) {
  return ModelArgs<real>{(const real*)a.dataf, a.datai, a.term_n, a.op_blob, (const real*)a.planf, a.plani,
                         a.plan_off, a.n_planf, a.n_plani, a.part_reals, a.elem_off, a.spart_off, a.phdr_off, a.stage,
                         OwnedHeader{a.owned[0], a.owned[1], a.owned[2], a.owned[3], a.owned[4], a.owned[5], a.owned[6], a.owned[7]}};
}

// per-lane evaluation context handed to the generated model code
template <typename real> struct Ctx {
  int tl;             // lane within the team, 0 .. 32 T - 1
  int lane;           // lane within the warp
  int warp;           // warp within the team
  int bar;            // named barrier of the team
  int my_spart;       // part entry this lane parks its share of the packed reduction at, or -1
  int lp_first;       // first part entry of the log-density partials (one per warp)
  real* xs;           // shared memory [DP]: constrained values of this chain (gather source)
  real* gx;           // shared memory [DP]: adjoint scatter target (models with unplanned gathers only)
  real* part;         // shared memory [part_reals]: chunk / warp partial sums, grouped by gradient element
  real* red;          // shared memory [2][4][T]: cross-warp reductions of the sampler (double buffered)
  int32_t* ctl;       // shared memory [4]: team control words
  real* scr;          // shared memory tree scratch of the team (when it fits)
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

template <int T> __device__ __forceinline__ void team_sync(int bar) {
  if constexpr (T == 1) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(32 * T) : "memory");
}

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

// sum of T consecutive shared-memory reals in index order (vector loads where the type allows)
template <int T, typename real> __device__ __forceinline__ real sum_consecutive(const real* p) {
  if constexpr (sizeof(real) == 4 && T == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    return ((v.x + v.y) + v.z) + v.w;
  } else if constexpr (sizeof(real) == 4 && T == 2) {
    const float2 v = *reinterpret_cast<const float2*>(p);
    return v.x + v.y;
  } else if constexpr (sizeof(real) == 8 && T == 2) {
    const double2 v = *reinterpret_cast<const double2*>(p);
    return v.x + v.y;
  } else if constexpr (sizeof(real) == 8 && T == 4) {
    const double2 v = *reinterpret_cast<const double2*>(p), w = *reinterpret_cast<const double2*>(p + 2);
    return ((v.x + v.y) + w.x) + w.y;
  } else if constexpr (sizeof(real) == 4 && T % 4 == 0) {
    real s = 0;
#pragma unroll
    for (int w = 0; w < T; w += 4) {
      const float4 v = *reinterpret_cast<const float4*>(p + w);
      s = (((s + v.x) + v.y) + v.z) + v.w;
    }
    return s;
  } else {
    real s = p[0];
#pragma unroll
    for (int w = 1; w < T; ++w) s += p[w];
    return s;
  }
}

// Team-wide sums of P per-lane values; every lane of the team receives all P totals (bit-identical).
// `red` must hold P*T reals and alternate between two buffers from one call to the next.
template <int T, int P, typename real>
__device__ __forceinline__ void team_sum(real (&v)[P], const Ctx<real>& c, real* red) {
  constexpr int per = 32 / P;
  warp_sum_packed<P, real>(v, c.lane);
  if constexpr (T == 1) {
    const real t = v[0];
#pragma unroll
    for (int k = 0; k < P; ++k) v[k] = __shfl_sync(FULL, t, k * per);
  } else {
    if ((c.lane % per) == 0) red[(c.lane / per) * T + c.warp] = v[0];
    team_sync<T>(c.bar);
#pragma unroll
    for (int k = 0; k < P; ++k) v[k] = sum_consecutive<T, real>(red + k * T);
  }
}

}  // namespace pm4
