// pm4_glm.cu -- dense-GLM building block: C[M x N] = A[M x K] * B[N x K]^T on the 5th-generation
// tensor cores (tcgen05.mma kind::tf32, accumulator in tensor memory), FP32 inputs split into two
// TF32 terms each ("3xTF32": hi*hi + lo*hi + hi*lo) so that the product keeps FP32-level accuracy --
// the X*beta contraction of the hierarchical logistic regression (BASELINE configs[2]; the
// reference evaluates it as a broadcasted tf.tensordot inside tfd.Bernoulli's log_prob,
// /root/reference/pymc4/distributions/discrete.py:75-78).
//
// Called by libpm4b200 (pm4_core.cu) for PM4_T_BERNOULLI_LOGIT_GLM terms: pm4_logp_grad adds it to the per-chain
// pass of the generated module, pm4_hmc_run drives it once per lock-step leapfrog (pm4_lockstep.cuh).
//
// Test kernel: one CTA computes one 128 x 128 output tile.  Operands are staged by the CTA's threads into shared
// memory in the canonical K-major "interleave" (no-swizzle) layout the UMMA descriptors describe:
// 8-row x 16-byte core matrices, LBO between the two core matrices of one K-step, SBO between
// 8-row groups.  One thread issues the MMAs and commits them to an mbarrier; the four warps read
// the accumulator back with tcgen05.ld (warp w owns TMEM lanes 32w .. 32w+31).
#include <cuda_runtime.h>
#include <algorithm>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/pm4b200.h"

namespace {

constexpr int TILE = 128;     // M and N of one MMA tile
constexpr int UMMA_K = 8;     // tf32 elements per tcgen05.mma (32 bytes)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 in bits
// [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M x N
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of element (row, k) in the staged operand: core matrix = 8 rows x 4 tf32
__device__ __forceinline__ uint32_t op_offset(int row, int k, int K) {
  const uint32_t sbo = (uint32_t)(K / 4) * 128u;
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * 128u + (uint32_t)(row & 7) * 16u + (uint32_t)(k & 3) * 4u;
}

// C[128 x 128] (row-major, ldc) = A[128 x K] (row-major, lda) * B[128 x K]^T (row-major, ldb); K % 8 == 0
__global__ void __launch_bounds__(128, 1) gemm_nt_3xtf32_kernel(const float* __restrict__ A, int lda,
                                                                const float* __restrict__ B, int ldb,
                                                                float* __restrict__ C, int ldc, int K, int passes,
                                                                int* __restrict__ error) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t op_bytes = (uint32_t)TILE * (uint32_t)K * 4u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + op_bytes;
  unsigned char* b_hi = a_lo + op_bytes;
  unsigned char* b_lo = b_hi + op_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + op_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TILE));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // stage the operands: split every FP32 value into two TF32 terms
  for (int idx = tid; idx < TILE * K; idx += blockDim.x) {
    const int row = idx / K, k = idx - row * K;
    const float a = A[(size_t)row * lda + k], b = B[(size_t)row * ldb + k];
    const float ah = to_tf32(a), bh = to_tf32(b);
    const uint32_t off = op_offset(row, k, K);
    *reinterpret_cast<float*>(a_hi + off) = ah;
    *reinterpret_cast<float*>(a_lo + off) = to_tf32(a - ah);
    *reinterpret_cast<float*>(b_hi + off) = bh;
    *reinterpret_cast<float*>(b_lo + off) = to_tf32(b - bh);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = make_idesc(TILE, TILE);
    const uint32_t sbo = (uint32_t)(K / 4) * 128u, lbo = 128u;
    uint32_t acc = 0;
    for (int pass = 0; pass < passes; ++pass) {  // hi*hi, lo*hi, hi*lo
      const unsigned char* ap = pass == 1 ? a_lo : a_hi;
      const unsigned char* bp = pass == 2 ? b_lo : b_hi;
      for (int ks = 0; ks < K / UMMA_K; ++ks) {
        const uint64_t da = make_desc(smem_u32(ap) + (uint32_t)ks * 256u, lbo, sbo);
        const uint64_t db = make_desc(smem_u32(bp) + (uint32_t)ks * 256u, lbo, sbo);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
            "l"(da), "l"(db), "r"(idesc), "r"(acc)
            : "memory");
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  // wait for the accumulator (bounded: a malformed descriptor must not hang the GPU)
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
    if (!done && clock64() - t0 > 2000000000ll) break;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!done) {
    if (tid == 0 && error) *error = 1;
  } else {
    // warp w reads TMEM lanes 32w .. 32w+31 (= output rows), 32 columns at a time
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < TILE; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) C[(size_t)row * ldc + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TILE));
}


// ------------------------------------------------------------------------------------------
// Bernoulli-logit GLM likelihood for C chains in lock-step:
//   eta[n, c] = sum_p X[n, p] beta[c, p]
//   lp[c]    += sum_n y_n eta - softplus(eta)            (Bernoulli(probs = sigmoid(eta)).log_prob(y))
//   g[c, p]  += sum_n (y_n - sigmoid(eta[n, c])) X[n, p]
// Kernel 1 (eta): one CTA per (128 observations) x (128 chains): X tile and beta tile staged K-major,
//   3xTF32 MMAs into TMEM, epilogue from TMEM: residuals r = y - sigmoid(eta) to global, TRANSPOSED
//   (Rt[Cp, Np]: a lane owns an observation, so the stores of one chain are coalesced), the tile's column
//   sums of the log-likelihood to LP[tile, Cp] (deterministic: no atomics).
// Kernel 2 (grad): one CTA per (observation split) x (128 chains), looping over 32-observation tiles:
//   Rt tile and Xt tile (the covariates transposed once at model creation) staged K-major -- the
//   contraction runs over observations -- with 16-byte loads and stores, 3xTF32 MMAs accumulating
//   G[chain, p] in TMEM over all tiles of the split, then G to GP[split, Cp, Pp].
//   (MN-major TF32 operands in the no-swizzle layout produced zeros on this part; transposing the
//   operands in memory keeps both GEMMs on the K-major path.)
// Kernel 3: fixed-order reductions of LP and GP in double precision into lp[C] and grad[C, D].
// ------------------------------------------------------------------------------------------
constexpr int KT2 = 32;  // observations per tile of the gradient kernel (60 KB of operands: two CTAs per SM overlap staging and MMA)
constexpr int NT = 512;  // threads per CTA of the GLM kernels: 16 warps stage the operands and share the epilogue
                         // (warp w reads TMEM lanes 32 (w % 4) .. + 31, column block w / 4)

__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
// bounded wait on an mbarrier phase; false on time-out
__device__ __forceinline__ bool mbar_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 4000000000ll) return false;
  }
  return true;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// sum of 32 per-lane values over the warp, transposed: lane j ends up with the total of v[j]
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
  int off = 16;
#pragma unroll
  for (int h = 32; h > 1; h >>= 1) {
    const int half = h >> 1;
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < half; ++k) {
      const float send = up ? v[k] : v[k + half];
      const float keep = up ? v[k + half] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  // value index held by lane l after the butterfly: bit-reversed interleave -> lane l holds column brev5(l)? no:
  // the step with offset `off` sends the upper half of the live values to the lanes with that bit set, so lane l
  // holds value ((l>>4)&1)*16 + ((l>>3)&1)*8 + ((l>>2)&1)*4 + ((l>>1)&1)*2 + (l&1) = l
  return v[0];
}

__global__ void __launch_bounds__(NT, 1) glm_eta_kernel(const float* __restrict__ X, const float* __restrict__ y, int n,
                                                         int P, const float* __restrict__ theta, int ldt, int base, int C,
                                                         float* __restrict__ Rt, int Np, int Cp,
                                                         float* __restrict__ LP, int* __restrict__ error) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int K = (P + 7) & ~7, K4 = K >> 2;
  const uint32_t op_bytes = (uint32_t)TILE * (uint32_t)K * 4u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + op_bytes;
  unsigned char* b_hi = a_lo + op_bytes;
  unsigned char* b_lo = b_hi + op_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + op_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  float* colsum = reinterpret_cast<float*>(mbar + 2);  // [4][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * TILE, n_tiles = (n + TILE - 1) / TILE;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * TILE));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  auto split_store = [&](const float4& v, unsigned char* hi, unsigned char* lo, uint32_t off) {
    const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
  };
  // the CTA's chain tile beta[c0 .. c0+127, :] is staged ONCE (zero padded, split into two TF32 terms); the CTA then
  // walks over observation tiles blockIdx.x, blockIdx.x + gridDim.x, ...
  // (index map: 8 consecutive lanes take the 8 rows of one core matrix at the same k -- 128 contiguous bytes of
  //  shared memory, no bank conflicts -- and the next 8 lanes the next 4 covariates of the same rows)
  for (int idx = tid; idx < TILE * K4; idx += NT) {
    const int q = idx >> 3, row = (q / K4) * 8 + (idx & 7), k = (q % K4) * 4;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c0 + row < C && k < P) {
      const float* tp = theta + (size_t)(c0 + row) * ldt + base + k;
      b.x = tp[0];
      if (k + 1 < P) b.y = tp[1];
      if (k + 2 < P) b.z = tp[2];
      if (k + 3 < P) b.w = tp[3];
    }
    split_store(b, b_hi, b_lo, op_offset(row, k, K));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int lg = warp & 3, row = lg * 32 + lane;
  uint32_t parity = 0;
  // X[n0 .. n0+127, :] travels through registers, 16 bytes at a time (P is a multiple of 4: a group of 4 covariates
  // is whole or absent; rows of X are 16-byte aligned): the loads of the NEXT tile are issued before the MMAs of this
  // one complete, so their latency hides behind the tensor core and the epilogue
  constexpr int AU = (TILE * 26 + NT - 1) / NT;  // K <= 104: at most 7 groups per thread
  float4 ra[AU];
  const bool vec4 = (P & 3) == 0;
  auto fetch = [&](int n0) {
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      const int idx = tid + u * NT, q = idx >> 3, r = (q / K4) * 8 + (idx & 7), k = (q % K4) * 4;
      ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < TILE * K4 && n0 + r < n && k < P) {
        const float* xp = X + (size_t)(n0 + r) * P + k;
        if (vec4) ra[u] = __ldg(reinterpret_cast<const float4*>(xp));  // P % 4 == 0: rows are 16-byte aligned
        else {
          ra[u].x = xp[0];
          if (k + 1 < P) ra[u].y = xp[1];
          if (k + 2 < P) ra[u].z = xp[2];
          if (k + 3 < P) ra[u].w = xp[3];
        }
      }
    }
  };
  // The accumulator is double-buffered in TMEM (2 x 128 columns): the MMAs of tile t + 1 are issued BEFORE the
  // epilogue of tile t, so the tensor core works while the CUDA cores do the Bernoulli arithmetic and the stores.
  // One X buffer is enough: it is rewritten only after the MMAs that read it have completed.
  auto stage_and_issue = [&](uint32_t buf) {   // registers -> shared memory (split), then the 3 x K/8 MMAs
#pragma unroll
    for (int u = 0; u < AU; ++u) {
      const int idx = tid + u * NT, q = idx >> 3, r = (q / K4) * 8 + (idx & 7), k = (q % K4) * 4;
      if (idx < TILE * K4) split_store(ra[u], a_hi, a_lo, op_offset(r, k, K));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t idesc = make_idesc(TILE, TILE);
      const uint32_t sbo = (uint32_t)(K / 4) * 128u, lbo = 128u;
      uint32_t acc = 0;
      for (int pass = 0; pass < 3; ++pass) {  // hi*hi, lo*hi, hi*lo
        const unsigned char* ap = pass == 1 ? a_lo : a_hi;
        const unsigned char* bp = pass == 2 ? b_lo : b_hi;
        for (int ks = 0; ks < K / UMMA_K; ++ks) {
          mma_tf32(tmem + buf * (uint32_t)TILE, make_desc(smem_u32(ap) + (uint32_t)ks * 256u, lbo, sbo),
                   make_desc(smem_u32(bp) + (uint32_t)ks * 256u, lbo, sbo), idesc, acc);
          acc = 1;
        }
      }
      mma_commit(mbar);
    }
  };
  const int stride = (int)gridDim.x;
  uint32_t buf = 0;
  if ((int)blockIdx.x < n_tiles) {
    fetch(blockIdx.x * TILE);
    stage_and_issue(0);
    if ((int)blockIdx.x + stride < n_tiles) fetch(((int)blockIdx.x + stride) * TILE);
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += stride, buf ^= 1u) {
    const int n0 = tile * TILE;
    const bool ok = mbar_wait(mbar, parity);  // the MMAs of this tile
    parity ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!ok) {  // CTA-uniform (every thread waits on the same phase)
      if (tid == 0 && error) *error = 1;
      break;
    }
    if (tile + stride < n_tiles) {  // next tile: its X rows are in registers, its MMAs go to the other accumulator
      stage_and_issue(buf ^ 1u);
      if (tile + 2 * stride < n_tiles) fetch((tile + 2 * stride) * TILE);
    }
    const bool live = n0 + row < n;
    const float yv = live ? y[n0 + row] : 0.0f;
    for (int cc = (warp >> 2) * 32; cc < TILE; cc += (NT / 128) * 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + buf * (uint32_t)TILE + (uint32_t)cc, v);
      float ll[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float eta = __uint_as_float(v[j]);
        // softplus(eta) = max(eta, 0) + log1p(exp(-|eta|)); sigmoid from the same exponential
        // (MUFU ex2 / lg2 / rcp: 1 + e lies in (1, 2], so the absolute error of each term stays ~1e-7)
        const float e = __expf(-fabsf(eta));
        const float sp = fmaxf(eta, 0.0f) + __logf(1.0f + e);
        const float rc = __fdividef(1.0f, 1.0f + e);
        const float sg = eta >= 0.0f ? rc : e * rc;
        ll[j] = live ? yv * eta - sp : 0.0f;
        v[j] = __float_as_uint(live ? yv - sg : 0.0f);
      }
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; ++j) Rt[(size_t)(c0 + cc + j) * Np + n0 + row] = __uint_as_float(v[j]);
      }
      const float cs = warp_colsum32(ll, lane);
      colsum[lg * TILE + cc + lane] = cs;
    }
    __syncthreads();
    if (tid < TILE)
      LP[(size_t)tile * Cp + c0 + tid] = ((colsum[tid] + colsum[TILE + tid]) + colsum[2 * TILE + tid]) + colsum[3 * TILE + tid];
    // this accumulator and colsum are reused two / one tiles from now
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * TILE));
}

// G[chain, p] = sum over the observations of this split of Rt[chain, n] * Xt[p, n]
__global__ void __launch_bounds__(NT, 2) glm_grad_kernel(const float* __restrict__ Xt, int ldx, int n, int P,
                                                          const float* __restrict__ Rt, int Np, int Cp,
                                                          float* __restrict__ GP, int Pp, int n_splits,
                                                          int* __restrict__ error) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t a_bytes = (uint32_t)TILE * KT2 * 4u, b_bytes = (uint32_t)Pp * KT2 * 4u;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + a_bytes;
  unsigned char* b_hi = a_lo + a_bytes;
  unsigned char* b_lo = b_hi + b_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + b_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, c0 = blockIdx.y * TILE;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TILE));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int n_tiles = (n + KT2 - 1) / KT2;
  uint32_t acc = 0, parity = 0;
  bool ok = true;
  // Each thread carries one tile's share of the operands in registers: 2 x 16 bytes of Rt (128 rows x 32
  // observations) and up to 2 x 16 bytes of Xt (Pp rows).  The loads of tile t + 1 are issued right after the
  // MMAs of tile t, so the memory latency overlaps the tensor core (and the co-resident CTA).
  constexpr int Q = KT2 / 4;  // 16-byte groups per row
  float4 ra[2], rb[2];
  auto fetch = [&](int nb) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + u * NT, q = idx >> 3, row = (q / Q) * 8 + (idx & 7), k = (q % Q) * 4;
      ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      rb[u] = ra[u];
      const bool whole = nb + k + 3 < n;  // observations >= n: Rt is not written there, both operands read as zero
      if (row < TILE) {
        const float* p = Rt + (size_t)(c0 + row) * Np + nb + k;
        if (whole) ra[u] = *reinterpret_cast<const float4*>(p);
        else {
          if (nb + k < n) ra[u].x = p[0];
          if (nb + k + 1 < n) ra[u].y = p[1];
          if (nb + k + 2 < n) ra[u].z = p[2];
        }
      }
      if (row < P) {
        const float* p = Xt + (size_t)row * ldx + nb + k;
        if (whole) rb[u] = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (nb + k < n) rb[u].x = p[0];
          if (nb + k + 1 < n) rb[u].y = p[1];
          if (nb + k + 2 < n) rb[u].z = p[2];
        }
      }
    }
  };
  auto split_store = [&](const float4& v, unsigned char* hi, unsigned char* lo, uint32_t off) {
    const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
  };
  if (split < n_tiles) fetch(split * KT2);
  for (int t = split; t < n_tiles && ok; t += n_splits) {
    // the operand buffers are reused: wait until the tensor core has consumed the previous tile
    if (acc) {
      ok = mbar_wait(mbar, parity);
      parity ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (!ok) break;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = tid + u * NT, q = idx >> 3, row = (q / Q) * 8 + (idx & 7), k = (q % Q) * 4;
      const uint32_t off = op_offset(row, k, KT2);
      if (row < TILE) split_store(ra[u], a_hi, a_lo, off);
      if (row < Pp) split_store(rb[u], b_hi, b_lo, off);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      const uint32_t idesc = make_idesc(TILE, Pp);
      const uint32_t sbo = (uint32_t)Q * 128u, lbo = 128u;
      for (int pass = 0; pass < 3; ++pass) {
        const unsigned char* ap = pass == 1 ? a_lo : a_hi;
        const unsigned char* bp = pass == 2 ? b_lo : b_hi;
        for (int ks = 0; ks < KT2 / UMMA_K; ++ks)
          mma_tf32(tmem, make_desc(smem_u32(ap) + (uint32_t)ks * 256u, lbo, sbo),
                   make_desc(smem_u32(bp) + (uint32_t)ks * 256u, lbo, sbo), idesc, (pass | ks) ? 1u : acc);
      }
      mma_commit(mbar);
    }
    acc = 1;  // CTA-uniform: the accumulator holds at least one tile from here on
    if (t + n_splits < n_tiles) fetch((t + n_splits) * KT2);
  }
  if (ok && acc) {  // the last tile's MMAs
    ok = mbar_wait(mbar, parity);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (!ok) {
    if (tid == 0 && error) *error = 2;
  } else {
    const int lg = warp & 3, row = lg * 32 + lane;  // chain within the tile
    float* dst = GP + ((size_t)split * Cp + c0 + row) * Pp;
    for (int cc = (warp >> 2) * 32; cc < Pp; cc += (NT / 128) * 32) {
      uint32_t v[32];
      if (acc) tmem_ld32(tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)cc, v);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (cc + j < Pp) dst[cc + j] = acc ? __uint_as_float(v[j]) : 0.0f;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TILE));
}

// first level of the log-likelihood reduction: group g of the observation tiles, 32 chains x 8 tile lanes per block,
// fixed order (LP2[g, c] = sum over the group's tiles, double)
constexpr int LP_GROUPS = 64;
__global__ void __launch_bounds__(256) glm_lp_kernel(const float* __restrict__ LP, int n_tiles, int Cp, double* __restrict__ LP2) {
  __shared__ double red[8][33];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + cx, g = blockIdx.y;
  const int per = (n_tiles + LP_GROUPS - 1) / LP_GROUPS, t0 = g * per, t1 = min(n_tiles, t0 + per);
  double s = 0;
  for (int t = t0 + ty; t < t1; t += 8) s += (double)LP[(size_t)t * Cp + c];
  red[ty][cx] = s;
  __syncthreads();
  if (ty == 0) {
    for (int k = 1; k < 8; ++k) s += red[k][cx];
    LP2[(size_t)g * Cp + c] = s;
  }
}

// fixed-order reductions in double precision: lp[c] += sum_groups LP2[g, c]; grad[c, base + p] += sum_splits GP[s, c, p]
__global__ void glm_reduce_kernel(const double* __restrict__ LP2, const float* __restrict__ GP, int n_splits,
                                  int Cp, int Pp, int C, int P, float* __restrict__ lp, float* __restrict__ grad, int ldg,
                                  int base) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C) {
    double s = 0;
    for (int g = 0; g < LP_GROUPS; ++g) s += LP2[(size_t)g * Cp + idx];
    lp[idx] += (float)s;
  }
  if (grad && idx < C * P) {
    const int c = idx / P, p = idx - c * P;
    double s = 0;
    for (int k = 0; k < n_splits; ++k) s += (double)GP[((size_t)k * Cp + c) * Pp + p];
    grad[(size_t)c * ldg + base + p] += (float)s;
  }
}


// ------------------------------------------------------------------------------------------
// Fused evaluation (value + gradient in ONE kernel; the residual matrix never leaves the SM):
//   per (32 observations) x (128 chains) tile
//     GEMM1   eta^T[chain, obs]  = beta[128 x K] . Xtile[32 x K]^T          -> TMEM (double-buffered, 2 x 32 columns)
//     epilogue r = y - sigmoid(eta), ll += y eta - softplus(eta); r split hi/lo and written back to TENSOR MEMORY as
//              GEMM2's A operand (beta, GEMM1's A operand, sits there too): with both A operands out of shared
//              memory an MMA only reads its small B block, so the 128 B/cycle shared-memory port no longer bounds
//              the 32-observation tiles (with A in shared memory every MMA re-read 4 KB of it)
//     GEMM2   G[chain, p]       += R[128 x 32] . Xt_tile[Pp x 32]^T          -> TMEM (Pp columns, lives for the whole CTA)
// The constant operand is PRE-SPLIT once at upload (glm_image_kernel): per tile the two TF32 terms of the X rows as
// GEMM1's B image and of the transposed tile as GEMM2's B image, already in the core-matrix layout the UMMA
// descriptors describe, so a tile arrives with two TMA bulk copies and no CUDA core touches it.  Roles (warp
// specialised, mbarrier pipeline): warp 0 = TMA producer, warp 1 = MMA issuer of GEMM1, warp 2 = of GEMM2, warps 3..18 = epilogue in two groups that take alternate tiles (warp w owns TMEM
// lanes 32 (w % 4) .. +31 = 32 chains, and 16 of the tile's 32 observations).  A lane of the accumulator is a CHAIN
// here, so the residuals of one chain are 16 consecutive K elements (columns) of GEMM2's A operand -- one tcgen05.st
// per TF32 term -- and the log-likelihood sum of a chain needs no cross-lane reduction.
// ------------------------------------------------------------------------------------------
constexpr int FT = 32;                               // observations per tile
constexpr int F_EPI = 16;                            // epilogue warps: two groups of 8 that take alternate tiles
constexpr int F_NQ = 2, F_CW = FT / F_NQ;            // column groups per tile (16 observations = two K steps of GEMM2)
constexpr int F_W_G1 = 1, F_W_G2 = 2, F_W_EPI = 3;   // warp roles: 0 TMA producer, 1 GEMM1 issuer, 2 GEMM2 issuer, 3.. epilogue
constexpr int F_THREADS = 32 * (F_W_EPI + F_EPI);
constexpr int F_NSX = 4, F_NSXT = 2;                 // pipeline stages of the two operand images
// tensor-memory map (512 columns, all of them): eta[buffer] | G | beta hi | beta lo | R[buffer][hi, lo].  Everything a
// tile touches is double-buffered, so GEMM1 of tile i + 1 and GEMM2 of tile i - 1 run under the epilogue of tile i, and
// the two epilogue groups (even / odd tiles) are half a period apart: the latencies of one (accumulator load, MUFU
// chains, tcgen05.st, barrier round trips) hide behind the arithmetic of the other.
constexpr uint32_t F_TMEM_COLS = 512;
constexpr uint32_t F_G_COL = 64, F_BHI_COL = 176, F_BLO_COL = 280, F_R_COL = 384;
enum { FB_FULL_X = 0, FB_EMPTY_X = FB_FULL_X + F_NSX, FB_FULL_XT = FB_EMPTY_X + F_NSX, FB_EMPTY_XT = FB_FULL_XT + F_NSXT,
       FB_ETA_FULL = FB_EMPTY_XT + F_NSXT, FB_ETA_EMPTY = FB_ETA_FULL + 2, FB_R_FULL = FB_ETA_EMPTY + 2 /* [buffer][column group] */,
       FB_R_EMPTY = FB_R_FULL + 2 * F_NQ, FB_G_DONE = FB_R_EMPTY + 2 * F_NQ, FB_COUNT };

__device__ __forceinline__ bool mbar_wait_short(uint64_t* mbar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (!done) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    if (!done && clock64() - t0 > 400000000ll) return false;
  }
  return true;
}
__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, uint64_t* mbar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A operand in TENSOR MEMORY (tcgen05.mma ".ts" form): lane i of the A block holds row i, one 32-bit column per tf32
// element of K.  A constant (or register-produced) A operand then costs no shared-memory bandwidth per MMA -- with the
// operand in shared memory every 128 x N x 8 MMA re-reads the 4 KB A block, which bounds small-N tiles at 128 bytes per
// cycle.  Test kernel for that form: same contract as gemm_nt_3xtf32_kernel.
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ float fast_ex2(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_lg2(float v) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float fast_rcp(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) gemm_ts_3xtf32_kernel(const float* __restrict__ A, int lda,
                                                                const float* __restrict__ B, int ldb,
                                                                float* __restrict__ C, int ldc, int K, int passes,
                                                                int* __restrict__ error) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t op_bytes = (uint32_t)TILE * (uint32_t)K * 4u;
  unsigned char* b_hi = smem;
  unsigned char* b_lo = b_hi + op_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + op_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t A_HI = 128, A_LO = 256;  // columns of the two A terms (K <= 104 each); D at 0 .. 127
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < TILE * K; idx += blockDim.x) {
    const int row = idx / K, k = idx - row * K;
    const float b = B[(size_t)row * ldb + k];
    const float bh = to_tf32(b);
    const uint32_t off = op_offset(row, k, K);
    *reinterpret_cast<float*>(b_hi + off) = bh;
    *reinterpret_cast<float*>(b_lo + off) = to_tf32(b - bh);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  {  // row = this thread's TMEM lane: its K values go to consecutive columns, 8 at a time
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = A[(size_t)row * lda + k0 + j];
        const float h = to_tf32(a);
        hi[j] = __float_as_uint(h);
        lo[j] = __float_as_uint(to_tf32(a - h));
      }
      tmem_st8(lane_addr + A_HI + (uint32_t)k0, hi);
      tmem_st8(lane_addr + A_LO + (uint32_t)k0, lo);
    }
    tmem_st_wait();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = make_idesc(TILE, TILE);
    const uint32_t sbo = (uint32_t)(K / 4) * 128u, lbo = 128u;
    uint32_t acc = 0;
    for (int pass = 0; pass < passes; ++pass) {  // hi*hi, lo*hi, hi*lo
      const uint32_t ac = pass == 1 ? A_LO : A_HI;
      const unsigned char* bp = pass == 2 ? b_lo : b_hi;
      for (int ks = 0; ks < K / UMMA_K; ++ks) {
        mma_tf32_ts(tmem, tmem + ac + (uint32_t)ks * UMMA_K, make_desc(smem_u32(bp) + (uint32_t)ks * 256u, lbo, sbo), idesc, acc);
        acc = 1;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
  }
  const bool done = mbar_wait_short(mbar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!done) {
    if (tid == 0 && error) *error = 1;
  } else {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < TILE; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) C[(size_t)row * ldc + c0 + j] = __uint_as_float(v[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// bytes of one tile of the pre-split image: [X rows hi | lo] (GEMM1's B, 32 x K) then [X^T hi | lo] (GEMM2's B, Pp x 32)
__host__ __device__ inline size_t fused_tile_bytes(int K, int Pp) { return 2 * (size_t)FT * K * 4 + 2 * (size_t)Pp * FT * 4; }

__global__ void __launch_bounds__(256) glm_image_kernel(const float* __restrict__ X, int n, int P, int K, int Pp,
                                                        unsigned char* __restrict__ img) {
  const int tile = blockIdx.x, n0 = tile * FT;
  unsigned char* xa_hi = img + (size_t)tile * fused_tile_bytes(K, Pp);
  unsigned char* xa_lo = xa_hi + (size_t)FT * K * 4;
  unsigned char* xb_hi = xa_lo + (size_t)FT * K * 4;
  unsigned char* xb_lo = xb_hi + (size_t)Pp * FT * 4;
  for (int idx = threadIdx.x; idx < FT * K; idx += blockDim.x) {
    const int r = idx / K, k = idx - r * K;
    const float v = (n0 + r < n && k < P) ? X[(size_t)(n0 + r) * P + k] : 0.0f;
    const float h = to_tf32(v);
    const uint32_t off = op_offset(r, k, K);
    *reinterpret_cast<float*>(xa_hi + off) = h;
    *reinterpret_cast<float*>(xa_lo + off) = to_tf32(v - h);
  }
  for (int idx = threadIdx.x; idx < Pp * FT; idx += blockDim.x) {
    const int k = idx / Pp, p = idx - k * Pp;  // consecutive threads read consecutive covariates of one observation
    const float v = (p < P && n0 + k < n) ? X[(size_t)(n0 + k) * P + p] : 0.0f;
    const float h = to_tf32(v);
    const uint32_t off = op_offset(p, k, FT);
    *reinterpret_cast<float*>(xb_hi + off) = h;
    *reinterpret_cast<float*>(xb_lo + off) = to_tf32(v - h);
  }
}

__global__ void __launch_bounds__(F_THREADS, 1) glm_fused_kernel(const unsigned char* __restrict__ img,
                                                                 const float* __restrict__ y, int n, int P,
                                                                 const float* __restrict__ theta, int ldt, int base, int C,
                                                                 double* __restrict__ LPp, float* __restrict__ GP, int Cp,
                                                                 int Pp, int* __restrict__ pace, const int* __restrict__ gate,
                                                                 int* __restrict__ error) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (gate && *(const volatile int*)gate == 0) return;  // batched lock-step iterations: nobody asked for this gradient
  const int K = (P + 7) & ~7, K4 = K >> 2;
  const uint32_t xa_part = (uint32_t)FT * K * 4u, xb_part = (uint32_t)Pp * FT * 4u;
  unsigned char* xa0 = smem;                           // F_NSX stages of [hi | lo] (GEMM1's B)
  unsigned char* xb0 = xa0 + (size_t)F_NSX * 2 * xa_part;  // F_NSXT stages of [hi | lo] (GEMM2's B)
  uint64_t* bars = reinterpret_cast<uint64_t*>(xb0 + (size_t)F_NSXT * 2 * xb_part);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + FB_COUNT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.y * TILE, n_tiles = (n + FT - 1) / FT;
  const int stride = (int)gridDim.x;
  const int nt = ((int)blockIdx.x < n_tiles) ? (n_tiles - (int)blockIdx.x + stride - 1) / stride : 0;
  const size_t tile_bytes = fused_tile_bytes(K, Pp);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(F_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int b = 0; b < FB_COUNT; ++b) {
      const uint32_t cnt = (b >= FB_ETA_EMPTY && b < FB_ETA_EMPTY + 2) ? (uint32_t)(F_EPI / 2)
                           : (b >= FB_R_FULL && b < FB_R_FULL + 2 * F_NQ) ? 4u
                                                                          : 1u;
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + b)), "r"(cnt));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const bool epi = warp >= F_W_EPI;
  // epilogue warps: TMEM lane quarter q, column group h of the tile, tile parity grp (group 0: even tiles, 1: odd)
  const int q = warp & 3, ew = (warp - F_W_EPI) >> 2, h = ew & (F_NQ - 1), grp = ew >> 1, row = q * 32 + lane;
  const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16);
  if (epi) {
    // the CTA's chain tile beta[c0 .. c0 + 127, :] goes to tensor memory once, split into its two TF32 terms: GEMM1's
    // A operand (lane = chain, one column per covariate), which then costs no shared-memory bandwidth
    const bool live_row = c0 + row < C;
    const float* tp = theta + (size_t)(c0 + row) * ldt + base;
    for (int k0 = 8 * ew; k0 < K; k0 += 8 * (F_EPI / 4)) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = (live_row && k0 + j < P) ? tp[k0 + j] : 0.0f;
        const float hv = to_tf32(v);
        hi[j] = __float_as_uint(hv);
        lo[j] = __float_as_uint(to_tf32(v - hv));
      }
      tmem_st8(lane_addr + F_BHI_COL + (uint32_t)k0, hi);
      tmem_st8(lane_addr + F_BLO_COL + (uint32_t)k0, lo);
    }
    tmem_st_wait();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  bool ok = true;

  if (warp == 0) {
    // ---- TMA producer ------------------------------------------------------------------------------------------
    // The CTAs of one observation group (same blockIdx.x, one per chain tile) stream the SAME tile images: the first
    // one to ask brings a tile in from HBM, the others find it in L2 -- as long as they stay close together.  They
    // drift apart by a few per cent over ~1700 tiles, which is more than L2 holds, so every PACE_EVERY tiles a
    // producer publishes its position and waits until the slowest CTA of its group is within PACE_WINDOW tiles
    // (`pace` is NULL when the grid is not co-resident).
    constexpr int PACE_EVERY = 8, PACE_WINDOW = 24;
    if (nt > 0) {
      auto src = [&](int i) { return img + (size_t)((int)blockIdx.x + i * stride) * tile_bytes; };
      auto load_x = [&](int i) { tma_load(xa0 + (size_t)(i % F_NSX) * 2 * xa_part, src(i), 2 * xa_part, bars + FB_FULL_X + (i % F_NSX)); };
      auto load_xt = [&](int i) { tma_load(xb0 + (size_t)(i % F_NSXT) * 2 * xb_part, src(i) + 2 * xa_part, 2 * xb_part, bars + FB_FULL_XT + (i % F_NSXT)); };
      if (lane == 0) {
        for (int i = 0; i < F_NSX && i < nt; ++i) load_x(i);
        for (int i = 0; i < F_NSXT && i < nt; ++i) load_xt(i);
      }
      int* const my_pace = pace ? pace + (size_t)blockIdx.x * gridDim.y : nullptr;
      for (int i = 0; i < nt && ok; ++i) {
        if (my_pace && (i % PACE_EVERY) == 0) {
          if (lane == 0) atomicMax(my_pace + blockIdx.y, i);
          const long long t0 = clock64();
          for (;;) {  // lane l watches chain tile l of the group
            int lo = i;
            for (int c = lane; c < (int)gridDim.y; c += 32) lo = min(lo, *((volatile int*)my_pace + c));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            if (lo + PACE_WINDOW >= i || clock64() - t0 > 20000000ll) break;
            __nanosleep(200);
          }
        }
        if (lane == 0) {
          if (i + F_NSX < nt) {  // the stage is free once GEMM1 of tile i has read it
            ok = mbar_wait_short(bars + FB_EMPTY_X + (i % F_NSX), (uint32_t)(i / F_NSX) & 1u);
            if (ok) load_x(i + F_NSX);
          }
          if (ok && i + F_NSXT < nt) {  // ... once GEMM2 of tile i has read it
            ok = mbar_wait_short(bars + FB_EMPTY_XT + (i % F_NSXT), (uint32_t)(i / F_NSXT) & 1u);
            if (ok) load_xt(i + F_NSXT);
          }
        }
        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
      }
      if (my_pace && lane == 0) atomicMax(my_pace + blockIdx.y, 0x7fffffff);  // done: never hold the others back
      if (!ok && lane == 0 && error) *error = 3;
    }
  } else if (warp == F_W_G1) {
    // ---- MMA issuer of GEMM1 (eta tiles): A = beta in tensor memory, B = the X rows of the tile in shared memory.
    // The whole warp runs the loop (waits are warp-uniform); one elected lane issues -- under `if (lane == 0)` the
    // compiler wraps every MMA in an election loop and one thread cannot issue 39 small MMAs per tile fast enough.
    // (a K step of 8 tf32 = 256 bytes = 16 units of the descriptor's 16-byte address field)
    if (nt > 0) {
      const uint32_t idesc1 = make_idesc(TILE, FT);
      const uint32_t sbo1 = (uint32_t)K4 * 128u, lbo = 128u;
      const int KS = K / UMMA_K;
      const uint32_t a_hi = tmem + F_BHI_COL, a_lo = tmem + F_BLO_COL;
      for (int i = 0; i < nt && ok; ++i) {
        const int s2 = i % F_NSX, bsel = i & 1;
        if (i >= 2) ok = mbar_wait_short(bars + FB_ETA_EMPTY + bsel, (uint32_t)((i - 2) >> 1) & 1u);  // the epilogue has read tile i - 2
        if (ok) ok = mbar_wait_short(bars + FB_FULL_X + s2, (uint32_t)(i / F_NSX) & 1u);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t dx_hi = make_desc(smem_u32(xa0 + (size_t)s2 * 2 * xa_part), lbo, sbo1);
          const uint64_t dx_lo = dx_hi + (uint64_t)(xa_part >> 4);
          const uint32_t d = tmem + (uint32_t)bsel * FT;
          mma_tf32_ts(d, a_hi, dx_hi, idesc1, 0u);  // hi*hi
#pragma unroll 4
          for (int ks = 1; ks < KS; ++ks) mma_tf32_ts(d, a_hi + (uint32_t)(ks * UMMA_K), dx_hi + (uint64_t)(ks * 16), idesc1, 1u);
#pragma unroll 4
          for (int ks = 0; ks < KS; ++ks) mma_tf32_ts(d, a_lo + (uint32_t)(ks * UMMA_K), dx_hi + (uint64_t)(ks * 16), idesc1, 1u);  // lo*hi
#pragma unroll 4
          for (int ks = 0; ks < KS; ++ks) mma_tf32_ts(d, a_hi + (uint32_t)(ks * UMMA_K), dx_lo + (uint64_t)(ks * 16), idesc1, 1u);  // hi*lo
          mma_commit(bars + FB_ETA_FULL + bsel);
          mma_commit(bars + FB_EMPTY_X + s2);
        }
        __syncwarp();
      }
      if (!ok && lane == 0 && error) *error = 4;
    }
  } else if (warp == F_W_G2) {
    // ---- MMA issuer of GEMM2 (gradient accumulator): A = the residuals in tensor memory (written by the epilogue),
    // B = the transposed tile
    if (nt > 0) {
      const uint32_t idesc2 = make_idesc(TILE, Pp);
      const uint32_t sbo2 = (uint32_t)(FT / 4) * 128u, lbo = 128u;
      const uint32_t d = tmem + F_G_COL;
      uint32_t acc_g = 0;
      for (int i = 0; i < nt && ok; ++i) {
        const int s2 = i % F_NSXT, bsel = i & 1;
        ok = mbar_wait_short(bars + FB_FULL_XT + s2, (uint32_t)(i / F_NSXT) & 1u);
        const uint64_t dt_hi = make_desc(smem_u32(xb0 + (size_t)s2 * 2 * xb_part), lbo, sbo2);
        const uint64_t dt_lo = make_desc(smem_u32(xb0 + (size_t)s2 * 2 * xb_part + xb_part), lbo, sbo2);
        const uint32_t r_hi = tmem + F_R_COL + (uint32_t)bsel * 64u, r_lo = r_hi + 32u;
        for (int hh = 0; hh < F_NQ && ok; ++hh) {  // the column groups of R (16 observations = two K steps) become ready independently
          ok = mbar_wait_short(bars + FB_R_FULL + F_NQ * bsel + hh, (uint32_t)(i >> 1) & 1u);
          if (!ok) break;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < F_CW / UMMA_K; ++kk) {
              const uint32_t ck = (uint32_t)(hh * (F_CW / UMMA_K) + kk) * UMMA_K;
              const uint64_t o = (uint64_t)((hh * (F_CW / UMMA_K) + kk) * 16);
              mma_tf32_ts(d, r_hi + ck, dt_hi + o, idesc2, (kk == 0) ? acc_g : 1u);
              mma_tf32_ts(d, r_lo + ck, dt_hi + o, idesc2, 1u);
              mma_tf32_ts(d, r_hi + ck, dt_lo + o, idesc2, 1u);
            }
            mma_commit(bars + FB_R_EMPTY + F_NQ * bsel + hh);
            if (hh == F_NQ - 1) {
              mma_commit(bars + FB_EMPTY_XT + s2);
              if (i == nt - 1) mma_commit(bars + FB_G_DONE);
            }
          }
          __syncwarp();
          acc_g = 1;
        }
      }
      if (!ok && lane == 0 && error) *error = 6;
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes 32 (w % 4) .. + 31 (chains), observations F_CW h .. F_CW h + F_CW - 1 of the
    // tiles of its group's parity ----
    double ll_acc = 0.0;
    float yv[F_CW];
    auto fetch_y = [&](int i) {  // the tile's observations of this column group (prefetched one tile of the group ahead)
      const int o0 = ((int)blockIdx.x + i * stride) * FT + F_CW * h;
      if (o0 + F_CW <= n) {
#pragma unroll
        for (int j4 = 0; j4 < F_CW / 4; ++j4) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(y + o0) + j4);
          yv[4 * j4] = a4.x; yv[4 * j4 + 1] = a4.y; yv[4 * j4 + 2] = a4.z; yv[4 * j4 + 3] = a4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < F_CW; ++j) yv[j] = o0 + j < n ? __ldg(y + o0 + j) : 0.0f;
      }
    };
    if (grp < nt) fetch_y(grp);
    for (int i = grp; i < nt && ok; i += 2) {
      const int o0 = ((int)blockIdx.x + i * stride) * FT + F_CW * h;
      ok = mbar_wait_short(bars + FB_ETA_FULL + grp, (uint32_t)(i >> 1) & 1u);
      if (!ok) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[F_CW];
      tmem_ld16(lane_addr + (uint32_t)grp * FT + (uint32_t)(F_CW * h), v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + FB_ETA_EMPTY + grp);
      uint32_t rh[F_CW], rl[F_CW];
      float ll = 0.0f;
      const int n_live = n - o0;  // observations of this group that exist (>= F_CW except in the last tile)
#pragma unroll
      for (int j = 0; j < F_CW; ++j) {
        const float eta = __uint_as_float(v[j]);
        // softplus(eta) = max(eta, 0) + log1p(exp(-|eta|)); sigmoid from the same exponential (MUFU ex2 / lg2 / rcp)
        const float e = fast_ex2(-1.4426950408889634f * fabsf(eta));
        const float d1 = 1.0f + e;
        const float sp = fmaxf(eta, 0.0f) + 0.6931471805599453f * fast_lg2(d1);
        const float rc = fast_rcp(d1);
        const float sg = eta >= 0.0f ? rc : e * rc;
        const bool live = j < n_live;
        ll += live ? yv[j] * eta - sp : 0.0f;
        const float r = live ? yv[j] - sg : 0.0f;
        const float hv = to_tf32(r);
        rh[j] = __float_as_uint(hv);
        rl[j] = __float_as_uint(to_tf32(r - hv));
      }
      ll_acc += (double)ll;
      if (i + 2 < nt) fetch_y(i + 2);
      if (i >= 2) {  // GEMM2 of tile i - 2 has read this column group of this R buffer
        ok = mbar_wait_short(bars + FB_R_EMPTY + F_NQ * grp + h, (uint32_t)((i - 2) >> 1) & 1u);
        if (!ok) break;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      // the residuals of a chain are consecutive K elements of GEMM2's A operand: straight into tensor memory
      tmem_st16(lane_addr + F_R_COL + (uint32_t)grp * 64u + (uint32_t)(F_CW * h), rh);
      tmem_st16(lane_addr + F_R_COL + (uint32_t)grp * 64u + 32u + (uint32_t)(F_CW * h), rl);
      tmem_st_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + FB_R_FULL + F_NQ * grp + h);
    }
    // per-chain partial of the log-likelihood (this CTA's observations, this half) and of the gradient
    LPp[((size_t)blockIdx.x * (F_EPI / 4) + ew) * Cp + c0 + row] = ok ? ll_acc : 0.0;
    if (ok && nt > 0) ok = mbar_wait_short(bars + FB_G_DONE, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* dst = GP + ((size_t)blockIdx.x * Cp + c0 + row) * Pp;
    for (int cc = 8 * ew; cc < Pp; cc += 8 * (F_EPI / 4)) {
      uint32_t v[8];
      if (ok && nt > 0) tmem_ld8(lane_addr + F_G_COL + (uint32_t)cc, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[cc + j] = (ok && nt > 0) ? __uint_as_float(v[j]) : 0.0f;
    }
    if (!ok && lane == 0 && error) *error = 5;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(F_TMEM_COLS));
}

// fixed-order reductions in double precision of the fused kernel's partials
__global__ void glm_fused_reduce_kernel(const double* __restrict__ LPp, const float* __restrict__ GP, int groups, int Cp,
                                        int Pp, int C, int P, float* __restrict__ lp, float* __restrict__ grad, int ldg, int base,
                                        const int* __restrict__ gate) {
  if (gate && *(const volatile int*)gate == 0) return;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < C) {
    double s = 0;
    for (int g = 0; g < (F_EPI / 4) * groups; ++g) s += LPp[(size_t)g * Cp + idx];
    lp[idx] += (float)s;
  }
  if (idx < C * P) {
    const int c = idx / P, p = idx - c * P;
    double s = 0;
    for (int g = 0; g < groups; ++g) s += (double)GP[((size_t)g * Cp + c) * Pp + p];
    grad[(size_t)c * ldg + base + p] += (float)s;
  }
}

}  // namespace

// test entry: one 128 x 128 tile; passes = 1 (plain TF32) or 3 (3xTF32); returns 0, a CUDA error code, or
// -62 (ETIME) when the MMA never completed
extern "C" int pm4_glm_gemm_tile(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int K, int passes,
                                 void* stream) {
  const bool ts = passes < 0;  // negative: the same product with the A operand in tensor memory
  if (ts) passes = -passes;
  if (K <= 0 || K % UMMA_K != 0 || K > 104 || passes < 1 || passes > 3) return -22;
  const size_t smem = 4 * (size_t)TILE * K * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(gemm_nt_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_ts_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int* d_err = nullptr;
  if ((e = cudaMalloc(&d_err, sizeof(int))) != cudaSuccess) return (int)e;
  cudaMemsetAsync(d_err, 0, sizeof(int), (cudaStream_t)stream);
  if (ts) gemm_ts_3xtf32_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, K, passes, d_err);
  else gemm_nt_3xtf32_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, lda, B, ldb, C, ldc, K, passes, d_err);
  int err = 0;
  e = cudaMemcpyAsync(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(d_err);
  if (e != cudaSuccess) return (int)e;
  return err ? -62 : 0;
}


// Adds the Bernoulli-logit GLM likelihood of C chains to lp[C] and grad[C, ldg] (grad may be NULL).
// X [n, P] row-major and its transpose Xt [P, ldxt] (ldxt >= n, a multiple of 4), y [n] (0/1 as float),
// theta [C, ldt] with beta at columns base .. base+P-1.  work: device scratch of pm4_glm_work_bytes(n, P, C)
// bytes.  Asynchronous on `stream`; *error_dev (device int, zeroed by the caller) becomes non-zero when a
// tensor-core step timed out.
static int glm_splits(int Cp) {
  int splits = 296 / (Cp / TILE);
  return splits < 1 ? 1 : splits;
}
static int64_t glm_np(int n) { return ((int64_t)n + KT2 - 1) / KT2 * KT2; }

extern "C" int64_t pm4_glm_work_bytes(int n, int P, int C) {
  const int64_t Cp = (C + TILE - 1) / TILE * TILE, Pp = (P + 15) / 16 * 16;
  const int64_t n_tiles = (n + TILE - 1) / TILE;
  return 4 * (glm_np(n) * Cp + n_tiles * Cp + (int64_t)glm_splits((int)Cp) * Cp * Pp) + 8 * LP_GROUPS * Cp + 256;
}

extern "C" int pm4_glm_logp_grad(const float* X, const float* Xt, int ldxt, const float* y, int n, int P,
                                 const float* theta, int ldt, int base, int C, float* lp, float* grad, int ldg,
                                 void* work, int* error_dev, void* stream) {
  if (n <= 0 || P <= 0 || P > 104 || C <= 0 || ldxt < n || ldxt % 4 != 0) return -22;
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = (C + TILE - 1) / TILE * TILE, Pp = (P + 15) / 16 * 16, K = (P + 7) & ~7;
  const int n_tiles = (n + TILE - 1) / TILE;
  const int Np = (int)glm_np(n);
  const int splits = glm_splits(Cp);
  float* Rt = (float*)work;
  float* LP = Rt + (size_t)Np * Cp;
  float* GP = LP + (size_t)n_tiles * Cp;
  double* LP2 = (double*)(((uintptr_t)(GP + (size_t)splits * Cp * Pp) + 15) & ~(uintptr_t)15);
  const size_t smem1 = 4 * (size_t)TILE * K * 4 + 16 + 4 * TILE * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(glm_eta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return (int)e;
  // one CTA per SM (213 KB of operands): two waves of 148 CTAs, each walking its share of the observation tiles
  const int eta_groups = std::max(1, std::min(n_tiles, splits));
  glm_eta_kernel<<<dim3(eta_groups, Cp / TILE), NT, smem1, st>>>(X, y, n, P, theta, ldt, base, C, Rt, Np, Cp, LP, error_dev);
  if (grad) {
    const size_t smem2 = 2 * (size_t)TILE * KT2 * 4 + 2 * (size_t)Pp * KT2 * 4 + 64;
    e = cudaFuncSetAttribute(glm_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return (int)e;
    glm_grad_kernel<<<dim3(splits, Cp / TILE), NT, smem2, st>>>(Xt, ldxt, n, P, Rt, Np, Cp, GP, Pp, splits, error_dev);
  }
  const int work_items = grad ? C * P : C;
  glm_lp_kernel<<<dim3(Cp / 32, LP_GROUPS), 256, 0, st>>>(LP, n_tiles, Cp, LP2);
  glm_reduce_kernel<<<(work_items + 255) / 256, 256, 0, st>>>(LP2, GP, splits, Cp, Pp, C, P, lp, grad, ldg, base);
  return (int)cudaGetLastError();
}

// ---- fused path: pre-split image of the constant operand + one kernel per evaluation ------------------------------
static int fused_groups(int n, int Cp) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (n + FT - 1) / FT;
  int groups = sms / (Cp / TILE);
  if (groups < 1) groups = 1;
  return groups > n_tiles ? n_tiles : groups;
}

extern "C" int64_t pm4_glm_image_bytes(int n, int P) {
  if (n <= 0 || P <= 0 || P > 104) return -1;
  const int K = (P + 7) & ~7, Pp = (P + 15) / 16 * 16;
  return (int64_t)((n + FT - 1) / FT) * (int64_t)fused_tile_bytes(K, Pp);
}

extern "C" int pm4_glm_build_image(const float* X, int n, int P, void* image, void* stream) {
  if (!X || !image || n <= 0 || P <= 0 || P > 104) return -22;
  const int K = (P + 7) & ~7, Pp = (P + 15) / 16 * 16;
  glm_image_kernel<<<(unsigned)((n + FT - 1) / FT), 256, 0, (cudaStream_t)stream>>>(X, n, P, K, Pp, (unsigned char*)image);
  return (int)cudaGetLastError();
}

extern "C" int64_t pm4_glm_fused_work_bytes(int n, int P, int C) {
  const int64_t Cp = (C + TILE - 1) / TILE * TILE, Pp = (P + 15) / 16 * 16;
  const int64_t groups = fused_groups(n, (int)Cp);
  return 8 * (F_EPI / 4) * groups * Cp + 4 * groups * Cp * Pp + 4 * groups * (Cp / TILE) + 256;
}

extern "C" int pm4_glm_fused_logp_grad(const void* image, const float* y, int n, int P, const float* theta, int ldt, int base,
                                       int C, float* lp, float* grad, int ldg, void* work, int* error_dev, const int* gate_dev,
                                       void* stream) {
  if (!image || !y || !theta || !lp || !grad || !work || n <= 0 || P <= 0 || P > 104 || C <= 0) return -22;
  cudaStream_t st = (cudaStream_t)stream;
  const int Cp = (C + TILE - 1) / TILE * TILE, Pp = (P + 15) / 16 * 16, K = (P + 7) & ~7;
  const int groups = fused_groups(n, Cp);
  double* LPp = (double*)work;
  float* GP = (float*)(LPp + (size_t)(F_EPI / 4) * groups * Cp);
  const size_t smem = (size_t)F_NSX * 2 * FT * K * 4 + (size_t)F_NSXT * 2 * Pp * FT * 4 + 8 * FB_COUNT + 16;
  cudaError_t e = cudaFuncSetAttribute(glm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // pacing of the CTAs that share tile images needs them co-resident (one CTA per SM): otherwise off
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int* pace = (int*)(GP + (size_t)groups * Cp * Pp);
  static const bool pace_off = getenv("PM4_GLM_PACE") && atoi(getenv("PM4_GLM_PACE")) == 0;
  if (groups * (Cp / TILE) > sms || Cp / TILE < 2 || pace_off) pace = nullptr;
  if (pace && (e = cudaMemsetAsync(pace, 0, sizeof(int) * (size_t)groups * (Cp / TILE), st)) != cudaSuccess) return (int)e;
  glm_fused_kernel<<<dim3(groups, Cp / TILE), F_THREADS, smem, st>>>((const unsigned char*)image, y, n, P, theta, ldt, base, C,
                                                                      LPp, GP, Cp, Pp, pace, gate_dev, error_dev);
  glm_fused_reduce_kernel<<<(C * P + 255) / 256, 256, 0, st>>>(LPp, GP, groups, Cp, Pp, C, P, lp, grad, ldg, base, gate_dev);
  return (int)cudaGetLastError();
}
