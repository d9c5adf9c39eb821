// pm4_gp.cu -- the marginal likelihood of GP regression for all chains at once (PM4_T_MVN_EXPQUAD).
//
// Reference path: ExpQuad.__call__ (/root/reference/pymc4/gp/cov.py:227-272, 463-469: tfp
// ExponentiatedQuadratic.matrix), stabilize (gp/util.py:15-21), MvNormal.log_prob
// (distributions/multivariate.py:159-199: tfd.MultivariateNormalFullCovariance) and their reverse-mode
// gradient, evaluated per chain under tf.vectorized_map: K [n, n] per chain, a Cholesky factorisation and its
// adjoint per leapfrog (~n^3 flops per gradient evaluation).
//
// Here one batched evaluation is, per chain c (grid z / y = chain):
//   K_c   = a^2 exp(-|X_i - X_j|^2 / (2 l^2)) + (noise + jitter) I          gp_build_kernel
//   L_c   = chol(K_c)            left-looking, 64-wide block columns:        gp_nt_kernel<CHOL> + gp_potrf_kernel + gp_trsm_kernel
//   Zt_c  = L_c^-T               block forward substitution (same kernels):  gp_nt_kernel<ZT> + gp_trsm_kernel
//   v = L^-1 y, alpha = K^-1 y                                               gp_v_kernel, gp_alpha_kernel
//   S     = sum_ij (alpha_i alpha_j - (Zt Zt^T)_ij) {K_se, K_se d^2, I}_ij   gp_nt_kernel<DOT>  (K^-1 is never stored)
//   lp    = -1/2 |v|^2 - sum log L_ii - n/2 log 2 pi,  d lp / d(l, a, noise) from S       gp_finish_kernel
// All of it in float64 (B200: 64 FP64 FMA/clk/SM, through DFMA or the FP64 tensor-core MMA), whatever the sampling dtype: the kernel matrix
// of a smooth covariance function has a condition number ~ a^2 n / noise, out of reach of float32 or of
// split-TF32 tensor-core arithmetic.  The O(n^3) work is three passes of ONE tile kernel, C[I, J] -= sum_k
// A[I, k] B[J, k] on 64 x 64 tiles of FP64 tensor-core MMAs (mma.sync m8n8k4): n^3/3 flops each for the factor, the inverse
// factor and the contraction with dK.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/pm4b200.h"
#include "pm4_philox.cuh"

namespace {

constexpr int NB = 64;   // block size of the factorisation
constexpr int KS = 16;   // k-slab of the tile kernel

struct GpShape {
  int n, np, nt, d, C;             // points, padded points (multiple of NB), tiles per side, features, chains in this chunk
  const double* X;                 // [n, d]
  const double* y;                 // [n]
  double* Kb;                      // [C, np, np]  kernel matrix -> Cholesky factor (lower)
  double* Zt;                      // [C, np, np]  L^-T (upper)
  double* hyp;                     // [C, 8]  l, a, v, diag, dl/dz, da/dz, dv/dz, logdet
  double* vec;                     // [C, 2, np]  v = L^-1 y, alpha = K^-1 y
  double* part;                    // [C, nt (nt + 1) / 2, 3]  per-tile partial sums
};

struct GpParams {  // how the three scalars are read from theta
  int elem[3];
  int tr[3];
  double shift[3], scale[3], cst[3];
  int noise_squared;
  double jitter;
};

template <typename real>
__global__ void gp_hyp_kernel(GpShape s, GpParams p, const real* __restrict__ theta, int ldt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= s.C) return;
  double x[3], dx[3];
  for (int k = 0; k < 3; ++k) {
    if (p.elem[k] < 0) { x[k] = p.cst[k]; dx[k] = 0; continue; }
    const double z = (double)theta[(size_t)c * ldt + p.elem[k]];
    if (p.tr[k] == PM4_TR_EXP) { x[k] = p.shift[k] + p.scale[k] * exp(z); dx[k] = x[k] - p.shift[k]; }
    else if (p.tr[k] == PM4_TR_SIGMOID) { const double sg = 1.0 / (1.0 + exp(-z)); x[k] = p.shift[k] + p.scale[k] * sg; dx[k] = p.scale[k] * sg * (1.0 - sg); }
    else { x[k] = z; dx[k] = 1.0; }
  }
  double* h = s.hyp + (size_t)c * 8;
  h[0] = x[0]; h[1] = x[1]; h[2] = x[2];
  h[3] = (p.noise_squared ? x[2] * x[2] : x[2]) + p.jitter;
  h[4] = dx[0]; h[5] = dx[1]; h[6] = dx[2];
  h[7] = 0.0;
}

__device__ __forceinline__ double dist2(const double* __restrict__ X, int d, int i, int j) {
  double d2 = 0;
  for (int f = 0; f < d; ++f) { const double df = X[(size_t)i * d + f] - X[(size_t)j * d + f]; d2 += df * df; }
  return d2;
}

// lower tiles of K (diagonal tiles in full); rows/columns >= n are padded with the identity
__global__ void __launch_bounds__(256) gp_build_kernel(GpShape s) {
  const int tj = blockIdx.x, ti = blockIdx.y, c = blockIdx.z;
  if (tj > ti) return;
  const double* h = s.hyp + (size_t)c * 8;
  const double a2 = h[1] * h[1], inv = -0.5 / (h[0] * h[0]), dg = h[3];
  double* K = s.Kb + (size_t)c * s.np * s.np;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = tj * NB + tx;
  for (int r = ty; r < NB; r += 4) {
    const int i = ti * NB + r;
    double v;
    if (i < s.n && j < s.n) v = a2 * exp(inv * dist2(s.X, s.d, i, j)) + (i == j ? dg : 0.0);
    else v = (i == j) ? 1.0 : 0.0;
    K[(size_t)i * s.np + j] = v;
  }
}

// ------------------------------------------------------------------------------------------
// the tile kernel: acc[I, J] = sum_{k in [k0, k1)} A[I, k] B[J, k], 64 x 64 per block, 4 warps x (4 x 4 MMA tiles)
// ------------------------------------------------------------------------------------------
enum { MODE_CHOL = 0, MODE_ZT = 1, MODE_DOT = 2 };

// 128 threads (4 warps) per 64 x 64 tile; warp (wr, wc) owns the 32 x 32 quadrant as 4 x 4 FP64 tensor-core tiles
// (mma.sync m8n8k4: the form cuBLAS DGEMM uses on this part).  Per k4 step a thread reads 8 doubles from shared
// memory for 16 MMAs = 4096 FMAs per warp: a quarter of the shared-memory traffic per flop of an 8 x 8 register
// tile on the DFMA pipe, which is what limited the previous version (LSU data pipe 69 % busy at 44 % FP64).
// Fragments (PTX ISA, mma.m8n8k4.f64): A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4],
// C[row = lane / 4][col = 2 (lane % 4) + {0, 1}].  Slabs are stored k-major with a row stride of 68 doubles:
// the 16 lanes of a half-warp then hit 16 distinct 8-byte banks.  The next slab is fetched into registers while
// the current one is multiplied.
constexpr int NT_THREADS = 128;
constexpr int SLD = NB + 4;  // slab row stride in doubles

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// element (mi, ni, e) of a thread's accumulators inside the 64 x 64 tile
__device__ __forceinline__ int tile_row(int wr, int lane, int mi) { return wr * 32 + mi * 8 + (lane >> 2); }
__device__ __forceinline__ int tile_col(int wc, int lane, int ni) { return wc * 32 + ni * 8 + (lane & 3) * 2; }

__device__ __forceinline__ void tile_nt(const double* __restrict__ A, const double* __restrict__ B, int ld, int k0, int k1,
                                        double (&acc)[4][4][2], double (*As)[SLD], double (*Bs)[SLD]) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, wr = warp >> 1, wc = warp & 1;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
  if (k0 >= k1) return;
  const int sr = t & 63, sk = (t >> 6) * 8;  // this thread stages row sr, slab columns sk .. sk + 7 of both slabs
  const double* ap = A + (size_t)sr * ld + sk;
  const double* bp = B + (size_t)sr * ld + sk;
  double4 a0 = *reinterpret_cast<const double4*>(ap + k0), a1 = *reinterpret_cast<const double4*>(ap + k0 + 4);
  double4 b0 = *reinterpret_cast<const double4*>(bp + k0), b1 = *reinterpret_cast<const double4*>(bp + k0 + 4);
  const int fr = lane >> 2, fk = lane & 3;
  for (int k = k0; k < k1; k += KS) {
    __syncthreads();
    As[sk + 0][sr] = a0.x; As[sk + 1][sr] = a0.y; As[sk + 2][sr] = a0.z; As[sk + 3][sr] = a0.w;
    As[sk + 4][sr] = a1.x; As[sk + 5][sr] = a1.y; As[sk + 6][sr] = a1.z; As[sk + 7][sr] = a1.w;
    Bs[sk + 0][sr] = b0.x; Bs[sk + 1][sr] = b0.y; Bs[sk + 2][sr] = b0.z; Bs[sk + 3][sr] = b0.w;
    Bs[sk + 4][sr] = b1.x; Bs[sk + 5][sr] = b1.y; Bs[sk + 6][sr] = b1.z; Bs[sk + 7][sr] = b1.w;
    __syncthreads();
    if (k + KS < k1) {
      a0 = *reinterpret_cast<const double4*>(ap + k + KS); a1 = *reinterpret_cast<const double4*>(ap + k + KS + 4);
      b0 = *reinterpret_cast<const double4*>(bp + k + KS); b1 = *reinterpret_cast<const double4*>(bp + k + KS + 4);
    }
#pragma unroll
    for (int k4 = 0; k4 < KS; k4 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        a[h] = As[k4 + fk][wr * 32 + h * 8 + fr];
        b[h] = Bs[k4 + fk][wc * 32 + h * 8 + fr];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
    }
  }
}

// MODE_CHOL, block column jb: tile (I = jb + blockIdx.x, jb) of L -= L[I, 0 : jb NB] L[jb, 0 : jb NB]^T
// MODE_ZT,   block ib:        tile (J = blockIdx.x <= ib, ib) of Zt = [J == ib] I - Zt[J, J NB : ib NB] L[ib, J NB : ib NB]^T
// MODE_DOT:  tile (I, J <= I) from the linear index blockIdx.x: G = Zt[I, I NB :] Zt[J, I NB :]^T, then the sums
template <int MODE>
__global__ void __launch_bounds__(NT_THREADS) gp_nt_kernel(GpShape s, int blk) {
  __shared__ __align__(16) double As[KS][SLD];
  __shared__ __align__(16) double Bs[KS][SLD];
  const int c = blockIdx.y, np = s.np;
  double* L = s.Kb + (size_t)c * np * np;
  double* Zt = s.Zt + (size_t)c * np * np;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5, wr = warp >> 1, wc = warp & 1;
  double acc[4][4][2];
  if (MODE == MODE_CHOL) {
    const int I = blk + blockIdx.x;
    tile_nt(L + (size_t)I * NB * np, L + (size_t)blk * NB * np, np, 0, blk * NB, acc, As, Bs);
    double* Ct = L + (size_t)I * NB * np + (size_t)blk * NB;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double2* p = reinterpret_cast<double2*>(Ct + (size_t)tile_row(wr, lane, i) * np + tile_col(wc, lane, j));
        double2 v = *p;
        v.x -= acc[i][j][0]; v.y -= acc[i][j][1];
        *p = v;
      }
  } else if (MODE == MODE_ZT) {
    const int J = blockIdx.x;
    tile_nt(Zt + (size_t)J * NB * np, L + (size_t)blk * NB * np, np, J * NB, blk * NB, acc, As, Bs);
    double* Ct = Zt + (size_t)J * NB * np + (size_t)blk * NB;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tile_row(wr, lane, i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = tile_col(wc, lane, j);
        double2 v;
        v.x = ((J == blk && r == col) ? 1.0 : 0.0) - acc[i][j][0];
        v.y = ((J == blk && r == col + 1) ? 1.0 : 0.0) - acc[i][j][1];
        *reinterpret_cast<double2*>(Ct + (size_t)r * np + col) = v;
      }
    }
  } else {
    // linear index -> (I, J <= I)
    int I = (int)((sqrt(8.0 * (double)blockIdx.x + 1.0) - 1.0) * 0.5);
    while ((I + 1) * (I + 2) / 2 <= (int)blockIdx.x) ++I;
    while (I * (I + 1) / 2 > (int)blockIdx.x) --I;
    const int J = (int)blockIdx.x - I * (I + 1) / 2;
    tile_nt(Zt + (size_t)I * NB * np, Zt + (size_t)J * NB * np, np, I * NB, np, acc, As, Bs);
    const double* h = s.hyp + (size_t)c * 8;
    const double a2 = h[1] * h[1], inv = -0.5 / (h[0] * h[0]);
    const double* alpha = s.vec + ((size_t)c * 2 + 1) * np;
    const double wt = (I == J) ? 1.0 : 2.0;
    double s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gi = I * NB + tile_row(wr, lane, i);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gj = J * NB + tile_col(wc, lane, j) + e;
          if (gi < s.n && gj < s.n) {
            const double W = alpha[gi] * alpha[gj] - acc[i][j][e];
            const double d2 = dist2(s.X, s.d, gi, gj);
            const double kse = a2 * exp(inv * d2);
            s1 += wt * W * kse;
            s2 += wt * W * kse * d2;
            if (gi == gj) s3 += W;
          }
        }
    }
    // block reduction in a fixed order
    __shared__ double red[3][NT_THREADS];
    __syncthreads();
    red[0][t] = s1; red[1][t] = s2; red[2][t] = s3;
    __syncthreads();
    for (int o = NT_THREADS / 2; o > 0; o >>= 1) {
      if (t < o) { red[0][t] += red[0][t + o]; red[1][t] += red[1][t + o]; red[2][t] += red[2][t + o]; }
      __syncthreads();
    }
    if (t < 3) s.part[((size_t)c * (s.nt * (s.nt + 1) / 2) + blockIdx.x) * 3 + t] = red[t][0];
  }
}

// Cholesky of the diagonal tile (jb, jb) in shared memory; adds sum log diag to hyp[7]
__global__ void __launch_bounds__(NB) gp_potrf_kernel(GpShape s, int jb) {
  __shared__ double S[NB][NB + 1];
  const int c = blockIdx.x, np = s.np, r = threadIdx.x;
  double* T = s.Kb + (size_t)c * np * np + (size_t)jb * NB * np + (size_t)jb * NB;
  for (int k = 0; k < NB; ++k) S[k][r] = T[(size_t)k * np + r];
  __syncthreads();
  for (int k = 0; k < NB; ++k) {
    const double dkk = sqrt(S[k][k]);  // NaN when the matrix is not positive definite: propagates to lp
    __syncthreads();
    if (r == k) S[k][k] = dkk;
    if (r > k) S[r][k] = S[r][k] / dkk;
    __syncthreads();
    if (r > k)
      for (int q = k + 1; q <= r; ++q) S[r][q] -= S[r][k] * S[q][k];
    __syncthreads();
  }
  for (int k = 0; k < NB; ++k) T[(size_t)k * np + r] = (r <= k) ? S[k][r] : 0.0;
  if (r == 0) {
    double ld = 0;
    for (int k = 0; k < NB; ++k) ld += log(S[k][k]);
    s.hyp[(size_t)c * 8 + 7] += ld;
  }
}

// rows [r0, r1) of M, column block jb:  x <- x L11^-T  with L11 = the diagonal tile (jb, jb) of L (one thread per row)
__global__ void __launch_bounds__(128) gp_trsm_kernel(GpShape s, int jb, int use_zt, int r0, int r1) {
  __shared__ double Ls[NB][NB + 1];
  __shared__ double rd[NB];
  const int c = blockIdx.y, np = s.np;
  const double* L11 = s.Kb + (size_t)c * np * np + (size_t)jb * NB * np + (size_t)jb * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Ls[e / NB][e % NB] = L11[(size_t)(e / NB) * np + (e % NB)];
  __syncthreads();
  if (threadIdx.x < NB) rd[threadIdx.x] = 1.0 / Ls[threadIdx.x][threadIdx.x];
  __syncthreads();
  const int row = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= r1) return;
  double* M = (use_zt ? s.Zt : s.Kb) + (size_t)c * np * np + (size_t)row * np + (size_t)jb * NB;
  double x[NB];
#pragma unroll
  for (int q = 0; q < NB; q += 4) {
    const double4 v = *reinterpret_cast<const double4*>(M + q);
    x[q] = v.x; x[q + 1] = v.y; x[q + 2] = v.z; x[q + 3] = v.w;
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    double tv = x[q];
#pragma unroll
    for (int k = 0; k < q; ++k) tv = fma(-x[k], Ls[q][k], tv);
    x[q] = tv * rd[q];
  }
#pragma unroll
  for (int q = 0; q < NB; q += 4) *reinterpret_cast<double4*>(M + q) = make_double4(x[q], x[q + 1], x[q + 2], x[q + 3]);
}

// v[k] = sum_{i <= k} Zt[i, k] y[i]   (v = L^-1 y)
__global__ void __launch_bounds__(256) gp_v_kernel(GpShape s) {
  const int c = blockIdx.y, np = s.np, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= np) return;
  const double* Zt = s.Zt + (size_t)c * np * np;
  double acc = 0;
  const int top = k < s.n ? k : s.n - 1;
  for (int i = 0; i <= top; ++i) acc = fma(Zt[(size_t)i * np + k], s.y[i], acc);
  s.vec[(size_t)c * 2 * np + k] = acc;
}

// alpha[i] = sum_{k >= i} Zt[i, k] v[k]   (alpha = L^-T v = K^-1 y), one warp per row
__global__ void __launch_bounds__(256) gp_alpha_kernel(GpShape s) {
  const int c = blockIdx.y, np = s.np, i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= np) return;
  const double* Zr = s.Zt + (size_t)c * np * np + (size_t)i * np;
  const double* v = s.vec + (size_t)c * 2 * np;
  double acc = 0;
  for (int k = (i & ~31) + lane; k < np; k += 32)
    if (k >= i) acc = fma(Zr[k], v[k], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s.vec[((size_t)c * 2 + 1) * np + i] = acc;
}

// lp and the gradient with respect to the unconstrained scalars (one warp per chain, fixed summation order)
template <typename real>
__global__ void __launch_bounds__(256) gp_finish_kernel(GpShape s, GpParams p, real* lp, int lp_stride, int accumulate,
                                                        real* grad, int ldg) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= s.C) return;
  const double* v = s.vec + (size_t)c * 2 * s.np;
  double quad = 0;
  for (int k = lane; k < s.n; k += 32) quad = fma(v[k], v[k], quad);
  const int ntile = s.nt * (s.nt + 1) / 2;
  const double* part = s.part + (size_t)c * ntile * 3;
  double s1 = 0, s2 = 0, s3 = 0;
  if (grad)
    for (int k = lane; k < ntile; k += 32) { s1 += part[k * 3]; s2 += part[k * 3 + 1]; s3 += part[k * 3 + 2]; }
  for (int o = 16; o > 0; o >>= 1) {
    quad += __shfl_xor_sync(0xffffffffu, quad, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s3 += __shfl_xor_sync(0xffffffffu, s3, o);
  }
  if (lane != 0) return;
  const double* h = s.hyp + (size_t)c * 8;
  const double val = -0.5 * quad - h[7] - 0.5 * (double)s.n * 1.8378770664093454835606594728112;
  real* out = lp + (size_t)c * lp_stride;
  *out = accumulate ? (real)((double)*out + val) : (real)val;
  if (grad) {
    const double l = h[0], a = h[1], nv = h[2];
    const double g[3] = {0.5 * s2 / (l * l * l), s1 / a, 0.5 * s3 * (p.noise_squared ? 2.0 * nv : 1.0)};
    for (int k = 0; k < 3; ++k)
      if (p.elem[k] >= 0) {
        real* gp = grad + (size_t)c * ldg + p.elem[k];
        *gp = (real)((double)*gp + g[k] * h[4 + k]);
      }
  }
}

// forward sampling: z ~ N(0, I) from the Philox stream of (state, element), then y = loc + L z
__global__ void __launch_bounds__(256) gp_z_kernel(GpShape s, long long state0, uint32_t elem0, uint32_t k0, uint32_t k1) {
  const int c = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.np) return;
  double z = 0.0;
  if (i < s.n) {
    const unsigned long long m = (unsigned long long)(state0 + c);
    const pm4::U4 u = pm4::philox4x32_10((uint32_t)m, (uint32_t)(m >> 32), elem0 + (uint32_t)i, pm4::RNG_PREDICTIVE, k0, k1);
    double n1;
    pm4::box_muller<double>(u.x, u.y, z, n1);
  }
  s.vec[(size_t)c * 2 * s.np + i] = z;
}

template <typename real>
__global__ void __launch_bounds__(256) gp_draw_kernel(GpShape s, const double* __restrict__ loc, real* __restrict__ out, int ldo) {
  const int c = blockIdx.y, i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= s.n) return;
  const double* Lr = s.Kb + (size_t)c * s.np * s.np + (size_t)i * s.np;
  const double* z = s.vec + (size_t)c * 2 * s.np;
  double acc = 0;
  for (int k = lane; k <= i; k += 32) acc = fma(Lr[k], z[k], acc);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[(size_t)c * ldo + i] = (real)(loc[i] + acc);
}

}  // namespace

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// workspace layout + how the scalars are read; hyper-parameters, kernel matrix and its Cholesky factor for C chains
static void gp_setup(GpShape& s, GpParams& p, const double* X, const double* y, int n, int d, const double* par,
                     const int* tr, const double* shift, const double* scale, int C, void* work) {
  s.n = n; s.np = round_up(n, NB); s.nt = s.np / NB; s.d = d; s.C = C; s.X = X; s.y = y;
  const size_t np = (size_t)s.np;
  double* w = (double*)work;
  s.Kb = w; w += (size_t)C * np * np;
  s.Zt = w; w += (size_t)C * np * np;
  s.hyp = w; w += (size_t)C * 8;
  s.vec = w; w += (size_t)C * 2 * np;
  s.part = w;
  for (int k = 0; k < 3; ++k) {
    p.elem[k] = (int)par[k];
    p.cst[k] = par[3 + k];
    p.tr[k] = tr ? tr[k] : PM4_TR_IDENTITY;
    p.shift[k] = shift ? shift[k] : 0.0;
    p.scale[k] = scale ? scale[k] : 1.0;
  }
  p.noise_squared = par[6] != 0.0;
  p.jitter = par[7];
}

static void gp_factorize(const GpShape& s, const GpParams& p, int dtype, const void* theta, int ldt, cudaStream_t st) {
  const int C = s.C;
  const unsigned cb = (unsigned)((C + 127) / 128);
  if (dtype == PM4_F64) gp_hyp_kernel<double><<<cb, 128, 0, st>>>(s, p, (const double*)theta, ldt);
  else gp_hyp_kernel<float><<<cb, 128, 0, st>>>(s, p, (const float*)theta, ldt);
  gp_build_kernel<<<dim3((unsigned)s.nt, (unsigned)s.nt, (unsigned)C), 256, 0, st>>>(s);
  // Cholesky, left-looking over block columns
  for (int jb = 0; jb < s.nt; ++jb) {
    if (jb > 0) gp_nt_kernel<MODE_CHOL><<<dim3((unsigned)(s.nt - jb), (unsigned)C), NT_THREADS, 0, st>>>(s, jb);
    gp_potrf_kernel<<<(unsigned)C, NB, 0, st>>>(s, jb);
    const int r0 = (jb + 1) * NB;
    if (r0 < s.np) gp_trsm_kernel<<<dim3((unsigned)((s.np - r0 + 127) / 128), (unsigned)C), 128, 0, st>>>(s, jb, 0, r0, s.np);
  }
}

extern "C" int64_t pm4_gp_work_bytes(int n, int C) {
  if (n <= 0 || C <= 0) return PM4_EINVAL;
  const size_t np = (size_t)round_up(n, NB), nt = np / NB;
  return (int64_t)((size_t)C * (2 * np * np + 8 + 2 * np + nt * (nt + 1) / 2 * 3) * sizeof(double));
}

// One batched evaluation for C chains.  X [n, d], y [n] (y - loc) double, device.  par: the 8-double record of
// include/pm4b200.h (PM4_T_MVN_EXPQUAD) on the HOST; tr/shift/scale: the transform of the free variable behind each
// of the three scalars (ignored where par[k] < 0).  theta [C, ldt] unconstrained; lp[c * lp_stride] (+)= the term;
// grad [C, ldg] += its gradient (NULL: value only).  work: pm4_gp_work_bytes(n, C) bytes.  dtype: PM4_F32 / PM4_F64
// is the type of theta, lp and grad; the factorisation is float64 either way.  Returns 0 or a cudaError_t.
extern "C" int pm4_gp_logp_grad(const double* X, const double* y, int n, int d, const double* par, const int* tr,
                                const double* shift, const double* scale, int dtype, const void* theta, int ldt, int C,
                                void* lp, int lp_stride, int accumulate, void* grad, int ldg, void* work, void* stream) {
  if (!X || !y || !par || !theta || !lp || !work || n <= 0 || d <= 0 || C <= 0) return PM4_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  GpShape s;
  GpParams p;
  gp_setup(s, p, X, y, n, d, par, tr, shift, scale, C, work);
  const size_t ntile = (size_t)s.nt * (s.nt + 1) / 2;
  gp_factorize(s, p, dtype, theta, ldt, st);
  // Zt = L^-T, block column by block column
  for (int ib = 0; ib < s.nt; ++ib) {
    gp_nt_kernel<MODE_ZT><<<dim3((unsigned)(ib + 1), (unsigned)C), NT_THREADS, 0, st>>>(s, ib);
    const int r1 = (ib + 1) * NB;
    gp_trsm_kernel<<<dim3((unsigned)((r1 + 127) / 128), (unsigned)C), 128, 0, st>>>(s, ib, 1, 0, r1);
  }
  gp_v_kernel<<<dim3((unsigned)((s.np + 255) / 256), (unsigned)C), 256, 0, st>>>(s);
  if (grad) {
    gp_alpha_kernel<<<dim3((unsigned)((s.np + 7) / 8), (unsigned)C), 256, 0, st>>>(s);
    gp_nt_kernel<MODE_DOT><<<dim3((unsigned)ntile, (unsigned)C), NT_THREADS, 0, st>>>(s, 0);
  }
  const unsigned fb = (unsigned)((C + 7) / 8);
  if (dtype == PM4_F64) gp_finish_kernel<double><<<fb, 256, 0, st>>>(s, p, (double*)lp, lp_stride, accumulate, (double*)grad, ldg);
  else gp_finish_kernel<float><<<fb, 256, 0, st>>>(s, p, (float*)lp, lp_stride, accumulate, (float*)grad, ldg);
  return (int)cudaGetLastError();
}

// Forward sampling of the observed vector (sample_posterior_predictive / sample_prior_predictive for an MvNormal with
// the ExpQuad + noise covariance, /root/reference/pymc4/forward_sampling.py:26-427): for each of the C states
// out[c, :] = loc + chol(K_c) z_c with z_c ~ N(0, I) drawn from the Philox stream of (state0 + c, elem0 + i, seed) -- the
// oracle draws the identical stream.  loc [n] device doubles; out [C, ldo] of `dtype`; the other arguments as above.
extern "C" int pm4_gp_sample(const double* X, const double* loc, int n, int d, const double* par, const int* tr,
                             const double* shift, const double* scale, int dtype, const void* theta, int ldt, int C,
                             int64_t state0, uint32_t elem0, uint64_t seed, void* out, int ldo, void* work, void* stream) {
  if (!X || !loc || !par || !theta || !out || !work || n <= 0 || d <= 0 || C <= 0) return PM4_EINVAL;
  cudaStream_t st = (cudaStream_t)stream;
  GpShape s;
  GpParams p;
  gp_setup(s, p, X, loc, n, d, par, tr, shift, scale, C, work);
  gp_factorize(s, p, dtype, theta, ldt, st);
  gp_z_kernel<<<dim3((unsigned)((s.np + 255) / 256), (unsigned)C), 256, 0, st>>>(s, (long long)state0, elem0, (uint32_t)seed,
                                                                                 (uint32_t)(seed >> 32));
  const dim3 g((unsigned)((n + 7) / 8), (unsigned)C);
  if (dtype == PM4_F64) gp_draw_kernel<double><<<g, 256, 0, st>>>(s, loc, (double*)out, ldo);
  else gp_draw_kernel<float><<<g, 256, 0, st>>>(s, loc, (float*)out, ldo);
  return (int)cudaGetLastError();
}

// number of kernel launches of one pm4_gp_logp_grad call (bench.py counts launches)
extern "C" int pm4_gp_launches(int n, int want_grad) {
  const int nt = round_up(n, NB) / NB;
  return 2 + (nt - 1) + nt + (nt - 1) + 2 * nt + 1 + (want_grad ? 2 : 0) + 1;
}
