// pm4_peaks.cu -- on-chip roofline denominators for the register-resident sampler kernels.
// MEASURED_PEAKS.json (driver-written) holds only the HBM copy bandwidth and the bf16 tensor
// throughput; the NUTS kernel of the radon-shaped model is bound by the FP32 pipe and by
// shared-memory reads, so bench.py measures those two on the box it runs on:
//   pm4_peak_fp32_tflops : dependent-free FFMA streams, 8 accumulators per thread
//   pm4_peak_smem_tbs    : conflict-free LDS.128 from shared memory
// Both time with CUDA events on the given stream after a warm-up launch.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(1024) fma_kernel(float* out, int iters, float a, float b) {
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = fmaf(acc[k], a, b);
  }
  float s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += acc[k];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the loop alive
}

__global__ void __launch_bounds__(1024) lds_kernel(float* out, int iters) {
  extern __shared__ float4 sm4[];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm4[i] = make_float4(i, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = sm4[(idx + k * 512) & 4095];  // 8 distinct rows: no load is redundant
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + 32) & 4095;
  }
  if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x;
}

template <class F> int timed(cudaStream_t st, F launch, float* ms) {
  cudaEvent_t e0, e1;
  if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1;
  launch();  // warm-up
  cudaEventRecord(e0, st);
  launch();
  cudaEventRecord(e1, st);
  cudaError_t e = cudaEventSynchronize(e1);
  if (e == cudaSuccess) e = cudaEventElapsedTime(ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (int)e;
}

}  // namespace

extern "C" int pm4_peak_fp32_tflops(int blocks_per_sm, int iters, void* stream, double* tflops) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out = nullptr;
  const int grid = sms * blocks_per_sm;
  if (cudaMalloc(&out, sizeof(float) * (size_t)grid * 1024) != cudaSuccess) return -1;
  float ms = 0;
  int rc = timed(st, [&] { fma_kernel<<<grid, 1024, 0, st>>>(out, iters, 1.000001f, 1e-7f); }, &ms);
  cudaFree(out);
  if (rc) return rc;
  *tflops = 2.0 * 8.0 * (double)iters * 1024.0 * (double)grid / (ms * 1e-3) / 1e12;
  return 0;
}

extern "C" int pm4_peak_smem_tbs(int iters, void* stream, double* tbs) {
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  float* out = nullptr;
  const int grid = sms * 2;
  if (cudaMalloc(&out, sizeof(float) * (size_t)grid * 1024) != cudaSuccess) return -1;
  cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  float ms = 0;
  int rc = timed(st, [&] { lds_kernel<<<grid, 1024, 65536, st>>>(out, iters); }, &ms);
  cudaFree(out);
  if (rc) return rc;
  *tbs = 16.0 * 8.0 * (double)iters * 1024.0 * (double)grid / (ms * 1e-3) / 1e12;
  return 0;
}
