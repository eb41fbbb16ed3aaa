// fp64_peak.cu -- DFMA-saturating microbenchmark: the measured FP64-pipe peak that the ERI
// roofline fraction is quoted against (MEASURED_PEAKS.json has no FP64 entry).
// 16 independent FMA chains per thread, 8 resident warps per SM sub-partition.
#include <cuda_runtime.h>

#include <string>

#include "../../include/rchem_eri.h"

namespace rchem {
int fail_public(int code, const std::string& msg);

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the math
}
}  // namespace rchem

extern "C" int rchem_fp64_peak(int device, int repeats, double* tflops_best, double* tflops_sustained) {
  using namespace rchem;
  if (!tflops_best) return fail_public(RCHEM_ERR_INVALID_ARG, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail_public(RCHEM_ERR_NO_DEVICE, "no CUDA device");
  if (cudaSetDevice(device) != cudaSuccess) return fail_public(RCHEM_ERR_CUDA, "cudaSetDevice");
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)) != cudaSuccess)
    return fail_public(RCHEM_ERR_CUDA, "cudaMalloc");
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const double flops = 2.0 * 16.0 * (double)iters * blocks * threads;
  double best = 0.0, total_ms = 0.0;
  if (repeats < 1) repeats = 1;
  for (int r = 0; r < repeats + 2; ++r) {
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r < 2) continue;  // warm-up
    total_ms += ms;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  cudaError_t e = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (e != cudaSuccess) return fail_public(RCHEM_ERR_CUDA, cudaGetErrorString(e));
  *tflops_best = best;
  if (tflops_sustained) *tflops_sustained = flops * repeats / (total_ms * 1e-3) / 1e12;
  return RCHEM_OK;
}
