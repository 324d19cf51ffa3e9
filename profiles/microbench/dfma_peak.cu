// Microbenchmark: achievable DFMA rate on this GPU (register-resident accumulators,
// no memory traffic), for warps/SM sweeps.  Build: nvcc -arch=sm_100a -O3 dfma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ACC>
__global__ void k(double* out, int iters, double a, double b) {
  double acc[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}
int main() {
  double* d; cudaMalloc(&d, 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int iters = 20000;
  for (int threads : {32, 64, 128, 256, 512, 1024}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<32><<<p.multiProcessorCount, threads>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<32><<<p.multiProcessorCount, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)p.multiProcessorCount * threads * 32.0 * iters;
    printf("threads/SM %4d: %.3f ms  %.2f TFMA/s  (%.2f TFLOP/s)  %.1f DFMA/clk/SM at %d MHz\n", threads, ms, fma / ms / 1e9,
           2 * fma / ms / 1e9, fma / (ms * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000);
  }
  return 0;
}
