// Microbenchmark: DFMA rate of an 8x8 register outer-product update, acc[t][r] += v[r]*w[t],
// the operand pattern of k_whist (three distinct 64-bit operands per DFMA, one reused).
#include <cstdio>
#include <cuda_runtime.h>
template <int TG, int FW, bool RT_ORDER>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int iters) {
  double acc[TG][FW];
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int r = 0; r < FW; ++r) acc[t][r] = 0.0;
  double v[FW], w[TG];
#pragma unroll
  for (int r = 0; r < FW; ++r) v[r] = in[threadIdx.x + r];
#pragma unroll
  for (int t = 0; t < TG; ++t) w[t] = in[threadIdx.x + 8 + t];
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (RT_ORDER) {
#pragma unroll
      for (int r = 0; r < FW; ++r)
#pragma unroll
        for (int t = 0; t < TG; ++t) acc[t][r] = fma(v[r], w[t], acc[t][r]);
    } else {
#pragma unroll
      for (int t = 0; t < TG; ++t)
#pragma unroll
        for (int r = 0; r < FW; ++r) acc[t][r] = fma(v[r], w[t], acc[t][r]);
    }
    // perturb operands so the compiler cannot hoist anything (cheap integer-free update)
#pragma unroll
    for (int r = 0; r < FW; ++r) v[r] = -v[r];
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TG; ++t)
#pragma unroll
    for (int r = 0; r < FW; ++r) s += acc[t][r];
  if (s == 12345.678) out[0] = s;
}
template <int TG, int FW, bool RT>
void run(const char* name, double* d, double* in, int sms, int clk) {
  const int iters = 4000;
  for (int threads : {128, 256}) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<TG, FW, RT><<<sms, threads>>>(d, in, 10);
    cudaEventRecord(e0);
    k<TG, FW, RT><<<sms, threads>>>(d, in, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * threads * TG * FW * iters;
    printf("%s threads/SM %4d: %.3f ms %.2f TFMA/s %.1f DFMA/clk/SM\n", name, threads, ms, fma / ms / 1e9,
           fma / (ms * 1e-3) / sms / (clk * 1e3));
  }
}
int main() {
  double *d, *in; cudaMalloc(&d, 8); cudaMalloc(&in, 8 * 4096); cudaMemset(in, 0, 8 * 4096);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  run<8, 8, true>("8x8 r-outer(v reused)", d, in, p.multiProcessorCount, p.clockRate);
  run<8, 8, false>("8x8 t-outer(w reused)", d, in, p.multiProcessorCount, p.clockRate);
  run<4, 8, true>("4x8 r-outer", d, in, p.multiProcessorCount, p.clockRate);
  run<8, 4, true>("8x4 r-outer", d, in, p.multiProcessorCount, p.clockRate);
  return 0;
}
