// Microbenchmark (B200): the fp64 issue rate that bounds k_segment.  Measures, per SM and clock,
//   (a) DFMA warp-instructions with 8 independent chains per thread and 32 warps / SM (the pipe's peak),
//   (b) the same stream with one integer instruction per DFMA (do the "other" instructions of the sweep issue for free?),
//   (c) one dependent DFMA chain per warp at 4 warps / scheduler (latency of a dependent fp64 operation),
//   (d) k_segment's mix: ~48 % fp64, 16 warps / SM, two chains per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS, int INTS>
__global__ void k(double* out, int iters, double a, double b, int m) {
  double x[CHAINS];
  int q[INTS > 0 ? INTS : 1];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-9 + c;
#pragma unroll
  for (int c = 0; c < (INTS > 0 ? INTS : 1); ++c) q[c] = threadIdx.x + c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
#pragma unroll
      for (int c = 0; c < INTS; ++c) q[c] = q[c] * m + i;
    }
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  int t = 0;
#pragma unroll
  for (int c = 0; c < (INTS > 0 ? INTS : 1); ++c) t += q[c];
  if (s == 12345.678 && t == 77) out[0] = s;
}
template <int CHAINS, int INTS>
static void run(const char* what, int blocks_per_sm, int threads, int sms, double mhz) {
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<CHAINS, INTS><<<sms * blocks_per_sm, threads>>>(d, 64, 1.0000001, 1e-9, 3);
  cudaEventRecord(e0);
  k<CHAINS, INTS><<<sms * blocks_per_sm, threads>>>(d, iters, 1.0000001, 1e-9, 3);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)blocks_per_sm * threads / 32;
  const double dfma = warps * iters * 8.0 * CHAINS, ints = warps * iters * 8.0 * INTS;
  const double cycles = ms * 1e-3 * mhz * 1e6;
  printf("%-64s %7.3f ms  DFMA/clk/SM %.3f  int/clk/SM %.3f  (fp64 pipe %.1f %% of 2 warp-instr/clk/SM)\n", what, ms, dfma / cycles, ints / cycles,
         100.0 * dfma / cycles / 2.0);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1e3;
  printf("%s, %d SMs, %.0f MHz (nominal; the rates below assume it)\n", p.name, p.multiProcessorCount, mhz);
  const int sms = p.multiProcessorCount;
  run<8, 0>("(a) 8 chains/thread, 32 warps/SM, DFMA only", 4, 256, sms, mhz);
  run<8, 8>("(b) 8 DFMA chains + 8 IMAD chains, 32 warps/SM", 4, 256, sms, mhz);
  run<8, 4>("(b') 8 DFMA chains + 4 IMAD chains, 32 warps/SM", 4, 256, sms, mhz);
  run<1, 0>("(c) 1 dependent chain/thread, 16 warps/SM", 2, 256, sms, mhz);
  run<2, 0>("(c') 2 chains/thread, 16 warps/SM", 2, 256, sms, mhz);
  run<2, 2>("(d) 2 DFMA + 2 IMAD chains/thread, 16 warps/SM", 2, 256, sms, mhz);
  run<1, 1>("(d') 1 DFMA + 1 IMAD chain/thread, 24 warps/SM", 3, 256, sms, mhz);
  run<4, 4>("(d'') 4 DFMA + 4 IMAD chains/thread, 16 warps/SM", 2, 256, sms, mhz);
  return 0;
}
