// gpu_math_check.cu -- TEST INFRASTRUCTURE (-m gpu).  Bit-for-bit comparison of the branch-free fp64 division /
// reciprocal / square root / divide-by-constant sequences of hpsim_b200/csrc/hps_math.cuh with nvcc's own
// operators on random operands covering the ranges the element maps produce (|x| in 2^-40 .. 2^40, random
// significands, plus significands close to all-zeros / all-ones).  Prints the number of mismatches per function.
//
//   gpu_math_check [log2 of the number of operand pairs, default 30]
#include <cstdio>
#include <cstdlib>

#include "../hpsim_b200/csrc/hps_kernels.cuh"

using namespace hps;

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {  // splitmix64
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double operand(unsigned long long r, unsigned long long mode) {
  unsigned long long mant = r & 0xfffffffffffffull;
  switch (mode & 7ull) {  // stress the rounding boundaries now and then
    case 0: mant &= 0xffull; break;                       // tiny significand
    case 1: mant |= 0xfffffffffff00ull; break;            // nearly all ones
    case 2: mant &= 0xfffff00000000ull; break;            // short significand
    default: break;
  }
  const long long e = 1023 - 40 + (long long)((r >> 52) % 81);
  return __longlong_as_double((long long)(((unsigned long long)e << 52) | mant));
}

__device__ __forceinline__ double uni52(unsigned long long r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

__global__ void check(unsigned long long n, unsigned long long seed, unsigned long long* bad) {
  unsigned long long local[5] = {0, 0, 0, 0, 0};
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long r1 = mix(seed + 2 * i), r2 = mix(seed + 2 * i + 1), m = mix(seed ^ i);
    const double a = (r1 >> 63 ? -1.0 : 1.0) * operand(r1, m), b = operand(r2, m >> 3);
    const double bs = (m >> 9) & 1 ? -b : b;
    if (__double_as_longlong(div_nr(a, bs)) != __double_as_longlong(a / bs)) local[0]++;
    if (__double_as_longlong(rcp_nr(bs)) != __double_as_longlong(1.0 / bs)) local[1]++;
    if (__double_as_longlong(sqrt_nr(b)) != __double_as_longlong(sqrt(b))) local[2]++;
    const double rb = 1.0 / bs;  // correctly rounded reciprocal, prepared "once"
    if (__double_as_longlong(udiv(a, bs, rb)) != __double_as_longlong(a / bs)) local[3]++;
    // fmod_twopi: arguments from 2 pi up to 2^31 (a particle tens of millions of RF periods away), dense near multiples of 2 pi
    const double tq = (double)(r1 % 2000000ull) * K_TWOPI + ((m >> 12) & 3 ? uni52(r2) * K_TWOPI : (uni52(r2) - 0.5) * 1e-9);
    const double tt = tq < K_TWOPI ? tq + K_TWOPI : tq;
    if (__double_as_longlong(fmod_twopi(tt)) != __double_as_longlong(fmod(tt, K_TWOPI))) local[4]++;
  }
  for (int k = 0; k < 5; ++k) if (local[k]) atomicAdd(&bad[k], local[k]);
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 30;
  const unsigned long long n = 1ull << lg;
  unsigned long long* bad;
  if (cudaMalloc(&bad, 5 * sizeof(unsigned long long)) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 2; }
  cudaMemset(bad, 0, 5 * sizeof(unsigned long long));
  check<<<148 * 8, 256>>>(n, 0x1234567ull, bad);
  unsigned long long h[5];
  if (cudaMemcpy(h, bad, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { fprintf(stderr, "kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 2; }
  // sqrt(0), and the reference particle's r = 0
  printf("MATHCHECK pairs %llu div %llu rcp %llu sqrt %llu udiv %llu fmod %llu\n", n, h[0], h[1], h[2], h[3], h[4]);
  return 0;
}
