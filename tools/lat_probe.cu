// Microbenchmark: dependent-issue latencies (cycles) that bound the single-CTA SVD kernels on B200:
// DFMA / DADD / DMUL chains, FP64<->FP32 conversion, MUFU.RSQ, double shuffle, __syncthreads at 512 threads, LDS.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(double* out, long long* cyc, double seed) {
  __shared__ double sm[1024];
  const int tid = threadIdx.x;
  sm[tid] = seed + tid; sm[tid + 512] = seed;
  __syncthreads();
  double a = seed + tid * 1e-9, b = 1.0000001, c = 1e-9;
  float f = (float)seed;
  const int N = 512;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (MODE == 0) a = fma(a, b, c);
    if (MODE == 1) a = a + c;
    if (MODE == 2) a = a * b;
    if (MODE == 3) { f = (float)a; a = (double)f + c; }              // cvt down + cvt up + add
    if (MODE == 4) { f = rsq(f) + 1.0f; }                           // MUFU + FADD
    if (MODE == 5) a += __shfl_xor_sync(0xffffffffu, a, 1);          // shuffle (2 x 32 bit) + DADD
    if (MODE == 6) { __syncthreads(); }
    if (MODE == 7) { a = sm[((int)a) & 511]; }                       // dependent LDS (cvt in the chain too)
    if (MODE == 8) { f = fmaf(f, 1.0001f, 1e-3f); }
    if (MODE == 9) { sm[tid] = a; __syncthreads(); a = sm[(tid + 32) & 511] + c; }  // exchange through smem
  }
  long long t1 = clock64();
  if (tid == 0) cyc[MODE] = (t1 - t0);
  out[MODE * 512 + tid] = a + f;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 16 * 512 * 8); cudaMallocManaged(&cyc, 16 * 8);
  const char* names[] = {"DFMA", "DADD", "DMUL", "F2F.down+up+DADD", "MUFU.RSQ+FADD", "SHFL.f64+DADD", "BAR.SYNC(512)", "I2F..LDS dep", "FFMA", "STS+BAR+LDS+DADD"};
  for (int threads : {32, 512}) {
    k<0><<<1, threads>>>(out, cyc, 1.0); k<1><<<1, threads>>>(out, cyc, 1.0); k<2><<<1, threads>>>(out, cyc, 1.0);
    k<3><<<1, threads>>>(out, cyc, 1.0); k<4><<<1, threads>>>(out, cyc, 1.0); k<5><<<1, threads>>>(out, cyc, 1.0);
    k<6><<<1, threads>>>(out, cyc, 1.0); k<7><<<1, threads>>>(out, cyc, 1.0); k<8><<<1, threads>>>(out, cyc, 1.0);
    k<9><<<1, threads>>>(out, cyc, 1.0);
    cudaDeviceSynchronize();
    for (int m = 0; m < 10; ++m) printf("threads=%3d %-20s %.1f cycles/iter\n", threads, names[m], cyc[m] / 512.0);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
