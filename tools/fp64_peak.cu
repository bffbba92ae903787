// FP64 peak microbenchmark for B200 (sm_100a): measures the roofline denominators that
// MEASURED_PEAKS.json does not carry -- DMMA (mma.sync f64) and plain DFMA throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu
// Prints one JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int ACC>
__global__ void k_dmma884(double* out, int iters) {
  double c0[ACC], c1[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void k_dmma1688(double* out, int iters) {
  double c[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
  double a[4], b[2];
  for (int j = 0; j < 4; ++j) a[j] = 1.0 + threadIdx.x * 1e-9 * (j + 1);
  for (int j = 0; j < 2; ++j) b[j] = 1.0 - threadIdx.x * 1e-9 * (j + 1);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ACC>
__global__ void k_dfma(double* out, int iters) {
  double c[ACC];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F launch, int reps) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch(); CK(cudaDeviceSynchronize());
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [", p.name, sms);
  bool first = true;
  auto report = [&](const char* name, int threads, int ctas_per_sm, double flops, double ms) {
    printf("%s\n {\"kernel\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}",
           first ? "" : ",", name, threads, ctas_per_sm, ms, flops / ms * 1e-9);
    first = false;
  };
  for (int threads : {128, 256, 512, 1024}) {
    for (int cps : {1, 2}) {
      if (threads * cps > 2048) continue;
      int grid = sms * cps;
      double warps = (double)grid * threads / 32;
      {
        double ms = time_ms([&] { k_dmma884<16><<<grid, threads>>>(out, iters); }, 5);
        report("dmma_m8n8k4_acc16", threads, cps, warps * iters * 16 * 512.0, ms);
      }
      {
        double ms = time_ms([&] { k_dmma1688<8><<<grid, threads>>>(out, iters); }, 5);
        report("dmma_m16n8k8_acc8", threads, cps, warps * iters * 8 * 2048.0, ms);
      }
      {
        double ms = time_ms([&] { k_dfma<16><<<grid, threads>>>(out, iters); }, 5);
        report("dfma_acc16", threads, cps, warps * 32 * iters * 16 * 2.0, ms);
      }
    }
  }
  // sustained DMMA: ~3 s back to back (power-capped clocks)
  {
    int grid = sms * 2, threads = 512;
    double warps = (double)grid * threads / 32;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_dmma884<16><<<grid, threads>>>(out, iters); CK(cudaDeviceSynchronize());
    int n = 0; CK(cudaEventRecord(e0));
    for (; n < 400; ++n) k_dmma884<16><<<grid, threads>>>(out, iters * 4);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    report("dmma_m8n8k4_sustained", threads, 2, warps * iters * 4 * 16 * 512.0 * n, ms);
  }
  printf("\n]}\n");
  CK(cudaGetLastError());
  return 0;
}
