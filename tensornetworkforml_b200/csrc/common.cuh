// Shared device/host helpers for libtnml (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "tnml.h"

#define TNML_CUDA_ERR(e) (-(1000 + (int)(e)))

// Number of kernels this library has launched in this process (statistics only; read through
// tnml_kernel_launches()).  Defined in contract.cu.
extern unsigned long long g_tnml_kernel_launches;
#define TNML_COUNT(n) (g_tnml_kernel_launches += (unsigned long long)(n))

// Return the launch status of the kernel(s) just enqueued (no synchronisation).
static inline int tnml_launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? TNML_OK : TNML_CUDA_ERR(e);
}

#define TNML_REQUIRE(cond) \
  do {                     \
    if (!(cond)) return TNML_ERR_INVALID; \
  } while (0)

#define TNML_F64_ONLY(dtype)                          \
  do {                                                \
    if ((dtype) == TNML_F32) return TNML_ERR_UNSUPPORTED; \
    if ((dtype) != TNML_F64) return TNML_ERR_INVALID;     \
  } while (0)

// Function attributes (dynamic shared memory opt-in, non-portable cluster size) are PER DEVICE: every launcher keeps one
// DeviceOnce and sets its attributes the first time it runs on each device.  Two host threads racing through the same
// launcher both set the (idempotent) attributes; the bit is published only afterwards.
static inline int tnml_current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d & 63;
}
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool needed(int dev) const { return !((done.load(std::memory_order_acquire) >> dev) & 1ULL); }
  void mark(int dev) { done.fetch_or(1ULL << dev, std::memory_order_release); }
};
// SM count of the current device (148 on B200), queried once per device.
static inline int tnml_num_sms() {
  static std::atomic<int> cache[64];
  const int d = tnml_current_device();
  int v = cache[d].load(std::memory_order_relaxed);
  if (v <= 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || v <= 0) v = 148;
    cache[d].store(v, std::memory_order_relaxed);
  }
  return v;
}

static inline int64_t tnml_align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline int tnml_cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

namespace tnml {


// ---- FP64 tensor-core MMA (SASS DMMA.8x8x4) ----------------------------------------------------------
// A 8x4 row-major : lane holds A[lane>>2][lane&3]
// B 4x8 col-major : lane holds B[lane&3][lane>>2]
// C 8x8           : lane holds C[lane>>2][2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- cp.async (LDGSTS) 8-byte copy with zero fill when !valid ---------------------------------------
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS UBLKCP) ----------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned phase) {
  unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(a),
      "r"(phase)
      : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bytes must be a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  unsigned b = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
               "l"(gmem_src), "r"(bytes), "r"(b)
               : "memory");
}

// C = alpha op(A) op(B) + beta C like tnml_gemm (bond.cu); returns at once on the device when *skip_if != 0.
int gemm_if(const double* skip_if, int tA, int tB, int M, int N, int K, double alpha, const double* A, int lda,
            const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t st);

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace tnml
