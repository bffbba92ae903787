// Projection / prediction f_n = B' . Phi~_n on FP64 tensor cores (DMMA).   NC:494-523
//
// For one label l and one 64-wide slice of the right bond:
//   T[b][(st, c)] = sum_a L[b][a] * B'[a][st][l][c]          (32 samples x 64 a) . (64 a x 256)   -> DMMA
//   f[b][l]      += sum_{st,c} pp[b][st] * R[b][c] * T[b][(st, c)]                                 -> epilogue
// The B' slice (64 x 256) stays resident in shared memory for the whole CTA (weight-stationary); samples are
// streamed in 32-row stages through a cp.async pipeline.  grid = (L * c_chunks, ksplit).
#include <cstdlib>

#include "common.cuh"
#include "f32_path.cuh"

namespace tnml {

constexpr int PJ_BM = 32, PJ_STAGES = 2, PJ_LS = 68, PJ_RS = 68, PJ_PS = 4, PJ_BS = 260;
constexpr int PJ_STAGE_DOUBLES = PJ_BM * (PJ_LS + PJ_RS + PJ_PS);
constexpr int PJ_B_DOUBLES = 64 * PJ_BS;
constexpr int PJ_RED_DOUBLES = 8 * PJ_BM;
constexpr int PJ_SMEM_BYTES = (PJ_B_DOUBLES + PJ_STAGES * PJ_STAGE_DOUBLES + PJ_RED_DOUBLES) * 8;

// FULL: the a-slice and the c-slice are complete 64s (no predicates around mma.sync, see k_grad).
template <bool FULL>
__global__ void __launch_bounds__(256, 1) k_project(const double* __restrict__ Bn, const double* __restrict__ pp,
                                                    const double* __restrict__ Lenv, const double* __restrict__ Renv,
                                                    double* __restrict__ fout, int64_t Ns, int Dl, int Dr, int L,
                                                    int a0, int c_chunks, int64_t chunk, int64_t fpart_stride,
                                                    int accumulate) {
  extern __shared__ __align__(16) double smem[];
  double* Bs = smem;
  double* stages = smem + PJ_B_DOUBLES;
  double* red = stages + PJ_STAGES * PJ_STAGE_DOUBLES;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int st = warp >> 1;           // sigma/tau pair of this warp's 32 columns
  const int cw = (warp & 1) * 32;     // first c (within the slice) of this warp

  const int cc = blockIdx.x % c_chunks;
  const int l = blockIdx.x / c_chunks;
  const int c0 = cc * 64;
  const int an = min(64, Dl - a0), cn = min(64, Dr - c0);
  const int64_t bstart = (int64_t)blockIdx.y * chunk;
  const int64_t bend = min(Ns, bstart + chunk);
  const int nst = bend > bstart ? (int)((bend - bstart + PJ_BM - 1) / PJ_BM) : 0;

  auto issue = [&](int it) {
    double* Ls = stages + (size_t)(it % PJ_STAGES) * PJ_STAGE_DOUBLES;
    double* Rs = Ls + PJ_BM * PJ_LS;
    double* Ps = Rs + PJ_BM * PJ_RS;
    const int64_t bb = bstart + (int64_t)it * PJ_BM;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int e = tid + 256 * i;  // 0..2047
      int r = e >> 6, cidx = e & 63;
      int64_t b = bb + r;
      bool rowok = b < bend;
      cp_async8(Ls + r * PJ_LS + cidx, Lenv + (rowok ? b : 0) * Dl + a0 + (cidx < an ? cidx : 0), rowok && cidx < an);
      cp_async8(Rs + r * PJ_RS + cidx, Renv + (rowok ? b : 0) * Dr + c0 + (cidx < cn ? cidx : 0), rowok && cidx < cn);
    }
    if (tid < PJ_BM * 4) {
      int r = tid >> 2, j = tid & 3;
      int64_t b = bb + r;
      bool rowok = b < bend;
      cp_async8(Ps + r * PJ_PS + j, pp + (rowok ? b : 0) * 4 + j, rowok);
    }
  };

  if (nst > 0) issue(0);
  cp_async_commit();

  // resident B' slice: Bs[a][st*64 + c] = B'[a0+a][sigma][l][tau][c0+c]
  for (int e = tid; e < 64 * 256; e += 256) {
    int a = e >> 8, n = e & 255;
    int s2 = n >> 6, c = n & 63;
    double v = 0.0;
    if (a < an && c < cn) {
      int sigma = s2 >> 1, tau = s2 & 1;
      v = Bn[((((size_t)(a0 + a) * 2 + sigma) * L + l) * 2 + tau) * Dr + c0 + c];
    }
    Bs[a * PJ_BS + n] = v;
  }

  // warp-uniform activity flags
  bool nt_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) nt_ok[j] = (cw + j * 8) < cn;
  const int k4max = FULL ? 64 : ((an + 3) & ~3);

  for (int it = 0; it < nst; ++it) {
    if (it + 1 < nst) issue(it + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const double* Ls = stages + (size_t)(it % PJ_STAGES) * PJ_STAGE_DOUBLES;
    const double* Rs = Ls + PJ_BM * PJ_LS;
    const double* Ps = Rs + PJ_BM * PJ_RS;

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (FULL || nt_ok[0]) {
#pragma unroll 4
      for (int k4 = 0; k4 < k4max; k4 += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = Ls[(i * 8 + g) * PJ_LS + k4 + t];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = Bs[(k4 + t) * PJ_BS + st * 64 + cw + j * 8 + g];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (FULL || nt_ok[j]) {
#pragma unroll
            for (int i = 0; i < 4; ++i) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
          }
      }
    }
    // epilogue: rowsum_b = sum_n pp[b][st] * R[b][c(n)] * T[b][n]; reduce over the 4 lanes of a row, then warps
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + g;
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = cw + j * 8 + 2 * t;
        s = fma(Rs[r * PJ_RS + c], acc[i][j][0], s);
        s = fma(Rs[r * PJ_RS + c + 1], acc[i][j][1], s);
      }
      s *= Ps[r * PJ_PS + st];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (t == 0) red[warp * PJ_BM + r] = s;
    }
    __syncthreads();
    if (tid < PJ_BM) {
      const int64_t b = bstart + (int64_t)it * PJ_BM + tid;
      if (b < bend) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * PJ_BM + tid];
        double* dst = fout + (size_t)cc * fpart_stride + b * L + l;
        *dst = accumulate ? (*dst + s) : s;
      }
    }
    // the next iteration's first __syncthreads orders the reads of red/stage before they are overwritten
  }
  cp_async_wait<0>();
}

// f[e] = (accumulate ? f[e] : 0) + sum_cc fpart[cc][e]   (fixed order)
__global__ void __launch_bounds__(256) k_fpart_reduce(const double* __restrict__ fpart, double* __restrict__ f, int64_t n,
                                                     int parts, int accumulate) {
  int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  double s = accumulate ? f[e] : 0.0;
  for (int i = 0; i < parts; ++i) s += fpart[(size_t)i * n + e];
  f[e] = s;
}

static void project_plan(int64_t Ns, int Dr, int L, int max_ctas, int* cols, int* ks, int64_t* chunk) {
  int c_chunks = tnml_cdiv(Dr, 64);
  *cols = L * c_chunks;
  const int budget = (max_ctas > 0 && max_ctas < tnml_num_sms()) ? max_ctas : tnml_num_sms();
  int k = budget / *cols;
  int kmax = tnml_cdiv(Ns, 2 * PJ_BM);
  if (k > kmax) k = kmax;
  if (k < 1) k = 1;
  int64_t ch = tnml_align_up((Ns + k - 1) / k, PJ_BM);
  *ks = tnml_cdiv(Ns, ch);
  *chunk = ch;
}

}  // namespace tnml

using namespace tnml;

extern "C" int64_t tnml_project_workspace_bytes(int64_t Ns, int32_t Dl, int32_t Dr, int32_t L) {
  (void)Dl;
  int c_chunks = tnml_cdiv(Dr, 64);
  const int64_t b64 = c_chunks > 1 ? (int64_t)c_chunks * Ns * L * 8 : 8, b32 = f32::project_workspace_bytes(Ns, Dl, Dr, L);
  return b64 > b32 ? b64 : b32;   // the query has no dtype argument: large enough for both variants
}

extern "C" int tnml_project(const void* B, const void* pp, const void* Lenv, const void* Renv, void* f, void* ws,
                            int64_t Ns, int32_t Dl, int32_t Dr, int32_t L, int32_t max_ctas, int32_t dtype,
                            tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(B && pp && Lenv && Renv && f && ws && Ns > 0 && Dl > 0 && Dr > 0 && L > 0);
  if (dtype == TNML_F32)
    return f32::project((const double*)B, (const float*)pp, (const float*)Lenv, (const float*)Renv, (float*)f, ws, Ns,
                        Dl, Dr, L, max_ctas, (cudaStream_t)stream);
  static DeviceOnce attr_once;
  const int attr_dev = tnml_current_device();
  if (attr_once.needed(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(k_project<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PJ_SMEM_BYTES);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    e = cudaFuncSetAttribute(k_project<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PJ_SMEM_BYTES);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    attr_once.mark(attr_dev);
  }
  int cols, ks;
  int64_t chunk;
  project_plan(Ns, Dr, L, max_ctas, &cols, &ks, &chunk);
  const int c_chunks = tnml_cdiv(Dr, 64), a_chunks = tnml_cdiv(Dl, 64);
  dim3 grid(cols, ks);
  // the left bond is contracted 64 rows at a time; successive launches accumulate (deterministic: stream order)
  const bool full = (Dl % 64 == 0) && (Dr % 64 == 0);
  auto launch = [&](double* dst, int a0, int64_t stride, int accumulate) {
    TNML_COUNT(1);
    if (full)
      k_project<true><<<grid, 256, PJ_SMEM_BYTES, (cudaStream_t)stream>>>(
          (const double*)B, (const double*)pp, (const double*)Lenv, (const double*)Renv, dst, Ns, Dl, Dr, L, a0,
          c_chunks, chunk, stride, accumulate);
    else
      k_project<false><<<grid, 256, PJ_SMEM_BYTES, (cudaStream_t)stream>>>(
          (const double*)B, (const double*)pp, (const double*)Lenv, (const double*)Renv, dst, Ns, Dl, Dr, L, a0,
          c_chunks, chunk, stride, accumulate);
  };
  for (int ac = 0; ac < a_chunks; ++ac) {
    if (c_chunks == 1) {
      launch((double*)f, ac * 64, 0, ac > 0);
    } else {
      launch((double*)ws, ac * 64, (int64_t)Ns * L, 0);
      TNML_COUNT(1);
      k_fpart_reduce<<<tnml_cdiv(Ns * L, 256), 256, 0, (cudaStream_t)stream>>>((const double*)ws, (double*)f, Ns * L,
                                                                                c_chunks, ac > 0);
    }
  }
  return tnml_launch_status();
}
