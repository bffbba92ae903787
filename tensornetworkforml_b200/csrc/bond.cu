// Batch-independent pieces of a bond update: small GEMMs (bond-tensor formation, norm environments, L2 term),
// regularisation + clipping + update.   NC:484, NC:728-761, NC:966-1179
#include "common.cuh"

namespace tnml {

// ---------------------------------------------------------------------------------------------------
// Row-major C = alpha * op(A) . op(B) + beta * C.  T x T tile (T = 64: 4x4 register tile, T = 32: 2x2), 256 threads,
// plain DFMA (on B200 the DFMA and DMMA peaks coincide; these products are at most a few hundred MFLOP and never
// leave L2).  The next k-slice is fetched into registers while the current one is multiplied: the kernels are short
// chains of k-slices on few CTAs, i.e. bound by the L2 latency of each slice, not by throughput.  T = 32 is chosen for
// small outputs (four times the CTAs: the warm-started split's 128 x 256 panels run on 32 instead of 8 SMs).
// TA: A stored K x M (element (m,k) at k*lda + m).  TB: B stored N x K (element (k,n) at n*ldb + k).
// skip_if (device, optional): the kernel returns at once when *skip_if != 0.
// ---------------------------------------------------------------------------------------------------
template <bool TA, bool TB, int T>
__global__ void __launch_bounds__(256) k_gemm(int M, int N, int K, double alpha, const double* __restrict__ A, int lda,
                                              const double* __restrict__ B, int ldb, double beta, double* __restrict__ C,
                                              int ldc, const double* __restrict__ skip_if) {
  constexpr int R = T / 16;            // register tile R x R
  constexpr int PT = T * 16 / 256;     // elements of each operand slice per thread (4 or 2)
  __shared__ double As[16][T + 1];
  __shared__ double Bs[16][T + 1];
  if (skip_if && *skip_if != 0.0) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  double acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.0;
  double pa[PT], pb[PT];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const int e = tid + 256 * i;
      int m, k;
      if (TA) { m = e % T; k = e / T; } else { k = e & 15; m = e >> 4; }
      pa[i] = (m0 + m < M && k0 + k < K) ? (TA ? A[(size_t)(k0 + k) * lda + m0 + m] : A[(size_t)(m0 + m) * lda + k0 + k]) : 0.0;
      int n, kk;
      if (TB) { kk = e & 15; n = e >> 4; } else { n = e % T; kk = e / T; }
      pb[i] = (n0 + n < N && k0 + kk < K) ? (TB ? B[(size_t)(n0 + n) * ldb + k0 + kk] : B[(size_t)(k0 + kk) * ldb + n0 + n]) : 0.0;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int i = 0; i < PT; ++i) {
      const int e = tid + 256 * i;
      int m, k;
      if (TA) { m = e % T; k = e / T; } else { k = e & 15; m = e >> 4; }
      As[k][m] = pa[i];
      int n, kk;
      if (TB) { kk = e & 15; n = e >> 4; } else { n = e % T; kk = e / T; }
      Bs[kk][n] = pb[i];
    }
    __syncthreads();
    if (k0 + 16 < K) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double a[R], b[R];
#pragma unroll
      for (int i = 0; i < R; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < R; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      double* c = C + (size_t)m * ldc + n;
      *c = (beta == 0.0) ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*c);
    }
  }
}

template <int T>
static void launch_gemm_t(int tA, int tB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                          int ldb, double beta, double* C, int ldc, cudaStream_t st, const double* skip_if) {
  dim3 grid(tnml_cdiv(N, T), tnml_cdiv(M, T));
  if (!tA && !tB) k_gemm<false, false, T><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, skip_if);
  else if (tA && !tB) k_gemm<true, false, T><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, skip_if);
  else if (!tA && tB) k_gemm<false, true, T><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, skip_if);
  else k_gemm<true, true, T><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, skip_if);
}

static int launch_gemm(int tA, int tB, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
                       int ldb, double beta, double* C, int ldc, cudaStream_t st, const double* skip_if = nullptr) {
  TNML_COUNT(1);
  // fewer than one 64 x 64 tile per SM: 32 x 32 tiles spread the product over four times as many SMs
  if ((int64_t)tnml_cdiv(M, 64) * tnml_cdiv(N, 64) < tnml_num_sms())
    launch_gemm_t<32>(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, st, skip_if);
  else
    launch_gemm_t<64>(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, st, skip_if);
  return tnml_launch_status();
}

// ---------------------------------------------------------------------------------------------------
// bond update, stage 1: d = dB - reg, block partial sums of |B|, |d|, <B, G>, |reg|
// ---------------------------------------------------------------------------------------------------
constexpr int BU_THREADS = 256;

__global__ void __launch_bounds__(BU_THREADS) k_bu_partial(const double* __restrict__ B, const double* __restrict__ dB,
                                                          const double* __restrict__ G, double* __restrict__ D,
                                                          double* __restrict__ partial, int n, double wd, int L2_flag) {
  __shared__ double red[4][BU_THREADS];
  int e = blockIdx.x * BU_THREADS + threadIdx.x;
  double sb = 0.0, sd = 0.0, sg = 0.0, sr = 0.0;
  if (e < n) {
    double b = B[e];
    double reg, gg = 0.0;
    if (L2_flag) { gg = G[e]; reg = 2 * wd * gg; }  // NC:1176
    else reg = wd * b;                                // NC:733
    double d = dB[e] - reg;                          // NC:730 / NC:734
    D[e] = d;
    sb = fabs(b); sd = fabs(d); sg = b * gg; sr = fabs(reg);   // |reg|: debug history, NC:747
  }
  red[0][threadIdx.x] = sb; red[1][threadIdx.x] = sd; red[2][threadIdx.x] = sg; red[3][threadIdx.x] = sr;
  __syncthreads();
  for (int s = BU_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int r = 0; r < 4; ++r) red[r][threadIdx.x] += red[r][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[4 * blockIdx.x] = red[0][0];
    partial[4 * blockIdx.x + 1] = red[1][0];
    partial[4 * blockIdx.x + 2] = red[2][0];
    partial[4 * blockIdx.x + 3] = red[3][0];
  }
}

// stage 2: every block re-reduces the partials in the same fixed order, then clips and updates its slice
__global__ void __launch_bounds__(BU_THREADS) k_bu_apply(const double* __restrict__ B, const double* __restrict__ D,
                                                        const double* __restrict__ partial, int nblocks,
                                                        double* __restrict__ Bnew, double* __restrict__ stats, int n,
                                                        double lr, double wd) {
  __shared__ double red[4][BU_THREADS];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += BU_THREADS) {
    s0 += partial[4 * i]; s1 += partial[4 * i + 1]; s2 += partial[4 * i + 2]; s3 += partial[4 * i + 3];
  }
  red[0][threadIdx.x] = s0; red[1][threadIdx.x] = s1; red[2][threadIdx.x] = s2; red[3][threadIdx.x] = s3;
  __syncthreads();
  for (int s = BU_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
#pragma unroll
      for (int r = 0; r < 4; ++r) red[r][threadIdx.x] += red[r][threadIdx.x + s];
    }
    __syncthreads();
  }
  const double sumB = red[0][0], sumD = red[1][0], l2 = red[2][0], sumR = red[3][0];
  const bool clip = sumD > sumB;  // NC:756
  const double ratio = clip ? sumD / sumB : 1.0;
  int e = blockIdx.x * BU_THREADS + threadIdx.x;
  if (e < n) {
    double d = D[e];
    if (clip) d = d / ratio;  // NC:757 (division by the ratio, like the reference)
    d = d * lr;               // NC:760
    Bnew[e] = B[e] + d;       // NC:761
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[0] = sumB; stats[1] = sumD; stats[2] = wd * l2; stats[3] = clip ? 1.0 : 0.0;
    stats[4] = sumB / n; stats[5] = sumD / n; stats[6] = sumR / n; stats[7] = 0.0;
  }
}

}  // namespace tnml

using namespace tnml;

namespace tnml {
int gemm_if(const double* skip_if, int tA, int tB, int M, int N, int K, double alpha, const double* A, int lda,
            const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t st) {
  return launch_gemm(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, st, skip_if);
}
}  // namespace tnml

extern "C" int tnml_gemm(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, double alpha, const void* A,
                         int32_t lda, const void* B, int32_t ldb, double beta, void* C, int32_t ldc, int32_t dtype,
                         tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && lda > 0 && ldb > 0 && ldc > 0);
  return launch_gemm(transA, transB, M, N, K, alpha, (const double*)A, lda, (const double*)B, ldb, beta, (double*)C, ldc,
                     (cudaStream_t)stream);
}

extern "C" int tnml_l2_term(const void* B, const void* EL, const void* ER, void* G, void* ws, int32_t Dl, int32_t Dr,
                            int32_t L, int32_t dtype, tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(B && EL && ER && G && ws && Dl > 0 && Dr > 0 && L > 0 && G != B);
  cudaStream_t st = (cudaStream_t)stream;
  double* T1 = (double*)ws;
  // T1[a'][(s,l,t,c)] = sum_a EL[a'][a] B[a][...] ; G[(a,s,l,t)][c'] = sum_c T1[...][c] ER[c][c']      NC:1129-1135
  int rc = launch_gemm(0, 0, Dl, 4 * L * Dr, Dl, 1.0, (const double*)EL, Dl, (const double*)B, 4 * L * Dr, 0.0, T1,
                       4 * L * Dr, st);
  if (rc) return rc;
  return launch_gemm(0, 0, Dl * 4 * L, Dr, Dr, 1.0, T1, Dr, (const double*)ER, Dr, 0.0, (double*)G, Dr, st);
}

extern "C" int64_t tnml_bond_update_workspace_bytes(int32_t Dl, int32_t Dr, int32_t L) {
  int64_t n = (int64_t)Dl * 4 * L * Dr;
  return (n + 4 * (int64_t)tnml_cdiv(n, BU_THREADS)) * 8;
}

extern "C" int tnml_bond_update(const void* B, const void* dB, const void* G, void* Bnew, void* stats, void* ws,
                                int32_t Dl, int32_t Dr, int32_t L, double lr, double wd, int32_t L2_flag, int32_t dtype,
                                tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(B && dB && Bnew && stats && ws && Dl > 0 && Dr > 0 && L > 0);
  TNML_REQUIRE(Bnew != B && Bnew != dB);
  if (L2_flag) TNML_REQUIRE(G != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  const int n = Dl * 4 * L * Dr;
  double* D = (double*)ws;
  double* partial = D + n;
  const int nb = tnml_cdiv(n, BU_THREADS);
  TNML_COUNT(2);
  k_bu_partial<<<nb, BU_THREADS, 0, st>>>((const double*)B, (const double*)dB, (const double*)G, D, partial, n, wd,
                                          L2_flag);
  k_bu_apply<<<nb, BU_THREADS, 0, st>>>((const double*)B, D, partial, nb, (double*)Bnew, (double*)stats, n, lr, wd);
  return tnml_launch_status();
}

extern "C" int tnml_norm_env_step(const void* Ein, const void* site, void* Eout, void* ws, int32_t Dl, int32_t Dr,
                                  int32_t left_moving, int32_t dtype, tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(Ein && site && Eout && ws && Dl > 0 && Dr > 0 && Ein != Eout);
  cudaStream_t st = (cudaStream_t)stream;
  const double* E = (const double*)Ein;
  const double* A = (const double*)site;
  double* T = (double*)ws;
  int rc;
  if (!left_moving) {
    // T[a'][(s,m)] = sum_a E[a'][a] A[a][(s,m)] ; Eout[m][m'] = sum_{(a',s)} A[(a',s)][m] T[(a',s)][m']
    rc = launch_gemm(0, 0, Dl, 2 * Dr, Dl, 1.0, E, Dl, A, 2 * Dr, 0.0, T, 2 * Dr, st);
    if (rc) return rc;
    rc = launch_gemm(1, 0, Dr, Dr, 2 * Dl, 1.0, A, Dr, T, Dr, 0.0, (double*)Eout, Dr, st);
  } else {
    // T[(a,s)][c'] = sum_c A[(a,s)][c] E[c][c'] ; Eout[a][a'] = sum_{(s,c)} A[a][(s,c)] T[a'][(s,c)]
    rc = launch_gemm(0, 0, 2 * Dl, Dr, Dr, 1.0, A, Dr, E, Dr, 0.0, T, Dr, st);
    if (rc) return rc;
    rc = launch_gemm(0, 1, Dl, Dl, 2 * Dr, 1.0, A, 2 * Dr, T, 2 * Dr, 0.0, (double*)Eout, Dl, st);
  }
  return rc;
}
