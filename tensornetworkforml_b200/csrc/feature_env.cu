// Feature map, input packing, environment advance (FP64 DMMA), forward's last contraction.
#include <cstdlib>

#include "common.cuh"
#include "f32_path.cuh"

namespace tnml {

// ---------------------------------------------------------------------------------------------------
// phi[s][b][:] = [sin(pi x/2), cos(pi x/2)]   (DG:165-167).  32x32 transpose tile: reads coalesced along s,
// writes coalesced along b.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_feature_map(const double* __restrict__ x, double2* __restrict__ phi, int64_t Ns,
                                                    int S) {
  __shared__ double tile[32][33];
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int64_t b = b0 + r;
    int s = s0 + tx;
    tile[r][tx] = (b < Ns && s < S) ? x[b * S + s] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int s = s0 + r;
    int64_t b = b0 + tx;
    if (b < Ns && s < S) {
      double v = tile[tx][r];
      double arg = 3.141592653589793 * v / 2;  // same operation order as np.pi*x/2
      phi[(int64_t)s * Ns + b] = make_double2(sin(arg), cos(arg));
    }
  }
}

// X[b][s][2] -> phi[s][b][2]   (NC:222-225 builds TX[i] = X[:, i, :])
__global__ void __launch_bounds__(256) k_pack_features(const double2* __restrict__ X, double2* __restrict__ phi,
                                                      int64_t Ns, int S) {
  __shared__ double2 tile[32][33];
  const int64_t b0 = (int64_t)blockIdx.x * 32;
  const int s0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    int64_t b = b0 + r;
    int s = s0 + tx;
    tile[r][tx] = (b < Ns && s < S) ? X[b * S + s] : make_double2(0.0, 0.0);
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int s = s0 + r;
    int64_t b = b0 + tx;
    if (b < Ns && s < S) phi[(int64_t)s * Ns + b] = tile[tx][r];
  }
}

// ---------------------------------------------------------------------------------------------------
// Environment advance: out[b][m] = sum_s phi[b][s] * sum_k E[b][k] W[k][s][m]
// A batched (Ns x K) . (K x 2M) GEMM on DMMA with the sigma-combination fused into the epilogue.
// CTA = 4 warps, 64 samples; warp = 16 samples x (2 sigma x 32 m).  K in chunks of 32, M in chunks of 32.
// ---------------------------------------------------------------------------------------------------
constexpr int EA_BM = 64, EA_KC = 32, EA_MC = 32;

__global__ void __launch_bounds__(128) k_env_advance(const double* __restrict__ E, const double2* __restrict__ phi,
                                                    const double* __restrict__ W, double* __restrict__ out, int64_t Ns,
                                                    int K, int M) {
  __shared__ double Es[EA_BM][EA_KC + 4];
  __shared__ double Ws[EA_KC][2 * EA_MC + 4];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t b0 = (int64_t)blockIdx.x * EA_BM;

  for (int mc = 0; mc < M; mc += EA_MC) {
    double acc[2][2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[i][s][n][0] = acc[i][s][n][1] = 0.0;

    for (int kc = 0; kc < K; kc += EA_KC) {
      __syncthreads();
      // E tile: 64 rows x 32 k
      {
        const int k = tid & 31;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int r = (tid >> 5) + 4 * i;
          int64_t b = b0 + r;
          Es[r][k] = (b < Ns && kc + k < K) ? E[b * K + kc + k] : 0.0;
        }
      }
      // W tile: 32 k x (2 sigma x 32 m)
      {
        const int mm = tid & 31;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int idx = (tid >> 5) + 4 * i;  // 0..63 -> (k, sigma)
          int k = idx >> 1, s = idx & 1;
          bool ok = (kc + k < K) && (mc + mm < M);
          Ws[k][s * EA_MC + mm] = ok ? W[((int64_t)(kc + k) * 2 + s) * M + mc + mm] : 0.0;
        }
      }
      __syncthreads();
      const int kmax = min(EA_KC, K - kc);
      for (int k4 = 0; k4 < kmax; k4 += 4) {
        double a0 = Es[warp * 16 + g][k4 + t];
        double a1 = Es[warp * 16 + 8 + g][k4 + t];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            double bv = Ws[k4 + t][s * EA_MC + n * 8 + g];
            dmma(acc[0][s][n][0], acc[0][s][n][1], a0, bv);
            dmma(acc[1][s][n][0], acc[1][s][n][1], a1, bv);
          }
      }
    }
    // epilogue: combine the two sigma planes with phi
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int64_t b = b0 + warp * 16 + i * 8 + g;
      if (b < Ns) {
        double2 p = phi[b];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          int m = mc + n * 8 + 2 * t;
          if (m < M) out[b * M + m] = p.x * acc[i][0][n][0] + p.y * acc[i][1][n][0];
          if (m + 1 < M) out[b * M + m + 1] = p.x * acc[i][0][n][1] + p.y * acc[i][1][n][1];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Environment advance, persistent TMA-staged version for K <= 64, M <= 64 (K % 4 == 0, M even): the site tensor W is
// staged ONCE per CTA in shared memory and the environment rows stream through a 3-stage ring, both with bulk
// asynchronous copies (cp.async.bulk, the 1-D TMA path: SASS UBLKCP) that complete on mbarriers -- one copy per
// (padded) row, so the DMMA fragment loads stay bank-conflict free (row strides = 8 mod 16 doubles / 4 mod 16).
// CTA = 8 warps = 4 (16 samples each) x 2 (32 output columns each, both sigma planes); tile = 64 samples.
// ---------------------------------------------------------------------------------------------------
constexpr int EB_BM = 64, EB_STAGES = 3, EB_ES = 68, EB_WS = 136;
constexpr int EB_SMEM_BYTES = (64 * EB_WS + EB_STAGES * EB_BM * EB_ES) * 8;

__global__ void __launch_bounds__(256, 1) k_env_advance_tma(const double* __restrict__ E, const double2* __restrict__ phi,
                                                           const double* __restrict__ W, double* __restrict__ out,
                                                           int64_t Ns, int K, int M) {
  extern __shared__ __align__(16) double smem[];
  double* Ws = smem;                         // [64][EB_WS]: W[k][sigma][m] at k*EB_WS + sigma*64 + m
  double* Es = smem + 64 * EB_WS;            // EB_STAGES x [EB_BM][EB_ES]
  __shared__ __align__(8) uint64_t wbar, full[EB_STAGES];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp >> 1, wc = warp & 1;   // 16-row group, 32-column half
  const int ntiles = (int)((Ns + EB_BM - 1) / EB_BM);

  if (tid == 0) {
    mbar_init(&wbar, 1);
    for (int s = 0; s < EB_STAGES; ++s) mbar_init(&full[s], 1);
    fence_mbar_init();
  }
  // columns m >= M of W are never copied: they must read as zero
  for (int e = tid; e < 64 * EB_WS; e += 256) Ws[e] = 0.0;
  fence_proxy_async();                       // generic-proxy zeros before the async-proxy copies into the same rows
  __syncthreads();
  auto issue = [&](int tile, int s) {        // warp 0: one bulk copy per environment row of the tile
    const int64_t b0 = (int64_t)tile * EB_BM;
    const int rows = (int)min((int64_t)EB_BM, Ns - b0);
    if (lane == 0) mbar_expect_tx(&full[s], (unsigned)(rows * K * 8));
    __syncwarp();
    for (int r = lane; r < rows; r += 32)
      tma_bulk_g2s(Es + ((size_t)s * EB_BM + r) * EB_ES, E + (b0 + r) * K, (unsigned)(K * 8), &full[s]);
  };
  if (warp == 0) {
    if (lane == 0) mbar_expect_tx(&wbar, (unsigned)(K * 2 * M * 8));
    __syncwarp();
    for (int i = lane; i < 2 * K; i += 32) {
      const int k = i >> 1, sg = i & 1;
      tma_bulk_g2s(Ws + (size_t)k * EB_WS + sg * 64, W + ((size_t)k * 2 + sg) * M, (unsigned)(M * 8), &wbar);
    }
    int tile = blockIdx.x;
    for (int s = 0; s < EB_STAGES - 1 && tile < ntiles; ++s, tile += gridDim.x) issue(tile, s);
  }
  mbar_wait(&wbar, 0);

  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it % EB_STAGES;
    const int64_t b0 = (int64_t)tile * EB_BM;
    double2 ph[2];                           // the epilogue's phi values: loaded before the MMA loop hides their latency
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t b = b0 + wr * 16 + i * 8 + g;
      ph[i] = b < Ns ? phi[b] : make_double2(0.0, 0.0);
    }
    mbar_wait(&full[s], (it / EB_STAGES) & 1);
    const double* Et = Es + (size_t)s * EB_BM * EB_ES;
    double acc[2][2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int sg = 0; sg < 2; ++sg)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[i][sg][n][0] = acc[i][sg][n][1] = 0.0;
#pragma unroll 2
    for (int k4 = 0; k4 < K; k4 += 4) {
      const double a0 = Et[(wr * 16 + g) * EB_ES + k4 + t];
      const double a1 = Et[(wr * 16 + 8 + g) * EB_ES + k4 + t];
#pragma unroll
      for (int sg = 0; sg < 2; ++sg)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const double bv = Ws[(k4 + t) * EB_WS + sg * 64 + wc * 32 + n * 8 + g];
          dmma(acc[0][sg][n][0], acc[0][sg][n][1], a0, bv);
          dmma(acc[1][sg][n][0], acc[1][sg][n][1], a1, bv);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t b = b0 + wr * 16 + i * 8 + g;
      if (b < Ns) {
        const double2 p = ph[i];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int m = wc * 32 + n * 8 + 2 * t;
          if (m + 1 < M) {
            *reinterpret_cast<double2*>(out + b * M + m) =
                make_double2(p.x * acc[i][0][n][0] + p.y * acc[i][1][n][0], p.x * acc[i][0][n][1] + p.y * acc[i][1][n][1]);
          } else if (m < M) {
            out[b * M + m] = p.x * acc[i][0][n][0] + p.y * acc[i][1][n][0];
          }
        }
      }
    }
    __syncthreads();                         // every warp is done with stage s: refill it with the tile 2 rounds ahead
    const int64_t nxt = (int64_t)tile + (int64_t)(EB_STAGES - 1) * gridDim.x;
    if (warp == 0 && nxt < ntiles) issue((int)nxt, (it + EB_STAGES - 1) % EB_STAGES);
  }
}

// Wt[c][s][a] = site[a][s][c]
__global__ void k_site_transpose(const double* __restrict__ site, double* __restrict__ Wt, int Dl, int Dr) {
  int n = Dl * 2 * Dr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int a = i % Dl, s = (i / Dl) & 1, c = i / (2 * Dl);
    Wt[i] = site[((int64_t)a * 2 + s) * Dr + c];
  }
}

// [a][s][l][c] <-> [a][l][s][c]
__global__ void k_label_site_swap(const double* __restrict__ in, double* __restrict__ out, int Dl, int Dr, int L,
                                  int to_left) {
  int n = Dl * 2 * L * Dr;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int c = i % Dr;
    int r = i / Dr;
    int a, s, l;
    if (to_left) {  // out index (a, l, s, c)
      s = r % 2; l = (r / 2) % L; a = r / (2 * L);
      out[i] = in[(((int64_t)a * 2 + s) * L + l) * Dr + c];
    } else {  // out index (a, s, l, c)
      l = r % L; s = (r / L) % 2; a = r / (2 * L);
      out[i] = in[(((int64_t)a * L + l) * 2 + s) * Dr + c];
    }
  }
}

// f[b][l] = sum L[b][a] phi[b][s] A[a][s][l][c] R[b][c]; one thread per (b, l).  Only used with Dl == 1 or
// Dr == 1 by forward(), so it is a thin FMA kernel.
__global__ void __launch_bounds__(256) k_site_predict(const double* __restrict__ Lenv, const double2* __restrict__ phi,
                                                     const double* __restrict__ A, const double* __restrict__ Renv,
                                                     double* __restrict__ f, int64_t Ns, int Dl, int Dr, int L) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Ns * L) return;
  int64_t b = idx / L;
  int l = (int)(idx % L);
  double2 p = phi[b];
  double sum = 0.0;
  for (int a = 0; a < Dl; ++a) {
    double la = Lenv[b * Dl + a];
    const double* A0 = A + (((int64_t)a * 2 + 0) * L + l) * Dr;
    const double* A1 = A + (((int64_t)a * 2 + 1) * L + l) * Dr;
    double s0 = 0.0, s1 = 0.0;
    for (int c = 0; c < Dr; ++c) {
      double r = Renv[b * Dr + c];
      s0 = fma(A0[c], r, s0);
      s1 = fma(A1[c], r, s1);
    }
    sum = fma(la, p.x * s0 + p.y * s1, sum);
  }
  f[idx] = sum;
}

}  // namespace tnml

namespace tnml {
// Synthetic images on the device (DG:42-50): label[b] drawn from a counter-based generator (splitmix64 of (seed, b)),
// x[b][s] = template[label[b]][s] * (1 - sigma) + u[b][s] * sigma with u uniform in [0, 1).  One thread per pixel.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long z) { return (double)(z >> 11) * (1.0 / 9007199254740992.0); }

__global__ void __launch_bounds__(256) k_generate_dataset(const double* __restrict__ templates, double* __restrict__ x,
                                                          int* __restrict__ labels, long long Ns, int S, int n_labels,
                                                          double sigma, double prob_first, unsigned long long seed) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= Ns * S) return;
  const long long b = e / S;
  const int s = (int)(e % S);
  const double ul = u01(splitmix64(seed * 0xD1342543DE82EF95ull + 2 * (unsigned long long)b + 1));
  int lab;
  if (prob_first >= 0.0) lab = ul < prob_first ? 0 : 1;                  // np.random.choice([0, 1], p=[p, 1-p])  DG:42
  else lab = min(n_labels - 1, (int)(ul * n_labels));
  if (s == 0) labels[b] = lab;
  const double un = u01(splitmix64((seed ^ 0xA5A5A5A5A5A5A5A5ull) * 0x9E3779B97F4A7C15ull + 2 * (unsigned long long)e));
  x[e] = templates[(long long)lab * S + s] * (1.0 - sigma) + un * sigma;   // DG:49-50
}
}  // namespace tnml

using namespace tnml;

extern "C" int tnml_generate_dataset(const void* templates, void* x, int32_t* labels, int64_t Ns, int32_t S,
                                     int32_t n_labels, double sigma, double prob_first, uint64_t seed,
                                     tnml_stream_t stream) {
  TNML_REQUIRE(templates && x && labels && Ns > 0 && S > 0 && n_labels > 0);
  TNML_REQUIRE(prob_first < 0.0 || n_labels == 2);
  TNML_COUNT(1);
  k_generate_dataset<<<tnml_cdiv(Ns * S, 256), 256, 0, (cudaStream_t)stream>>>(
      (const double*)templates, (double*)x, (int*)labels, Ns, S, n_labels, sigma, prob_first, (unsigned long long)seed);
  return tnml_launch_status();
}

extern "C" int tnml_feature_map(const void* x, void* phi, int64_t Ns, int32_t S, int32_t dtype, tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(x && phi && Ns > 0 && S > 0);
  if (dtype == TNML_F32) return f32::feature_map((const double*)x, (float*)phi, Ns, S, (cudaStream_t)stream);
  dim3 grid(tnml_cdiv(Ns, 32), tnml_cdiv(S, 32));
  TNML_COUNT(1);
  k_feature_map<<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, (double2*)phi, Ns, S);
  return tnml_launch_status();
}

extern "C" int tnml_pack_features(const void* X, void* phi, int64_t Ns, int32_t S, int32_t dtype,
                                  tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(X && phi && Ns > 0 && S > 0);
  if (dtype == TNML_F32) return f32::pack_features((const double*)X, (float*)phi, Ns, S, (cudaStream_t)stream);
  dim3 grid(tnml_cdiv(Ns, 32), tnml_cdiv(S, 32));
  TNML_COUNT(1);
  k_pack_features<<<grid, 256, 0, (cudaStream_t)stream>>>((const double2*)X, (double2*)phi, Ns, S);
  return tnml_launch_status();
}

extern "C" int tnml_env_advance(const void* E, const void* phi_p, const void* W, void* out, int64_t Ns, int32_t K,
                                int32_t M, int32_t dtype, tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(E && phi_p && W && out && Ns > 0 && K > 0 && M > 0);
  if (dtype == TNML_F32)
    return f32::env_advance((const float*)E, (const float*)phi_p, (const float*)W, (float*)out, Ns, K, M,
                            (cudaStream_t)stream);
  TNML_COUNT(1);
  static const int use_tma = [] {            // TNML_ENV_TMA=0: always the plain kernel (A/B knob)
    const char* ev = getenv("TNML_ENV_TMA");
    return (ev && atoi(ev) == 0) ? 0 : 1;
  }();
  static DeviceOnce attr_once;
  const int attr_dev = tnml_current_device();
  if (use_tma && attr_once.needed(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(k_env_advance_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, EB_SMEM_BYTES);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    attr_once.mark(attr_dev);
  }
  if (use_tma && K <= 64 && M <= 64 && K % 4 == 0 && M % 2 == 0 && Ns >= 4 * EB_BM &&
      (((uintptr_t)E | (uintptr_t)W | (uintptr_t)out) & 15) == 0) {
    const int ntiles = tnml_cdiv(Ns, EB_BM);
    k_env_advance_tma<<<ntiles < tnml_num_sms() ? ntiles : tnml_num_sms(), 256, EB_SMEM_BYTES, (cudaStream_t)stream>>>(
        (const double*)E, (const double2*)phi_p, (const double*)W, (double*)out, Ns, K, M);
    return tnml_launch_status();
  }
  k_env_advance<<<tnml_cdiv(Ns, EA_BM), 128, 0, (cudaStream_t)stream>>>((const double*)E, (const double2*)phi_p,
                                                                       (const double*)W, (double*)out, Ns, K, M);
  return tnml_launch_status();
}

extern "C" int tnml_site_transpose(const void* site, void* Wt, int32_t Dl, int32_t Dr, int32_t dtype,
                                   tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(site && Wt && Dl > 0 && Dr > 0);
  int n = Dl * 2 * Dr;
  TNML_COUNT(1);
  k_site_transpose<<<min(tnml_cdiv(n, 256), 1024), 256, 0, (cudaStream_t)stream>>>((const double*)site, (double*)Wt, Dl,
                                                                                   Dr);
  return tnml_launch_status();
}

extern "C" int tnml_label_site_swap(const void* in, void* out, int32_t Dl, int32_t Dr, int32_t L, int32_t to_left_layout,
                                    int32_t dtype, tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(in && out && in != out && Dl > 0 && Dr > 0 && L > 0);
  int n = Dl * 2 * L * Dr;
  TNML_COUNT(1);
  k_label_site_swap<<<min(tnml_cdiv(n, 256), 1024), 256, 0, (cudaStream_t)stream>>>((const double*)in, (double*)out, Dl,
                                                                                    Dr, L, to_left_layout);
  return tnml_launch_status();
}

extern "C" int tnml_site_predict(const void* Lenv, const void* phi_p, const void* A_label, const void* Renv, void* f,
                                 int64_t Ns, int32_t Dl, int32_t Dr, int32_t L, int32_t dtype, tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(Lenv && phi_p && A_label && Renv && f && Ns > 0 && Dl > 0 && Dr > 0 && L > 0);
  if (dtype == TNML_F32)
    return f32::site_predict((const float*)Lenv, (const float*)phi_p, (const float*)A_label, (const float*)Renv,
                             (float*)f, Ns, Dl, Dr, L, (cudaStream_t)stream);
  TNML_COUNT(1);
  k_site_predict<<<tnml_cdiv(Ns * L, 256), 256, 0, (cudaStream_t)stream>>>(
      (const double*)Lenv, (const double2*)phi_p, (const double*)A_label, (const double*)Renv, (double*)f, Ns, Dl, Dr, L);
  return tnml_launch_status();
}

extern "C" int tnml_convert_f32(const void* src_f64, void* dst_f32, int64_t n, tnml_stream_t stream) {
  TNML_REQUIRE(src_f64 && dst_f32 && n > 0);
  return f32::convert((const double*)src_f64, (float*)dst_f32, n, (cudaStream_t)stream);
}

extern "C" int tnml_site_weights_f32(const void* site_f64, void* Wt_f32, int32_t Dl, int32_t Dr, int32_t left_moving,
                                     tnml_stream_t stream) {
  TNML_REQUIRE(site_f64 && Wt_f32 && Dl > 0 && Dr > 0);
  return f32::site_weights((const double*)site_f64, (float*)Wt_f32, Dl, Dr, left_moving, (cudaStream_t)stream);
}
