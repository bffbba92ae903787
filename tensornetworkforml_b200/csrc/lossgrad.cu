// Activation + loss derivative + metrics (fused elementwise pass) and the gradient: a K = Ns reduction on
// FP64 tensor cores (DMMA) with a static split-K and a fixed-order second stage.
#include "common.cuh"
#include "f32_path.cuh"

namespace tnml {

// ---------------------------------------------------------------------------------------------------
// act / loss derivative / metrics.  One thread per sample; L is small (2..10).
// ---------------------------------------------------------------------------------------------------
constexpr int AL_THREADS = 256;
constexpr int AL_MAXL = 64;

__global__ void __launch_bounds__(AL_THREADS) k_act_lossder(const double* __restrict__ f, const int* __restrict__ y,
                                                           const double2* __restrict__ phi_p,
                                                           const double2* __restrict__ phi_q, double* __restrict__ q,
                                                           double* __restrict__ pp, double* __restrict__ partial,
                                                           int64_t Ns, int L, int act, int loss, double T) {
  __shared__ double red[3][AL_THREADS];
  const int64_t b = (int64_t)blockIdx.x * AL_THREADS + threadIdx.x;
  double n_ok = 0.0, abs_err = 0.0, abs_f = 0.0;
  if (b < Ns) {
    const double* fb = f + b * L;
    const int yb = y[b];
    double denom = 1.0, shift = 0.0;
    const bool softmax = act == TNML_ACT_SOFTMAX || act == TNML_ACT_SOFTMAX_STABLE;
    if (act == TNML_ACT_SOFTMAX_STABLE) {   // opt-in: subtract the largest logit (exact in infinite precision)
      shift = fb[0];
      for (int l = 1; l < L; ++l) shift = fmax(shift, fb[l]);
    }
    if (softmax) {  // NC:794 -- TNML_ACT_SOFTMAX is NOT max-stabilised, exactly like the reference
      denom = 0.0;
      for (int l = 0; l < L; ++l) denom += exp((fb[l] - shift) / T);
    }
    const double2 p = phi_p[b], r = phi_q[b];
    const double w0 = p.x * r.x, w1 = p.x * r.y, w2 = p.y * r.x, w3 = p.y * r.y;
    double best = 0.0;
    int arg = 0;
    for (int l = 0; l < L; ++l) {
      double v = fb[l], fa;
      abs_f += fabs(v);                                   // NC:744 (debug history: mean |f_orig|)
      if (act == TNML_ACT_LINEAR) fa = v;
      else if (act == TNML_ACT_SIGMOID) fa = 1.0 / (1.0 + exp(-v / T));  // NC:791
      else fa = exp((v - shift) / T) / denom;
      if (l == 0 || fa > best) { best = fa; arg = l; }  // np.argmax: first maximum
      const double yl = (l == yb) ? 1.0 : 0.0;
      abs_err += fabs(yl - fa);
      double g;
      if (loss == TNML_LOSS_MSE) g = yl - fa;  // NC:824
      else if (loss == TNML_LOSS_CROSS_ENTROPY)
        g = softmax ? (yl - yl * fa) / T : yl / fa;  // NC:828, NC:830
      else g = 1.0 / ((l == yb ? fa : fa - 1.0) + 1e-4);               // NC:832-833
      double* qo = q + (b * L + l) * 4;
      qo[0] = g * w0; qo[1] = g * w1; qo[2] = g * w2; qo[3] = g * w3;
    }
    n_ok = (arg == yb) ? 1.0 : 0.0;
    double* po = pp + b * 4;
    po[0] = w0; po[1] = w1; po[2] = w2; po[3] = w3;
  }
  red[0][threadIdx.x] = n_ok;
  red[1][threadIdx.x] = abs_err;
  red[2][threadIdx.x] = abs_f;
  __syncthreads();
  for (int s = AL_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
      red[2][threadIdx.x] += red[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    partial[3 * blockIdx.x] = red[0][0];
    partial[3 * blockIdx.x + 1] = red[1][0];
    partial[3 * blockIdx.x + 2] = red[2][0];
  }
}

// Standalone activation on the reference's own layout f[l][b] (L x Ns), one thread per sample.   NC:767-796
__global__ void __launch_bounds__(256) k_apply_act(const double* __restrict__ f, double* __restrict__ out, int64_t Ns,
                                                   int L, int act, double T) {
  const int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (b >= Ns) return;
  double denom = 1.0, shift = 0.0;
  if (act == TNML_ACT_SOFTMAX_STABLE) {
    shift = f[b];
    for (int l = 1; l < L; ++l) shift = fmax(shift, f[(int64_t)l * Ns + b]);
  }
  if (act == TNML_ACT_SOFTMAX || act == TNML_ACT_SOFTMAX_STABLE) {
    denom = 0.0;
    for (int l = 0; l < L; ++l) denom += exp((f[(int64_t)l * Ns + b] - shift) / T);
  }
  for (int l = 0; l < L; ++l) {
    const double v = f[(int64_t)l * Ns + b];
    double fa;
    if (act == TNML_ACT_LINEAR) fa = v;
    else if (act == TNML_ACT_SIGMOID) fa = 1.0 / (1.0 + exp(-v / T));
    else fa = exp((v - shift) / T) / denom;
    out[(int64_t)l * Ns + b] = fa;
  }
}

// Standalone loss derivative, dense target y[l][b] (the reference takes a one-hot float array).   NC:800-835
__global__ void __launch_bounds__(256) k_loss_der(const double* __restrict__ fa, const double* __restrict__ y,
                                                  double* __restrict__ out, int64_t n, int act, int loss, double T) {
  const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  const double f = fa[e], yy = y[e];
  double g;
  if (loss == TNML_LOSS_MSE) g = yy - f;
  else if (loss == TNML_LOSS_CROSS_ENTROPY)
    g = (act == TNML_ACT_SOFTMAX || act == TNML_ACT_SOFTMAX_STABLE) ? (yy - yy * f) / T : yy / f;
  else g = 1.0 / ((yy == 0.0 ? f - 1.0 : f) + 1e-4);
  out[e] = g;
}

// fixed-order final sum of the per-block partials (one block)
__global__ void __launch_bounds__(256) k_metrics_final(const double* __restrict__ partial, int nblocks,
                                                      double* __restrict__ metrics, double count) {
  __shared__ double red[3][256];
  double a = 0.0, e = 0.0, af = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) { a += partial[3 * i]; e += partial[3 * i + 1]; af += partial[3 * i + 2]; }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = e; red[2][threadIdx.x] = af;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
      red[2][threadIdx.x] += red[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { metrics[0] = red[0][0]; metrics[1] = red[1][0]; metrics[2] = count; metrics[3] = red[2][0]; }
}

// ---------------------------------------------------------------------------------------------------
// Gradient.  For one label l, one 64-wide slice of the left bond (a) and one 64-wide slice of the right
// bond (c):   dB_l[a][(st, c)] = sum_b L[b][a] * ( q[b][l][st] * R[b][c] ),   st = 2*sigma + tau
// i.e. a (64 x Ns) . (Ns x 256) GEMM whose B operand is formed on the fly from R and the 4 q values.
// CTA = 8 warps as 2 (a) x 4 (st): warp tile 32 x 64 -> 32 DMMA accumulator fragments.
// Samples are streamed in 16-row stages through a 3-deep cp.async pipeline.
// grid = (L * a_chunks * c_chunks, ksplit); each CTA writes its partial tile to ws[split].
// ---------------------------------------------------------------------------------------------------
constexpr int GR_BK = 16, GR_STAGES = 3, GR_LS = 68, GR_RS = 68, GR_QS = 8;
constexpr int GR_STAGE_DOUBLES = GR_BK * (GR_LS + GR_RS + GR_QS);
constexpr int GR_SMEM_BYTES = GR_STAGES * GR_STAGE_DOUBLES * 8;

// FULL: every a-slice and c-slice is a complete 64 (no per-fragment predicates in the MMA loop: a predicate around
// mma.sync costs a WARPSYNC + NOP + ISETP/BRA per DMMA, which halves the tensor-pipe issue rate).
template <bool FULL>
__global__ void __launch_bounds__(256, 1) k_grad(const double* __restrict__ q, const double* __restrict__ Lenv,
                                                 const double* __restrict__ Renv, double* __restrict__ ws, int64_t Ns,
                                                 int Dl, int Dr, int L, int a_chunks, int c_chunks, int64_t chunk) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;  // wn == st of this warp's columns

  int col = blockIdx.x;
  const int cc = col % c_chunks; col /= c_chunks;
  const int ac = col % a_chunks; col /= a_chunks;
  const int l = col;
  const int a0 = ac * 64, c0 = cc * 64;
  const int an = min(64, Dl - a0), cn = min(64, Dr - c0);

  const int64_t bstart = (int64_t)blockIdx.y * chunk;
  const int64_t bend = min(Ns, bstart + chunk);
  const int nst = bend > bstart ? (int)((bend - bstart + GR_BK - 1) / GR_BK) : 0;

  double acc[4][8][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // which fragments carry real rows / columns (warp-uniform)
  bool mt_ok[4], nt_ok[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) mt_ok[i] = (wm * 32 + i * 8) < an;
#pragma unroll
  for (int j = 0; j < 8; ++j) nt_ok[j] = (j * 8) < cn;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * GR_STAGE_DOUBLES; };
  auto issue = [&](int it) {
    double* Ls = stage_ptr(it % GR_STAGES);
    double* Rs = Ls + GR_BK * GR_LS;
    double* Qs = Rs + GR_BK * GR_RS;
    const int64_t bb = bstart + (int64_t)it * GR_BK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = tid + 256 * i;  // 0..1023
      int r = e >> 6, cidx = e & 63;
      int64_t b = bb + r;
      bool rowok = b < bend;
      cp_async8(Ls + r * GR_LS + cidx, Lenv + (rowok ? b : 0) * Dl + a0 + (cidx < an ? cidx : 0), rowok && cidx < an);
      cp_async8(Rs + r * GR_RS + cidx, Renv + (rowok ? b : 0) * Dr + c0 + (cidx < cn ? cidx : 0), rowok && cidx < cn);
    }
    if (tid < GR_BK * 4) {
      int r = tid >> 2, j = tid & 3;
      int64_t b = bb + r;
      bool rowok = b < bend;
      cp_async8(Qs + r * GR_QS + j, q + ((rowok ? b : 0) * L + l) * 4 + j, rowok);
    }
  };

  for (int s = 0; s < GR_STAGES - 1; ++s) {
    if (s < nst) issue(s);
    cp_async_commit();
  }
  for (int it = 0; it < nst; ++it) {
    cp_async_wait<GR_STAGES - 2>();
    __syncthreads();
    if (it + GR_STAGES - 1 < nst) issue(it + GR_STAGES - 1);
    cp_async_commit();
    const double* Ls = stage_ptr(it % GR_STAGES);
    const double* Rs = Ls + GR_BK * GR_LS;
    const double* Qs = Rs + GR_BK * GR_RS;
#pragma unroll
    for (int k4 = 0; k4 < GR_BK; k4 += 4) {
      const int kr = k4 + t;
      double af[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = Ls[kr * GR_LS + wm * 32 + i * 8 + g];
      const double qv = Qs[kr * GR_QS + wn];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (FULL || nt_ok[j]) {
          const double bv = qv * Rs[kr * GR_RS + j * 8 + g];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (FULL || mt_ok[i]) dmma(acc[i][j][0], acc[i][j][1], af[i], bv);
        }
      }
    }
  }
  cp_async_wait<0>();

  // partial tile -> ws[split] in bond-tensor layout [a][sigma][l][tau][c]
  double* out = ws + (size_t)blockIdx.y * ((size_t)Dl * 4 * L * Dr);
  const int sigma = wn >> 1, tau = wn & 1;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = wm * 32 + i * 8 + g;
    if (a < an) {
      double* row = out + ((((size_t)(a0 + a) * 2 + sigma) * L + l) * 2 + tau) * Dr + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = j * 8 + 2 * t;
        if (c < cn) row[c] = acc[i][j][0];
        if (c + 1 < cn) row[c + 1] = acc[i][j][1];
      }
    }
  }
}

// dB[e] = sum_i ws[i][e] in the fixed order i = 0 .. ks-1
__global__ void __launch_bounds__(256) k_grad_reduce(const double* __restrict__ ws, double* __restrict__ dB, int64_t n,
                                                    int ks) {
  int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (e >= n) return;
  double s = ws[e];
  for (int i = 1; i < ks; ++i) s += ws[(size_t)i * n + e];
  dB[e] = s;
}

static void grad_plan(int64_t Ns, int Dl, int Dr, int L, int* cols, int* ks, int64_t* chunk) {
  int a_chunks = tnml_cdiv(Dl, 64), c_chunks = tnml_cdiv(Dr, 64);
  *cols = L * a_chunks * c_chunks;
  int k = tnml_num_sms() / *cols;
  int kmax = tnml_cdiv(Ns, 4 * GR_BK);
  if (k > kmax) k = kmax;
  if (k < 1) k = 1;
  int64_t ch = tnml_align_up((Ns + k - 1) / k, GR_BK);
  *ks = tnml_cdiv(Ns, ch);
  *chunk = ch;
}

}  // namespace tnml

using namespace tnml;

extern "C" int64_t tnml_act_lossder_workspace_bytes(int64_t Ns) {
  return (int64_t)tnml_cdiv(Ns, AL_THREADS) * 3 * 8;
}

extern "C" int tnml_act_lossder(const void* f, const int32_t* y, const void* phi_p, const void* phi_q, void* q, void* pp,
                                void* metrics, void* ws, int64_t Ns, int32_t L, int32_t act, int32_t loss, double T,
                                int32_t dtype, tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(f && y && phi_p && phi_q && q && pp && metrics && ws && Ns > 0 && L > 0 && L <= AL_MAXL);
  TNML_REQUIRE(act >= 0 && act <= 3 && loss >= 0 && loss <= 2);
  if (dtype == TNML_F32)
    return f32::act_lossder((const float*)f, y, (const float*)phi_p, (const float*)phi_q, (float*)q, (float*)pp,
                            (double*)metrics, (double*)ws, Ns, L, act, loss, T, (cudaStream_t)stream);
  int nb = tnml_cdiv(Ns, AL_THREADS);
  TNML_COUNT(1);
  k_act_lossder<<<nb, AL_THREADS, 0, (cudaStream_t)stream>>>((const double*)f, y, (const double2*)phi_p,
                                                            (const double2*)phi_q, (double*)q, (double*)pp, (double*)ws,
                                                            Ns, L, act, loss, T);
  TNML_COUNT(1);
  k_metrics_final<<<1, 256, 0, (cudaStream_t)stream>>>((const double*)ws, nb, (double*)metrics, (double)Ns);
  return tnml_launch_status();
}

extern "C" int tnml_apply_act(const void* f, void* out, int64_t Ns, int32_t L, int32_t act, double T, int32_t dtype,
                              tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(f && out && Ns > 0 && L > 0 && act >= 0 && act <= 3);
  TNML_COUNT(1);
  k_apply_act<<<tnml_cdiv(Ns, 256), 256, 0, (cudaStream_t)stream>>>((const double*)f, (double*)out, Ns, L, act, T);
  return tnml_launch_status();
}

extern "C" int tnml_loss_derivative(const void* fa, const void* y, void* out, int64_t Ns, int32_t L, int32_t act,
                                    int32_t loss, double T, int32_t dtype, tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(fa && y && out && Ns > 0 && L > 0 && act >= 0 && act <= 3 && loss >= 0 && loss <= 2);
  TNML_COUNT(1);
  k_loss_der<<<tnml_cdiv(Ns * L, 256), 256, 0, (cudaStream_t)stream>>>((const double*)fa, (const double*)y, (double*)out,
                                                                      Ns * L, act, loss, T);
  return tnml_launch_status();
}

extern "C" int64_t tnml_grad_workspace_bytes(int64_t Ns, int32_t Dl, int32_t Dr, int32_t L) {
  int cols, ks;
  int64_t chunk;
  grad_plan(Ns, Dl, Dr, L, &cols, &ks, &chunk);
  const int64_t b64 = (int64_t)ks * Dl * 4 * L * Dr * 8, b32 = f32::grad_workspace_bytes(Ns, Dl, Dr, L);
  return b64 > b32 ? b64 : b32;   // the query has no dtype argument: large enough for both variants
}

extern "C" int tnml_grad(const void* q, const void* Lenv, const void* Renv, void* dB, void* ws, int64_t Ns, int32_t Dl,
                         int32_t Dr, int32_t L, int32_t dtype, tnml_stream_t stream) {
  TNML_REQUIRE(dtype == TNML_F64 || dtype == TNML_F32);
  TNML_REQUIRE(q && Lenv && Renv && dB && ws && Ns > 0 && Dl > 0 && Dr > 0 && L > 0);
  if (dtype == TNML_F32) {   // q = loss derivative [Ns][L] followed (16-byte aligned) by a copy of pp [Ns][4]
    const float* g = (const float*)q;
    return f32::grad(g, g + ((Ns * L + 3) & ~(int64_t)3), (const float*)Lenv, (const float*)Renv, (double*)dB, ws, Ns,
                     Dl, Dr, L, (cudaStream_t)stream);
  }
  static DeviceOnce attr_once;
  const int attr_dev = tnml_current_device();
  if (attr_once.needed(attr_dev)) {
    cudaError_t e = cudaFuncSetAttribute(k_grad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    e = cudaFuncSetAttribute(k_grad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GR_SMEM_BYTES);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    attr_once.mark(attr_dev);
  }
  int cols, ks;
  int64_t chunk;
  grad_plan(Ns, Dl, Dr, L, &cols, &ks, &chunk);
  dim3 grid(cols, ks);
  TNML_COUNT(1);
  if (Dl % 64 == 0 && Dr % 64 == 0)
    k_grad<true><<<grid, 256, GR_SMEM_BYTES, (cudaStream_t)stream>>>((const double*)q, (const double*)Lenv,
                                                                    (const double*)Renv, (double*)ws, Ns, Dl, Dr, L,
                                                                    Dl / 64, Dr / 64, chunk);
  else
    k_grad<false><<<grid, 256, GR_SMEM_BYTES, (cudaStream_t)stream>>>((const double*)q, (const double*)Lenv,
                                                                     (const double*)Renv, (double*)ws, Ns, Dl, Dr, L,
                                                                     tnml_cdiv(Dl, 64), tnml_cdiv(Dr, 64), chunk);
  int64_t n = (int64_t)Dl * 4 * L * Dr;
  TNML_COUNT(1);
  k_grad_reduce<<<tnml_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>((const double*)ws, (double*)dB, n, ks);
  return tnml_launch_status();
}
