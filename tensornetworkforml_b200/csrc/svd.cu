// SVD split of the updated bond tensor by one-sided (Hestenes) Jacobi.   NC:839-962
//
// Mx (R x C, row-major view of B') is short in one direction (n = min(R, C) <= 128 here) and long in the other
// (Nl = max(R, C)).  Pipeline, all on one stream, no host synchronisation:
//   gram      G  = Mx Mx^T (short side), split over the long side, fixed-order partial sums
//   jacobi    one-sided Jacobi on the rows of G inside ONE CTA (G lives in shared memory): rows converge to
//             lambda_k u_k^T, lambda = sigma^2; sorted descending
//   rows      Y  = U1^T Mx                                   (the rotated matrix, rows nearly orthogonal)
//   gram+jacobi again on Y: restores absolute accuracy eps*sigma_max for the small singular values, which a
//             single Gram pass loses (it squares the condition number)
//   rows      long factor  = S^-1/2 U^T Mx   = sqrt(S) Vh      written into the destination site layout
//   short     short factor = U sqrt(S)                         written into the destination site layout
#include "common.cuh"

namespace tnml {

struct Idx3 {  // i -> (i / (n2*n3)) * s1 + ((i / n3) % n2) * s2 + (i % n3) * s3
  int n2, n3;
  long long s1, s2, s3;
  __host__ __device__ long long operator()(int i) const {
    return (long long)(i / (n2 * n3)) * s1 + (long long)((i / n3) % n2) * s2 + (long long)(i % n3) * s3;
  }
};

constexpr int SVD_MAXN = 128;
constexpr int GRAM_LC = 32;  // long-side columns per CTA

// partial[blk][i*n + j] = sum_{l in chunk} In(i,l) In(j,l),  In(s,l) = X[s*ss + l*sl]
__global__ void __launch_bounds__(256) k_gram(const double* __restrict__ X, long long ss, long long sl, int n, int Nl,
                                              double* __restrict__ partial) {
  __shared__ double V[16][SVD_MAXN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int l0 = blockIdx.x * GRAM_LC;
  const int lend = min(Nl, l0 + GRAM_LC);
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int lb = l0; lb < lend; lb += 16) {
    __syncthreads();
    for (int e = tid; e < 16 * SVD_MAXN; e += 256) {
      int lj, s;
      if (sl == 1) { lj = e & 15; s = e >> 4; } else { s = e & (SVD_MAXN - 1); lj = e >> 7; }
      double v = 0.0;
      if (s < n && lb + lj < lend) v = X[(long long)s * ss + (long long)(lb + lj) * sl];
      V[lj][s] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int lj = 0; lj < 16; ++lj) {
      double a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { a[i] = V[lj][ty + 16 * i]; b[i] = V[lj][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
  }
  double* out = partial + (size_t)blockIdx.x * n * n;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = ty + 16 * i;
    if (r >= n) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = tx + 16 * j;
      if (c < n) out[r * n + c] = acc[i][j];
    }
  }
}

// One-sided (Hestenes) Jacobi on the rows of the symmetric PSD matrix G (n x n) held in shared memory, zero-padded
// to the compile-time size NP (32, 64 or 128; a zero row never rotates).  One CTA of 8*NP threads; each HALF-warp
// owns one row pair per round (NP/2 pairs per round), a lane holds NP/16 elements of each row as double2 vectors.
// Per round and pair: one dot product (the squared row norms are cached, updated by the rotation and refreshed
// once per sweep), a branch-free rotation-parameter evaluation (tan in single precision after a power-of-two
// rescale, then c = rsqrt(1 + t^2), s = c t in double: the rotation is orthogonal to double precision whatever
// the accuracy of t, which only affects the convergence rate), and the rotation itself.  The kernel is bound by
// instruction issue, so everything is unrolled at compile time and the round-robin schedule is incremental.
// Output: Vt[k][:] = k-th eigenvector (unit), lam[k] = k-th eigenvalue, descending; info[0] = sweeps used.
__device__ __forceinline__ double half_sum(double v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 1/sqrt(w) for w in [1, 2]: single-precision seed + two Newton steps in double (no special cases, no branches)
__device__ __forceinline__ double rsqrt_1_2(double w) {
  double r = (double)rsqrtf((float)w);
  const double h = 0.5 * w;
  r = fma(r, fma(-h * r, r, 0.5), r);
  r = fma(r, fma(-h * r, r, 0.5), r);
  return r;
}

template <int NP>
__global__ void __launch_bounds__(8 * NP, 1) k_jacobi(const double* __restrict__ partial, int nparts, int n,
                                                      double* __restrict__ Vt, double* __restrict__ lam, int max_sweeps,
                                                      double tol, double* __restrict__ info) {
  constexpr int V = NP / 32;  // double2 vectors per lane and row
  extern __shared__ __align__(16) double W[];  // NP x NP
  __shared__ double nrm2[NP];
  __shared__ int rot_count;
  const int tid = threadIdx.x, hw = tid >> 4, l16 = tid & 15;
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int NT = 8 * NP, NW = NT / 32;

  for (int e = tid; e < NP * NP; e += NT) {
    const int r = e / NP, c = e % NP;
    double s = 0.0;
    if (r < n && c < n)
      for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n * n + r * n + c];
    W[e] = s;
  }
  if (tid == 0) rot_count = 0;
  __syncthreads();

  const double tol2 = tol * tol;
  int sweeps_done = 0;
  // circle method: position 0 is fixed, positions 1..NP-1 rotate; half-warp hw plays position hw against NP-1-hw
  int ra = (hw == 0) ? 0 : hw - 1, rb = NP - 2 - hw;   // (position - 1 + round) mod (NP - 1)

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    {  // refresh the cached squared norms of rows hw and hw + NP/2
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = hw + rr * (NP / 2);
        const double2* row = reinterpret_cast<const double2*>(W + r * NP);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < V; ++k) { double2 v = row[l16 + 16 * k]; s = fma(v.x, v.x, s); s = fma(v.y, v.y, s); }
        s = half_sum(s);
        if (l16 == 0) nrm2[r] = s;
      }
    }
    __syncthreads();
    for (int round = 0; round < NP - 1; ++round) {
      const int p = (hw == 0) ? 0 : 1 + ra;
      const int q = 1 + rb;
      double2* x = reinterpret_cast<double2*>(W + p * NP);
      double2* y = reinterpret_cast<double2*>(W + q * NP);
      double2 xv[V], yv[V];
      double ga = 0.0;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        xv[k] = x[l16 + 16 * k];
        yv[k] = y[l16 + 16 * k];
        ga = fma(xv[k].x, yv[k].x, ga);
        ga = fma(xv[k].y, yv[k].y, ga);
      }
      ga = half_sum(ga);
      const double al = nrm2[p], be = nrm2[q];
      if (ga * ga > tol2 * al * be) {
        const double de = be - al, ta = 2.0 * ga;
        const double mx = fmax(fabs(de), fabs(ta));
        const int ex = (__double2hiint(mx) >> 20) & 0x7ff;
        if (ex > 0 && ex < 2040) {
          const double sc = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023-ex): mx*sc in [1,2)
          const float df = (float)(de * sc), tf = (float)(ta * sc);
          const float h = sqrtf(fmaf(df, df, tf * tf));
          const double t = (double)__fdividef(tf, df + copysignf(h, df));  // |t| <= 1
          const double cs = rsqrt_1_2(fma(t, t, 1.0));
          const double sn = cs * t;
#pragma unroll
          for (int k = 0; k < V; ++k) {
            double2 a = xv[k], b = yv[k], xo, yo;
            xo.x = cs * a.x - sn * b.x; xo.y = cs * a.y - sn * b.y;
            yo.x = sn * a.x + cs * b.x; yo.y = sn * a.y + cs * b.y;
            x[l16 + 16 * k] = xo;
            y[l16 + 16 * k] = yo;
          }
          if (l16 == 0) {
            nrm2[p] = al - t * ga;
            nrm2[q] = be + t * ga;
            rot_count = 1;
          }
        }
      }
      ra = (ra + 1 == NP - 1) ? 0 : ra + 1;
      rb = (rb + 1 == NP - 1) ? 0 : rb + 1;
      __syncthreads();
    }
    sweeps_done = sweep + 1;
    const int any = rot_count;
    __syncthreads();
    if (tid == 0) rot_count = 0;
    __syncthreads();
    if (!any) break;
  }
  if (tid == 0 && info) info[0] = (double)sweeps_done;

  // row norms = eigenvalues; rank them (descending, ties by index) and emit the unit rows in that order
  for (int r = warp; r < n; r += NW) {
    double s = 0.0;
    for (int idx = lane; idx < n; idx += 32) { double v = W[r * NP + idx]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) nrm2[r] = sqrt(s);
  }
  __syncthreads();
  for (int r = warp; r < n; r += NW) {
    const double mine = nrm2[r];
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      double other = nrm2[o];
      rank += (other > mine) || (other == mine && o < r);
    }
    const double inv = mine > 0.0 ? 1.0 / mine : 0.0;
    for (int idx = lane; idx < n; idx += 32) Vt[(size_t)rank * n + idx] = W[r * NP + idx] * inv;
    if (lane == 0) lam[rank] = mine;
  }
}

static void launch_jacobi(const double* partial, int nparts, int n, double* Vt, double* lam, double tol, double* info,
                          cudaStream_t st) {
  TNML_COUNT(1);
  if (n > 64) k_jacobi<128><<<1, 1024, 128 * 128 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, info);
  else if (n > 32) k_jacobi<64><<<1, 512, 64 * 64 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, info);
  else k_jacobi<32><<<1, 256, 32 * 32 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, info);
}

// Out[k][l] = scale_k * sum_s Vt[k][s] In(s,l), k < kmax; scale_k = lam_k^(-1/4) if lam != nullptr else 1
// written at out + k*kstride + map(l).  grid = (ceil(Nl/32), ceil(kmax/32)), 256 threads.
__global__ void __launch_bounds__(256) k_rows(const double* __restrict__ X, long long ss, long long sl, int n, int Nl,
                                              const double* __restrict__ Vt, const double* __restrict__ lam, int kmax,
                                              double* __restrict__ out, long long kstride, Idx3 map) {
  extern __shared__ __align__(16) double sm[];
  double* Vs = sm;                 // [32][n+1]
  double* Is = sm + 32 * (n + 1);  // [n][33]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int e = tid; e < 32 * n; e += 256) {
    int k = e / n, s = e % n;
    Vs[k * (n + 1) + s] = (k0 + k < kmax) ? Vt[(size_t)(k0 + k) * n + s] : 0.0;
  }
  for (int e = tid; e < 32 * n; e += 256) {
    int s, lj;
    if (sl == 1) { lj = e & 31; s = e >> 5; } else { s = e % n; lj = e / n; }
    Is[s * 33 + lj] = (l0 + lj < Nl) ? X[(long long)s * ss + (long long)(l0 + lj) * sl] : 0.0;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int k = warp + 8 * i;
    double s = 0.0;
    for (int t = 0; t < n; ++t) s = fma(Vs[k * (n + 1) + t], Is[t * 33 + lane], s);
    if (k0 + k < kmax && l0 + lane < Nl) {
      double sc = 1.0;
      if (lam) { double lv = lam[k0 + k]; sc = lv > 0.0 ? 1.0 / sqrt(sqrt(lv)) : 0.0; }
      out[(long long)(k0 + k) * kstride + map(l0 + lane)] = sc * s;
    }
  }
}

// Fs[s][k] = lam_k^(1/4) * sum_t Vt1[t][s] * Vt2[k][t]   (U = U1 U2); Vt2 == nullptr -> U = U1.
// Also emits the singular values sqrt(lam).
__global__ void __launch_bounds__(256) k_short(const double* __restrict__ Vt1, const double* __restrict__ Vt2,
                                               const double* __restrict__ lam, int n, int m, double* __restrict__ out,
                                               long long kstride, Idx3 map, double* __restrict__ svals) {
  int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < n) svals[idx] = sqrt(lam[idx]);
  if (idx >= n * m) return;
  const int s = idx % n, k = idx / n;
  double v;
  if (Vt2) {
    v = 0.0;
    for (int t = 0; t < n; ++t) v = fma(Vt1[(size_t)t * n + s], Vt2[(size_t)k * n + t], v);
  } else {
    v = Vt1[(size_t)k * n + s];
  }
  out[map(s) + (long long)k * kstride] = sqrt(sqrt(lam[k])) * v;
}

struct SvdPlan {
  int R, C, n, Nl, nparts;
  bool rows_short;
  size_t off_partial, off_vt1, off_vt2, off_lam1, off_lam2, off_Y, total;
};

static SvdPlan svd_plan(int Dl, int Dr, int L, int left_dir) {
  SvdPlan p;
  p.R = left_dir ? 2 * Dl * L : 2 * Dl;
  p.C = left_dir ? 2 * Dr : 2 * L * Dr;
  p.rows_short = p.R <= p.C;
  p.n = p.rows_short ? p.R : p.C;
  p.Nl = p.rows_short ? p.C : p.R;
  p.nparts = tnml_cdiv(p.Nl, GRAM_LC);
  size_t o = 0;
  p.off_partial = o; o += (size_t)p.nparts * p.n * p.n;
  p.off_vt1 = o; o += (size_t)p.n * p.n;
  p.off_vt2 = o; o += (size_t)p.n * p.n;
  p.off_lam1 = o; o += p.n;
  p.off_lam2 = o; o += p.n;
  p.off_Y = o; o += (size_t)p.n * p.Nl;
  p.total = o;
  return p;
}

}  // namespace tnml

using namespace tnml;

extern "C" int64_t tnml_svd_split_workspace_bytes(int32_t Dl, int32_t Dr, int32_t L, int32_t left_dir) {
  return (int64_t)svd_plan(Dl, Dr, L, left_dir).total * 8;
}

extern "C" int tnml_svd_split(const void* Bnew, void* site_p, void* site_q, void* svals, void* ws, int32_t Dl, int32_t Dr,
                              int32_t L, int32_t m, int32_t left_dir, int32_t refine, int32_t dtype,
                              tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(Bnew && site_p && site_q && svals && ws && Dl > 0 && Dr > 0 && L > 0);
  SvdPlan p = svd_plan(Dl, Dr, L, left_dir);
  TNML_REQUIRE(m > 0 && m <= p.n);
  if (p.n > SVD_MAXN) return TNML_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_jacobi<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    e = cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (32 * (SVD_MAXN + 1) + SVD_MAXN * 33) * 8);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    attr_set = true;
  }
  double* w = (double*)ws;
  double *partial = w + p.off_partial, *vt1 = w + p.off_vt1, *vt2 = w + p.off_vt2, *lam1 = w + p.off_lam1,
         *lam2 = w + p.off_lam2, *Y = w + p.off_Y;
  const double* X = (const double*)Bnew;
  const long long ss = p.rows_short ? p.C : 1, sl = p.rows_short ? 1 : p.C;
  const int n = p.n, Nl = p.Nl;
  const double tol_final = sqrt((double)n) * 2.220446049250313e-16;
  const size_t rsmem = (size_t)(32 * (n + 1) + n * 33) * 8;

  // destination maps (see tnml.h): rows of Mx -> site_p, columns of Mx -> site_q
  Idx3 rowmap, colmap;
  long long row_k, col_k;
  if (!left_dir) {
    rowmap = Idx3{1, 1, (long long)m, 0, 0}; row_k = 1;                                  // site_p[a][s][k]
    colmap = Idx3{2, Dr, (long long)Dr, (long long)L * Dr, 1}; col_k = 2LL * L * Dr;      // site_q[k][t][l][c]
  } else {
    rowmap = Idx3{2, L, (long long)L * 2 * m, (long long)m, 2LL * m}; row_k = 1;          // site_p[a][l][s][k]
    colmap = Idx3{1, 1, 1, 0, 0}; col_k = 2LL * Dr;                                       // site_q[k][t][c]
  }
  double* dst_short = (double*)(p.rows_short ? site_p : site_q);
  double* dst_long = (double*)(p.rows_short ? site_q : site_p);
  const Idx3 map_short = p.rows_short ? rowmap : colmap, map_long = p.rows_short ? colmap : rowmap;
  const long long k_short_stride = p.rows_short ? row_k : col_k, k_long_stride = p.rows_short ? col_k : row_k;

  TNML_COUNT(1);
  k_gram<<<p.nparts, 256, 0, st>>>(X, ss, sl, n, Nl, partial);
  launch_jacobi(partial, p.nparts, n, vt1, lam1, tol_final, (double*)svals + n, st);
  if (refine) {
    const Idx3 dense{1, 1, 1, 0, 0};
    TNML_COUNT(1);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(n, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, nullptr, n, Y, Nl, dense);
    TNML_COUNT(1);
    k_gram<<<p.nparts, 256, 0, st>>>(Y, Nl, 1, n, Nl, partial);
    launch_jacobi(partial, p.nparts, n, vt2, lam2, tol_final, (double*)svals + n + 1, st);
    TNML_COUNT(1);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(Y, Nl, 1, n, Nl, vt2, lam2, m, dst_long,
                                                                        k_long_stride, map_long);
    TNML_COUNT(1);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, vt2, lam2, n, m, dst_short, k_short_stride, map_short,
                                                   (double*)svals);
  } else {
    TNML_COUNT(1);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, lam1, m, dst_long,
                                                                        k_long_stride, map_long);
    TNML_COUNT(1);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, nullptr, lam1, n, m, dst_short, k_short_stride, map_short,
                                                   (double*)svals);
  }
  return tnml_launch_status();
}
