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

// One-sided Jacobi on the rows of the symmetric PSD matrix G (n x n) held in shared memory.
// Output: Vt[k][:] = k-th eigenvector (unit), lam[k] = k-th eigenvalue, descending.
__global__ void __launch_bounds__(1024, 1) k_jacobi(const double* __restrict__ partial, int nparts, int n,
                                                    double* __restrict__ Vt, double* __restrict__ lam, int max_sweeps) {
  extern __shared__ __align__(16) double W[];  // n x n
  __shared__ double nrm[SVD_MAXN];
  __shared__ int rot_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;

  for (int e = tid; e < n * n; e += blockDim.x) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n * n + e];
    W[e] = s;
  }
  if (tid == 0) rot_flag = 0;
  __syncthreads();

  const int np = n + (n & 1);  // players in the round-robin tournament (one dummy if n is odd)
  const int npairs = np >> 1;
  const double tol = sqrt((double)n) * 2.220446049250313e-16;

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    for (int round = 0; round < np - 1; ++round) {
      for (int pi = warp; pi < npairs; pi += nwarps) {
        // circle method: position 0 is fixed, positions 1..np-1 rotate
        int pa = (pi == 0) ? 0 : 1 + (pi - 1 + round) % (np - 1);
        int pb = 1 + (np - 2 - pi + round) % (np - 1);
        int p = min(pa, pb), q = max(pa, pb);
        if (q >= n) continue;  // dummy player
        double* x = W + (size_t)p * n;
        double* y = W + (size_t)q * n;
        double xv[4], yv[4];
        double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int idx = lane + 32 * k;
          xv[k] = idx < n ? x[idx] : 0.0;
          yv[k] = idx < n ? y[idx] : 0.0;
          al = fma(xv[k], xv[k], al);
          be = fma(yv[k], yv[k], be);
          ga = fma(xv[k], yv[k], ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        if (fabs(ga) > tol * sqrt(al * be) && al > 0.0 && be > 0.0) {
          double zeta = (be - al) / (2.0 * ga);
          double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          double cs = 1.0 / sqrt(1.0 + tt * tt);
          double sn = cs * tt;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            int idx = lane + 32 * k;
            if (idx < n) {
              x[idx] = cs * xv[k] - sn * yv[k];
              y[idx] = sn * xv[k] + cs * yv[k];
            }
          }
          if (lane == 0) rot_flag = 1;
        }
      }
      __syncthreads();
    }
    int any = rot_flag;
    __syncthreads();
    if (tid == 0) rot_flag = 0;
    __syncthreads();
    if (!any) break;
  }

  // row norms = eigenvalues; rank them (descending, ties by index) and emit the unit rows in that order
  for (int r = warp; r < n; r += nwarps) {
    double s = 0.0;
    for (int idx = lane; idx < n; idx += 32) { double v = W[(size_t)r * n + idx]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) nrm[r] = sqrt(s);
  }
  __syncthreads();
  for (int r = warp; r < n; r += nwarps) {
    const double mine = nrm[r];
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      double other = nrm[o];
      rank += (other > mine) || (other == mine && o < r);
    }
    const double inv = mine > 0.0 ? 1.0 / mine : 0.0;
    for (int idx = lane; idx < n; idx += 32) Vt[(size_t)rank * n + idx] = W[(size_t)r * n + idx] * inv;
    if (lane == 0) lam[rank] = mine;
  }
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
    cudaError_t e = cudaFuncSetAttribute(k_jacobi, cudaFuncAttributeMaxDynamicSharedMemorySize, SVD_MAXN * SVD_MAXN * 8);
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
  const int jthreads = n >= 64 ? 1024 : (n >= 32 ? 512 : 256);
  const size_t jsmem = (size_t)n * n * 8;
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

  k_gram<<<p.nparts, 256, 0, st>>>(X, ss, sl, n, Nl, partial);
  k_jacobi<<<1, jthreads, jsmem, st>>>(partial, p.nparts, n, vt1, lam1, 40);
  if (refine) {
    const Idx3 dense{1, 1, 1, 0, 0};
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(n, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, nullptr, n, Y, Nl, dense);
    k_gram<<<p.nparts, 256, 0, st>>>(Y, Nl, 1, n, Nl, partial);
    k_jacobi<<<1, jthreads, jsmem, st>>>(partial, p.nparts, n, vt2, lam2, 40);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(Y, Nl, 1, n, Nl, vt2, lam2, m, dst_long,
                                                                        k_long_stride, map_long);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, vt2, lam2, n, m, dst_short, k_short_stride, map_short,
                                                   (double*)svals);
  } else {
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, lam1, m, dst_long,
                                                                        k_long_stride, map_long);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, nullptr, lam1, n, m, dst_short, k_short_stride, map_short,
                                                   (double*)svals);
  }
  return tnml_launch_status();
}
