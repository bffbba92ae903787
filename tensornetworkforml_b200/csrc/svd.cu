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
#include <cstdlib>

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
                                              double* __restrict__ partial, const double* __restrict__ skip_flag) {
  __shared__ double V[16][SVD_MAXN + 1];
  if (skip_flag && *skip_flag != 0.0) return;   // second pass not needed (see k_jacobi)
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
// Preconditioning (Drmac-Veselic): before the sweeps G is replaced in place by its diagonally pivoted Cholesky
// factor, G = sum_k r_k r_k^T (row r_k stored in the physical row of its pivot, so no permutation is ever applied).
// Orthogonalising the rows of R instead of the rows of G works on the spectrum sigma instead of sigma^2 and needs
// about half the sweeps; the rows converge to sigma_k v_k^T with v_k the eigenvectors of G.
// Output: Vt[k][:] = k-th eigenvector (unit), lam[k] = k-th eigenvalue, descending; info[0] = sweeps used.
__device__ __forceinline__ float rsqrt_approx(float x) {   // one MUFU.RSQ, no slow path
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Sum four values over the 32 lanes of a warp and leave all four sums on every lane: packed butterfly, 10 double
// shuffles instead of 20 (shuffles share the MIO queue with shared-memory traffic, which bounds this kernel).
__device__ __forceinline__ void warp_sum4(double& g0, double& g1, double& g2, double& g3, int lane) {
  const bool hi = lane & 16;
  double ka = hi ? g2 : g0, kb = hi ? g3 : g1;
  ka += __shfl_xor_sync(0xffffffffu, hi ? g0 : g2, 16);
  kb += __shfl_xor_sync(0xffffffffu, hi ? g1 : g3, 16);
  const bool h8 = lane & 8;
  double v = h8 ? kb : ka;
  v += __shfl_xor_sync(0xffffffffu, h8 ? ka : kb, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  g0 = __shfl_sync(0xffffffffu, v, 0);
  g1 = __shfl_sync(0xffffffffu, v, 8);
  g2 = __shfl_sync(0xffffffffu, v, 16);
  g3 = __shfl_sync(0xffffffffu, v, 24);
}

// Four independent row-pair rotations held in registers: rows x[i], y[i] (E elements per lane), cached squared
// norms nx[i], ny[i].  Returns true if any pair had a relative inner product above 1e-8 (a "large" rotation).
template <int E>
__device__ __forceinline__ bool rotate4(double (&x)[4][E], double (&y)[4][E], double (&nx)[4], double (&ny)[4],
                                        double tol2, int lane) {
  double g[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < E; ++k) acc = fma(x[i][k], y[i][k], acc);
    g[i] = acc;
  }
  warp_sum4(g[0], g[1], g[2], g[3], lane);
  bool any = false;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double al = nx[i], be = ny[i], ga = g[i];
    const int ex = (__double2hiint(al + be) >> 20) & 0x7ff;
    const double g2 = ga * ga, ab = al * be;
    if (g2 > tol2 * ab && ex > 0 && ex < 2040) {
      const double sc = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023-ex): (al+be)*sc in [1,2)
      // tan, cos, sin in single precision: t = 2ga / (de + sign(de) sqrt(de^2 + 4ga^2)), |t| <= 1
      const float df = (float)((be - al) * sc), tf = (float)((ga + ga) * sc);
      const float hh = fmaf(df, df, tf * tf);                      // in [~1e-30, 8]: no range issues
      const float h = hh * rsqrt_approx(hh);                       // sqrt via MUFU.RSQ
      const float t0 = __fdividef(tf, df + copysignf(h, df));
      const float cf = rsqrt_approx(fmaf(t0, t0, 1.0f));
      double cs = (double)cf, sn = (double)(cf * t0);
      // exact renormalisation in double: nu = (cs^2 + sn^2)^(-1/2) = 1 - e/2 + 3e^2/8, e ~ 1e-7 -> error ~ e^3
      const double e = fma(cs, cs, fma(sn, sn, -1.0));
      const double nu = fma(e, fma(e, 0.375, -0.5), 1.0);
      cs *= nu;
      sn *= nu;
#pragma unroll
      for (int k = 0; k < E; ++k) {
        const double a = x[i][k], b = y[i][k];
        x[i][k] = cs * a - sn * b;
        y[i][k] = sn * a + cs * b;
      }
      const double tg = (double)t0 * ga;
      nx[i] = al - tg;
      ny[i] = be + tg;
      any |= g2 > 1e-16 * ab;
    }
  }
  return any;
}

// NP: padded matrix size (32, 64, 128).  Rows are grouped in NP/4 blocks of 4; one WARP owns one block pair per
// block-round (circle method over the blocks), keeps the 8 rows in registers (lane holds elements lane + 32k) and
// performs all 16 cross rotations (4 sets of 4 independent ones) -- plus, in the first block-round of a sweep, the
// 6 + 6 rotations inside the two blocks -- before writing the rows back: the matrix crosses shared memory once per
// block-round instead of once per rotation round.
template <int NP>
__global__ void __launch_bounds__(NP * 4, 1) k_jacobi(const double* __restrict__ partial, int nparts, int n,
                                                      double* __restrict__ Vt, double* __restrict__ lam,
                                                      int max_sweeps, double tol, int use_chol,
                                                      double* __restrict__ info, double* __restrict__ skip_flag,
                                                      const double* __restrict__ lam_prev) {
  // Second-pass protocol: the first pass (use_chol = 1) sets *skip_flag = 1 when lambda_min / lambda_max > 1e-7
  // (sigma ratio > 3e-4: a single Gram pass is then already accurate to ~1e-12 sigma_max for every singular value);
  // the second pass (use_chol = 0) sees the flag, returns the identity rotation and the first-pass eigenvalues.
  if (!use_chol && skip_flag && *skip_flag != 0.0) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Vt[e] = (e / n == e % n) ? 1.0 : 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) lam[e] = lam_prev[e];
    if (threadIdx.x == 0 && info) info[0] = 0.0;
    return;
  }
  constexpr int NB = NP / 4;           // row blocks
  constexpr int NW = NB / 2;           // warps = block pairs per block-round
  constexpr int NT = NW * 32;
  constexpr int E = NP / 32;           // elements per lane and row
  extern __shared__ __align__(16) double W[];  // NP x NP
  __shared__ double nrm2[NP];
  __shared__ int rot_count;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int e = tid; e < NP * NP; e += NT) {
    const int r = e / NP, c = e % NP;
    double s = 0.0;
    if (r < n && c < n)
      for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * n * n + r * n + c];
    W[e] = s;
  }
  if (tid == 0) rot_count = 0;
  __syncthreads();

  // ---- diagonally pivoted Cholesky, in place and LEFT-looking: W <- R with G = R^T R ----
  // Step k picks the largest remaining diagonal entry c of the Schur complement (kept incrementally in diag[]),
  // forms only that row of the Schur complement, S[c][j] = G[c][j] - sum_{m<k} R_m[c] R_m[j], scales it and stores
  // it in physical row c (rows of not-yet-eliminated indices still hold G).  No trailing-matrix update, no
  // permutation: the rows of R are simply scattered by pivot.
  if (use_chol) {
    constexpr int NG = NT / NP;                          // thread groups splitting the sum over previous rows (4)
    __shared__ double part[NG][NP], diag[NP];
    __shared__ unsigned char active[NP];
    __shared__ int piv[NP];
    __shared__ double piv_floor;
    __shared__ int piv_idx;
    for (int j = tid; j < NP; j += NT) { active[j] = j < n; diag[j] = j < n ? W[j * NP + j] : 0.0; }
    __syncthreads();
    auto select_pivot = [&](bool first) {                // warp 0
      double best = -1.0;
      int bi = -1;
      for (int j = lane; j < n; j += 32)
        if (active[j]) { double v = diag[j]; if (v > best) { best = v; bi = j; } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        if (first) piv_floor = best * (double)n * 2.220446049250313e-16;
        piv_idx = (bi >= 0 && best > piv_floor) ? bi : -1;
      }
    };
    if (warp == 0) select_pivot(true);
    __syncthreads();
    const int j = tid % NP, gq = tid / NP;
    for (int k = 0; k < n; ++k) {
      const int c = piv_idx;
      if (c < 0) break;                                  // numerically rank deficient from here on
      double acc = 0.0;                                  // this group's share of sum_m R_m[c] R_m[j]
#pragma unroll 4
      for (int m = gq; m < k; m += NG) {
        const int rm = piv[m];
        acc = fma(W[rm * NP + c], W[rm * NP + j], acc);
      }
      part[gq][j] = acc;
      __syncthreads();
      if (gq == 0) {
        double scc = W[c * NP + c], sj = W[c * NP + j];
#pragma unroll
        for (int g = 0; g < NG; ++g) { scc -= part[g][c]; sj -= part[g][j]; }
        const double d = sqrt(fmax(scc, piv_floor));
        double r = 0.0;
        if (j == c) r = d;
        else if (active[j]) { r = sj / d; diag[j] = fma(-r, r, diag[j]); }
        W[c * NP + j] = r;
        if (j == c) { active[c] = 0; piv[k] = c; }
      }
      __syncthreads();
      if (warp == 0) select_pivot(false);
      __syncthreads();
    }
    // Rows never eliminated (numerical rank deficiency): the Schur complement is below the resolution of this
    // pass.  Give them a tiny multiple of the unit vectors so that the sweeps still complete an orthonormal basis
    // (the second pass resolves the true small singular values inside that subspace).
    const double tiny = sqrt(piv_floor);
    for (int e = tid; e < NP * NP; e += NT)
      if (active[e / NP]) W[e] = (e / NP == e % NP) ? tiny : 0.0;
    __syncthreads();
  }

  const double tol2 = tol * tol;
  int sweeps_done = 0;
  // circle method over the NB blocks: position 0 is fixed, positions 1..NB-1 rotate; warp w plays w against NB-1-w
  int ra = (warp == 0) ? 0 : warp - 1, rb = NB - 2 - warp;   // (position - 1 + round) mod (NB - 1)

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    for (int r = warp; r < NP; r += NW) {   // refresh the cached squared row norms
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < E; ++k) { const double v = W[r * NP + lane + 32 * k]; s = fma(v, v, s); }
      s = warp_sum(s);
      if (lane == 0) nrm2[r] = s;
    }
    __syncthreads();
    bool rotated = false;
    for (int round = 0; round < NB - 1; ++round) {
      const int bi = (warp == 0) ? 0 : 1 + ra;
      const int bj = 1 + rb;
      double a[4][E], b[4][E], na[4], nb[4];
      double* const wa = W + 4 * bi * NP + lane;
      double* const wb = W + 4 * bj * NP + lane;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int k = 0; k < E; ++k) {
          a[i][k] = wa[i * NP + 32 * k];
          b[i][k] = wb[i * NP + 32 * k];
        }
        na[i] = nrm2[4 * bi + i];
        nb[i] = nrm2[4 * bj + i];
      }
      if (round == 0) {
        // pairs inside each block, once per sweep: (0,1)(2,3) | (0,2)(1,3) | (0,3)(1,2) for both blocks at once
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int p0 = 0, q0 = s + 1;                       // (0,1) (0,2) (0,3)
          const int p1 = (s == 0) ? 2 : 1, q1 = (s == 2) ? 2 : 3;   // (2,3) (1,3) (1,2)
          double x[4][E], y[4][E], nx[4], ny[4];
#pragma unroll
          for (int k = 0; k < E; ++k) {
            x[0][k] = a[p0][k]; y[0][k] = a[q0][k]; x[1][k] = a[p1][k]; y[1][k] = a[q1][k];
            x[2][k] = b[p0][k]; y[2][k] = b[q0][k]; x[3][k] = b[p1][k]; y[3][k] = b[q1][k];
          }
          nx[0] = na[p0]; ny[0] = na[q0]; nx[1] = na[p1]; ny[1] = na[q1];
          nx[2] = nb[p0]; ny[2] = nb[q0]; nx[3] = nb[p1]; ny[3] = nb[q1];
          rotated |= rotate4<E>(x, y, nx, ny, tol2, lane);
#pragma unroll
          for (int k = 0; k < E; ++k) {
            a[p0][k] = x[0][k]; a[q0][k] = y[0][k]; a[p1][k] = x[1][k]; a[q1][k] = y[1][k];
            b[p0][k] = x[2][k]; b[q0][k] = y[2][k]; b[p1][k] = x[3][k]; b[q1][k] = y[3][k];
          }
          na[p0] = nx[0]; na[q0] = ny[0]; na[p1] = nx[1]; na[q1] = ny[1];
          nb[p0] = nx[2]; nb[q0] = ny[2]; nb[p1] = nx[3]; nb[q1] = ny[3];
        }
      }
      // the 16 pairs across the two blocks: set s pairs a[i] with b[(i+s)&3]
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        double y[4][E], ny[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int k = 0; k < E; ++k) y[i][k] = b[(i + s) & 3][k];
          ny[i] = nb[(i + s) & 3];
        }
        rotated |= rotate4<E>(a, y, na, ny, tol2, lane);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int k = 0; k < E; ++k) b[(i + s) & 3][k] = y[i][k];
          nb[(i + s) & 3] = ny[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int k = 0; k < E; ++k) {
          wa[i * NP + 32 * k] = a[i][k];
          wb[i * NP + 32 * k] = b[i][k];
        }
        if (lane == 0) { nrm2[4 * bi + i] = na[i]; nrm2[4 * bj + i] = nb[i]; }
      }
      ra = (ra + 1 == NB - 1) ? 0 : ra + 1;
      rb = (rb + 1 == NB - 1) ? 0 : rb + 1;
      __syncthreads();
    }
    // Stop rule: cyclic Jacobi converges quadratically, so a sweep in which every rotated pair had a relative
    // inner product below 1e-8 leaves all of them below 1e-15 -- no separate verification sweep is needed.
    if (rotated && lane == 0) rot_count = 1;
    sweeps_done = sweep + 1;
    __syncthreads();
    const int any = rot_count;
    __syncthreads();
    if (tid == 0) rot_count = 0;
    __syncthreads();
    if (!any) break;
  }
  if (tid == 0 && info) info[0] = (double)sweeps_done;

  // squared row norms rank the rows (descending, ties by index); emit the unit rows in that order.
  // With the Cholesky factor the rows are sigma_k v_k^T (|row|^2 = lambda_k); without it lambda_k v_k^T.
  for (int r = warp; r < n; r += NW) {
    double s = 0.0;
    for (int idx = lane; idx < n; idx += 32) { double v = W[r * NP + idx]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) nrm2[r] = s;
  }
  __syncthreads();
  for (int r = warp; r < n; r += NW) {
    const double mine = nrm2[r];
    int rank = 0;
    for (int o = 0; o < n; ++o) {
      double other = nrm2[o];
      rank += (other > mine) || (other == mine && o < r);
    }
    const double nr = sqrt(mine);
    const double inv = mine > 0.0 ? 1.0 / nr : 0.0;
    for (int idx = lane; idx < n; idx += 32) Vt[(size_t)rank * n + idx] = W[r * NP + idx] * inv;
    if (lane == 0) lam[rank] = use_chol ? mine : nr;
  }
  if (use_chol && skip_flag && warp == 0) {
    double mn = 1e300, mx = 0.0;
    for (int r = lane; r < n; r += 32) { mn = fmin(mn, nrm2[r]); mx = fmax(mx, nrm2[r]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) *skip_flag = (mn > 1e-7 * mx) ? 1.0 : 0.0;
  }
}

static cudaError_t jacobi_prepare() {
  return cudaFuncSetAttribute(k_jacobi<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 128 * 8);
}

static void launch_jacobi(const double* partial, int nparts, int n, double* Vt, double* lam, double tol, int use_chol,
                          double* info, double* skip, const double* lam_prev, cudaStream_t st) {
  TNML_COUNT(1);
  if (n > 64)
    k_jacobi<128><<<1, 512, 128 * 128 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, use_chol, info, skip, lam_prev);
  else if (n > 32)
    k_jacobi<64><<<1, 256, 64 * 64 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, use_chol, info, skip, lam_prev);
  else
    k_jacobi<32><<<1, 128, 32 * 32 * 8, st>>>(partial, nparts, n, Vt, lam, 40, tol, use_chol, info, skip, lam_prev);
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
  size_t off_partial, off_vt1, off_vt2, off_lam1, off_lam2, off_Y, off_skip, total;
};

static SvdPlan svd_plan_rc(int R, int C);
static SvdPlan svd_plan(int Dl, int Dr, int L, int left_dir) {
  return svd_plan_rc(left_dir ? 2 * Dl * L : 2 * Dl, left_dir ? 2 * Dr : 2 * L * Dr);
}

// Shared implementation: X = R x C row-major matrix, rowmap/colmap = where row i / column j of the factors land.
static int svd_core(const double* X, SvdPlan p, int m, int refine, double* dst_rows, Idx3 rowmap, long long row_k,
                    double* dst_cols, Idx3 colmap, long long col_k, double* svals, double* w, cudaStream_t st) {
  if (p.n > SVD_MAXN) return TNML_ERR_UNSUPPORTED;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = jacobi_prepare();
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    e = cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (32 * (SVD_MAXN + 1) + SVD_MAXN * 33) * 8);
    if (e != cudaSuccess) return TNML_CUDA_ERR(e);
    attr_set = true;
  }
  double *partial = w + p.off_partial, *vt1 = w + p.off_vt1, *vt2 = w + p.off_vt2, *lam1 = w + p.off_lam1,
         *lam2 = w + p.off_lam2, *Y = w + p.off_Y, *skip = w + p.off_skip;
  const long long ss = p.rows_short ? p.C : 1, sl = p.rows_short ? 1 : p.C;
  const int n = p.n, Nl = p.Nl;
  const double tol_final = sqrt((double)n) * 2.220446049250313e-16;
  const size_t rsmem = (size_t)(32 * (n + 1) + n * 33) * 8;
  double* dst_short = p.rows_short ? dst_rows : dst_cols;
  double* dst_long = p.rows_short ? dst_cols : dst_rows;
  const Idx3 map_short = p.rows_short ? rowmap : colmap, map_long = p.rows_short ? colmap : rowmap;
  const long long k_short_stride = p.rows_short ? row_k : col_k, k_long_stride = p.rows_short ? col_k : row_k;

  TNML_COUNT(1);
  k_gram<<<p.nparts, 256, 0, st>>>(X, ss, sl, n, Nl, partial, nullptr);
  launch_jacobi(partial, p.nparts, n, vt1, lam1, tol_final, 1, svals + n, refine == 1 ? skip : nullptr, nullptr, st);
  if (refine) {
    const Idx3 dense{1, 1, 1, 0, 0};
    TNML_COUNT(4);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(n, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, nullptr, n, Y, Nl, dense);
    k_gram<<<p.nparts, 256, 0, st>>>(Y, Nl, 1, n, Nl, partial, refine == 1 ? skip : nullptr);
    launch_jacobi(partial, p.nparts, n, vt2, lam2, tol_final, 0, svals + n + 1, refine == 1 ? skip : nullptr, lam1, st);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(Y, Nl, 1, n, Nl, vt2, lam2, m, dst_long,
                                                                        k_long_stride, map_long);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, vt2, lam2, n, m, dst_short, k_short_stride, map_short, svals);
  } else {
    TNML_COUNT(2);
    k_rows<<<dim3(tnml_cdiv(Nl, 32), tnml_cdiv(m, 32)), 256, rsmem, st>>>(X, ss, sl, n, Nl, vt1, lam1, m, dst_long,
                                                                        k_long_stride, map_long);
    k_short<<<tnml_cdiv(n * m, 256), 256, 0, st>>>(vt1, nullptr, lam1, n, m, dst_short, k_short_stride, map_short, svals);
  }
  return tnml_launch_status();
}

static SvdPlan svd_plan_rc(int R, int C) {
  SvdPlan p;
  p.R = R;
  p.C = C;
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
  p.off_skip = o; o += 1;
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
  return svd_core((const double*)Bnew, p, m, refine, (double*)site_p, rowmap, row_k, (double*)site_q, colmap, col_k,
                  (double*)svals, (double*)ws, (cudaStream_t)stream);
}

extern "C" int64_t tnml_svd_workspace_bytes(int32_t R, int32_t C) { return (int64_t)svd_plan_rc(R, C).total * 8; }

extern "C" int tnml_svd(const void* Mx, void* US, void* SVh, void* svals, void* ws, int32_t R, int32_t C, int32_t m,
                        int32_t refine, int32_t dtype, tnml_stream_t stream) {
  TNML_F64_ONLY(dtype);
  TNML_REQUIRE(Mx && US && SVh && svals && ws && R > 0 && C > 0);
  SvdPlan p = svd_plan_rc(R, C);
  TNML_REQUIRE(m > 0 && m <= p.n);
  const Idx3 rowmap{1, 1, (long long)m, 0, 0};   // US[i][k]
  const Idx3 colmap{1, 1, 1, 0, 0};              // SVh[k][j]
  return svd_core((const double*)Mx, p, m, refine, (double*)US, rowmap, 1, (double*)SVh, colmap, C, (double*)svals,
                  (double*)ws, (cudaStream_t)stream);
}
