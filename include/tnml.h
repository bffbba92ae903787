/* tnml.h -- C ABI of the B200-native MPS-classifier sweep (libtnml.so).
 *
 * The reference (francescovidaich964/TensorNetworkForML) is pure Python + NumPy and has NO FFI of its own
 * (SURVEY.md section 2.1 / 8b): its boundary is the Python API of Network_class.py (NC), Tensor_class.py (TC)
 * and custom_linalg_tools.py (CLT).  This header is the C ABI that sits UNDER that Python API; every entry
 * point names the reference code it replaces (file:line under /root/reference/TensorNetwork).
 *
 * Conventions
 *  - plain C types only; all data pointers are DEVICE pointers unless a parameter says "host".
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing and keeps no
 *    global state: the caller owns every buffer, including workspaces sized by the *_workspace_bytes queries.
 *  - return value: 0 = ok; TNML_ERR_* (< 0) for argument errors; -(1000 + cudaError_t) for CUDA launch errors.
 *  - dtype: TNML_F64 is the parity path (FP64, DMMA tensor cores).  TNML_F32 is the FP32/TF32 variant: the PER-SAMPLE
 *    arrays (phi, env, f, q, pp and the weight operand Wt / A_label handed to tnml_env_advance / tnml_site_predict, see
 *    tnml_site_weights_f32 / tnml_convert_f32) are FP32, the tile-aligned contractions run on tcgen05 tensor cores (kind::tf32, FP32 accumulation in TMEM; ragged
 *    bond dimensions fall back to FP32 FMA kernels), while site tensors, bond tensors, the gradient sum dB, clipping
 *    and the SVD split stay FP64.  Entry points that accept TNML_F32: tnml_feature_map / tnml_pack_features (x, X stay
 *    FP64, phi is FP32), tnml_env_advance, tnml_site_predict, tnml_act_lossder, tnml_grad (dB FP64), tnml_project
 *    (B FP64); every other entry point is batch-independent and FP64 only (TNML_ERR_UNSUPPORTED for TNML_F32).
 *    With TNML_F32 the q buffer of tnml_act_lossder / tnml_grad holds the loss derivative g[Ns][L] followed, at the next
 *    multiple of 4 elements, by a copy of pp[Ns][4] (Ns*L + 4*Ns + 4 floats): the gradient is then the GEMM
 *    dB = sum_b (g L)(x)(pp R) of two Khatri-Rao operands formed on the fly.
 *
 * Device layouts ("canonical"), all row-major, d = 2 physical components:
 *    phi   [S][Ns][2]        feature-mapped input, site-major
 *    env   [Ns][D]           per-sample environment (left env of sites < p, or right env of sites >= p)
 *    site  [Dl][2][Dr]       plain site tensor          (a, sigma, c)
 *    label site, right sweep [Dl][2][L][Dr]  (a, sigma, l, c);  left sweep [Dl][L][2][Dr]  (a, l, sigma, c)
 *    B     [Dl][2][L][2][Dr] bond tensor                (a, sigma, l, tau, c)
 *    f, q  [Ns][L], [Ns][L][4]   outputs; q = lossder * phi_p(sigma) * phi_q(tau)
 */
#ifndef TNML_H
#define TNML_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tnml_stream_t; /* cudaStream_t */

enum { TNML_F64 = 0, TNML_F32 = 1 };
/* TNML_ACT_SOFTMAX is the reference's softmax, NOT max-stabilised (NC:794: exp(f/T) overflows to inf/inf = nan for
 * |f|/T > ~709); TNML_ACT_SOFTMAX_STABLE is the opt-in exp((f - max f)/T) / sum form: identical where the reference
 * is finite, finite everywhere. */
enum { TNML_ACT_LINEAR = 0, TNML_ACT_SIGMOID = 1, TNML_ACT_SOFTMAX = 2, TNML_ACT_SOFTMAX_STABLE = 3 };   /* NC:127 */
enum { TNML_LOSS_MSE = 0, TNML_LOSS_CROSS_ENTROPY = 1, TNML_LOSS_FULL_CROSS_ENT = 2 }; /* NC:132 */
enum {
  TNML_OK = 0,
  TNML_ERR_INVALID = -1,     /* bad dimension / null pointer / enum out of range */
  TNML_ERR_UNSUPPORTED = -2, /* dtype or size not implemented on this build */
  TNML_ERR_WORKSPACE = -3    /* workspace too small */
};

int tnml_version(void);
/* Kernels launched by this library in this process so far (statistics for bench.py's gpu_launches). */
uint64_t tnml_kernel_launches(void);
/* Human-readable text for a return code (static storage). */
const char* tnml_error_string(int code);

/* Asynchronous device-to-device copy of nbytes on `stream`. */
int tnml_copy(void* dst, const void* src, int64_t nbytes, tnml_stream_t stream);

/* A one-thread kernel that idles for about `ns` nanoseconds (at most 100 us) on `stream`: lets a caller make a kernel on
 * one stream eligible slightly after a kernel on another (see tnml_svd_split_ev). */
int tnml_delay(int64_t ns, tnml_stream_t stream);

/* Page-lock a caller-owned HOST buffer in place (and release it), so the host->device copy of the input batch is a
 * direct DMA transfer instead of a staged one.  Returns 0 (registered now), 1 (was already registered) or < 0. */
int tnml_host_register(void* host_ptr, uint64_t bytes);
int tnml_host_unregister(void* host_ptr);

/* ---- a1: feature map + input packing ------------------------------------------------------------
 * tnml_feature_map   : phi[s][b][:] = [sin(pi x[b][s]/2), cos(pi x[b][s]/2)]  (sin first)   DG:165-167, NC:152-155
 * tnml_pack_features : X[b][s][:] (the (Ns,S,2) array the reference API takes) -> phi[s][b][:]   NC:222-225 */
int tnml_feature_map(const void* x, void* phi, int64_t Ns, int32_t S, int32_t dtype, tnml_stream_t stream);
/* Synthetic images generated on the device (the caller side of DG:6-52): labels[b] from a counter-based generator
 * (prob_first >= 0: two labels, label 0 with that probability like np.random.choice at DG:42; prob_first < 0: uniform
 * over n_labels), x[b][s] = templates[labels[b]][s] * (1 - sigma) + uniform[0,1) * sigma (DG:49-50).  templates is a
 * device array [n_labels][S].  Deterministic in (seed, Ns, S); NOT NumPy's stream -- seeded reference scripts keep using
 * the host generator of data_generator.py. */
int tnml_generate_dataset(const void* templates, void* x, int32_t* labels, int64_t Ns, int32_t S, int32_t n_labels,
                          double sigma, double prob_first, uint64_t seed, tnml_stream_t stream);
int tnml_pack_features(const void* X, void* phi, int64_t Ns, int32_t S, int32_t dtype, tnml_stream_t stream);

/* ---- a5/a10: environment advance ------------------------------------------------------------------
 * out[b][m] = sum_sigma phi_p[b][sigma] * sum_k E[b][k] * W[k][sigma][m]
 * W is the site tensor seen from the side the environment comes from: right-moving (left env) W = site
 * [Dl][2][Dr] as stored; left-moving (right env) W[c][sigma][a] = site[a][sigma][c] (tnml_site_transpose).
 * Replaces contract(As[i],TX[i]) + contract(cum, A_TX) : NC:227-255, NC:637-642, NC:669-674 (CLT:81-84). */
int tnml_env_advance(const void* E, const void* phi_p, const void* W, void* out, int64_t Ns, int32_t K, int32_t M,
                     int32_t dtype, tnml_stream_t stream);
/* Wt[c][sigma][a] = site[a][sigma][c] */
int tnml_site_transpose(const void* site, void* Wt, int32_t Dl, int32_t Dr, int32_t dtype, tnml_stream_t stream);

/* ---- a5: last contraction of forward ----------------------------------------------------------------
 * f[b][l] = sum L[b][a] phi_p[b][sigma] A[a][sigma][l][c] R[b][c]   (label site in right-sweep layout)
 * NC:242 / NC:255 (r_cum_contraction[0] / l_cum_contraction[-1]). */
int tnml_site_predict(const void* Lenv, const void* phi_p, const void* A_label, const void* Renv, void* f, int64_t Ns,
                      int32_t Dl, int32_t Dr, int32_t L, int32_t dtype, tnml_stream_t stream);

/* ---- a11/a12/a7: activation, loss derivative, metrics ------------------------------------------------
 * fa = act(f / T) (NC:767-796, softmax NOT max-stabilised), g = dloss(fa, onehot(y)) (NC:800-835),
 * q[b][l][2*sigma+tau] = g[b][l] * phi_p[b][sigma] * phi_q[b][tau]   (operand of the gradient GEMM)
 * pp[b][2*sigma+tau]   = phi_p[b][sigma] * phi_q[b][tau]             (operand of the projection epilogue)
 * metrics[0] = number of samples with argmax(fa) == y, metrics[1] = sum |onehot(y) - fa|   (NC:697-702),
 * metrics[2] = Ns (the local sample count, so that the sharded sums can be all-reduced as they are),
 * metrics[3] = sum |f| (NC:744, debug history)
 * The two sums are reduced in a fixed order (deterministic).  ws: tnml_act_lossder_workspace_bytes(Ns). */
int64_t tnml_act_lossder_workspace_bytes(int64_t Ns);
int tnml_act_lossder(const void* f, const int32_t* y, const void* phi_p, const void* phi_q, void* q, void* pp,
                     void* metrics, void* ws, int64_t Ns, int32_t L, int32_t act, int32_t loss, double T,
                     int32_t dtype, tnml_stream_t stream);

/* Standalone forms on the reference's own (L, Ns) layout, for the public methods Network.apply_act_func (NC:767-796)
 * and Network.compute_loss_derivate (NC:800-835); y is the dense (one-hot) target the reference passes. */
int tnml_apply_act(const void* f, void* out, int64_t Ns, int32_t L, int32_t act, double T, int32_t dtype,
                   tnml_stream_t stream);
int tnml_loss_derivative(const void* fa, const void* y, void* out, int64_t Ns, int32_t L, int32_t act, int32_t loss,
                         double T, int32_t dtype, tnml_stream_t stream);

/* ---- a10: gradient, a K = Ns tensor-core reduction ---------------------------------------------------
 * dB[a][sigma][l][tau][c] = sum_b q[b][l][sigma,tau] L[b][a] R[b][c]          NC:625-646 + NC:710
 * Split-K over samples with a static decomposition and a fixed-order second stage => bitwise reproducible.
 * ws: tnml_grad_workspace_bytes(...). */
int64_t tnml_grad_workspace_bytes(int64_t Ns, int32_t Dl, int32_t Dr, int32_t L);
int tnml_grad(const void* q, const void* Lenv, const void* Renv, void* dB, void* ws, int64_t Ns, int32_t Dl, int32_t Dr,
              int32_t L, int32_t dtype, tnml_stream_t stream);

/* ---- a9: bond tensor formation and projection ----------------------------------------------------------
 * tnml_gemm: C = alpha * op(A) * op(B) + beta * C, row-major, used for B = A_p . A_q (NC:484), the L2 term and the
 * norm environments (NC:966-1179); transX != 0 means the stored matrix is the transpose.
 * tnml_project: f[b][l] = sum B[a][sigma][l][tau][c] L[b][a] pp[b][sigma,tau] R[b][c]     NC:494-523 */
int tnml_gemm(int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, double alpha, const void* A, int32_t lda,
              const void* B, int32_t ldb, double beta, void* C, int32_t ldc, int32_t dtype, tnml_stream_t stream);
int64_t tnml_project_workspace_bytes(int64_t Ns, int32_t Dl, int32_t Dr, int32_t L);
/* max_ctas: 0 = fill all SMs; otherwise cap the grid (the caller runs the SVD split beside the projection and
 * leaves whole GPCs free for its thread-block clusters). */
int tnml_project(const void* B, const void* pp, const void* Lenv, const void* Renv, void* f, void* ws, int64_t Ns,
                 int32_t Dl, int32_t Dr, int32_t L, int32_t max_ctas, int32_t dtype, tnml_stream_t stream);

/* ---- a10/a14: regularisation, clipping, update --------------------------------------------------------
 * tnml_l2_term     : G = E_L . B . E_R (the derivative of the squared norm w.r.t. B)        NC:1129-1135
 *                    EL [Dl][Dl], ER [Dr][Dr] norm environments; ws: Dl*4*L*Dr elements.  Independent of the
 *                    gradient, so the caller may run it on another stream while tnml_grad runs.
 * tnml_bond_update : reg = L2_flag ? 2 wd G : wd B                                       NC:728-734, NC:1176
 *                    dB -= reg ; if sum|dB| > sum|B| : dB /= (sum|dB| / sum|B|) ; B' = B + lr dB     NC:755-761
 * stats[0..7] = { sum|B|, sum|dB| (after reg, before clip), wd*<B, G> (0 if !L2_flag), clipped?,
 *                 mean|B|, mean|dB|, mean|reg|, 0 }                       (NC:741-747 debug history)
 * G is ignored when !L2_flag.  Bnew may alias neither B nor dB.  All sums are fixed-order (deterministic). */
int tnml_l2_term(const void* B, const void* EL, const void* ER, void* G, void* ws, int32_t Dl, int32_t Dr, int32_t L,
                 int32_t dtype, tnml_stream_t stream);
int64_t tnml_bond_update_workspace_bytes(int32_t Dl, int32_t Dr, int32_t L);
int tnml_bond_update(const void* B, const void* dB, const void* G, void* Bnew, void* stats, void* ws, int32_t Dl,
                     int32_t Dr, int32_t L, double lr, double wd, int32_t L2_flag, int32_t dtype, tnml_stream_t stream);

/* ---- a14: norm environments ----------------------------------------------------------------------------
 * right-moving: Eout[m][m'] = sum_{a,a',s} Ein[a][a'] A[a][s][m] A[a'][s][m']          NC:1004-1029
 * left-moving : Eout[a][a'] = sum_{c,c',s} A[a][s][c] A[a'][s][c'] Ein[c][c']          NC:1035-1061
 * ws: 2*Dl*Dr elements. */
int tnml_norm_env_step(const void* Ein, const void* site, void* Eout, void* ws, int32_t Dl, int32_t Dr,
                       int32_t left_moving, int32_t dtype, tnml_stream_t stream);

/* ---- a13: SVD split by one-sided Jacobi -----------------------------------------------------------------
 * Mx = B' viewed as R x C (right sweep: R = 2 Dl, C = 2 L Dr; left sweep: R = 2 Dl L, C = 2 Dr)   NC:528-560
 * U S Vh = svd(Mx); keep m; left site = U[:, :m] sqrt(S), right site = sqrt(S) Vh[:m]     NC:887-925, NC:947-960
 * Implementation: Gram matrix of the short side, one-sided (Hestenes) Jacobi on it inside one CTA, a second
 * pass on the rotated matrix to restore full accuracy for small singular values, then the two factors are
 * written straight into the destination site layouts:
 *   right sweep: site_p[a][s][m] (plain),      site_q[m][tau][l][c] (label site, right-sweep layout)
 *   left  sweep: site_p[a][l][s][m] (label site, left-sweep layout), site_q[m][tau][c] (plain)
 * svals must hold min(R, C) + 2 values: all singular values, descending, then two diagnostics (Jacobi sweeps used by
 * the first and second pass).  m is chosen by the caller (truncation rule).
 * refine: 0 = single Gram pass (small singular values only accurate to sqrt(eps) sigma_max); 1 = second pass unless the
 * first pass finds sigma_min/sigma_max > 3e-4 (then it is provably unnecessary at the 1e-11 level); 2 = always;
 * 3 = like 1, but when all m kept singular values lie in the accurate leading block the second pass -- which then
 * only improves the REPORTED values of the discarded tail, not the factors -- is left to tnml_svd_split_tail, which the
 * caller enqueues later on any stream ordered after this call (off the critical path of the sweep).  Bnew and the
 * workspace must stay untouched until that call has run; svals is complete only after it. */
int64_t tnml_svd_split_workspace_bytes(int32_t Dl, int32_t Dr, int32_t L, int32_t left_dir);
int tnml_svd_split(const void* Bnew, void* site_p, void* site_q, void* svals, void* ws, int32_t Dl, int32_t Dr,
                   int32_t L, int32_t m, int32_t left_dir, int32_t refine, int32_t dtype, tnml_stream_t stream);

/* tnml_svd_split_ev: the same call with a cudaEvent_t (as void*, may be NULL) that is recorded on `stream` once the Gram
 * matrix is complete, right before the Cholesky kernel that also reserves the SMs of the sweeps' thread-block cluster:
 * a caller that runs the projection on another stream makes it wait for this event, so that the projection cannot fill
 * those SMs first.  For sizes that take another path the event is not recorded (create it in the recorded state). */
int tnml_svd_split_ev(const void* Bnew, void* site_p, void* site_q, void* svals, void* ws, int32_t Dl, int32_t Dr,
                      int32_t L, int32_t m, int32_t left_dir, int32_t refine, int32_t dtype, tnml_stream_t stream,
                      void* gram_done_event);
/* record: NULL = refine now (second pass on the small block, svals complete when this call has run); otherwise a device
 * buffer of tnml_svd_tail_record_bytes() bytes that only receives the small block's Gram matrix (no cluster kernel on
 * this path) -- tnml_svd_tail_batch then solves `nrec` consecutive records at once, one CTA each, and writes the tail
 * singular values into svals + i * svals_stride (doubles) for record i.  Same Dl, Dr, L, m, left_dir as the split. */
int tnml_svd_split_tail(const void* Bnew, void* svals, void* ws, void* record, int32_t Dl, int32_t Dr, int32_t L,
                        int32_t m, int32_t left_dir, int32_t dtype, tnml_stream_t stream);
/* Warm-started split (the replacement of np.linalg.svd at NC:887 for a bond that is visited again and again by the
 * sweeps of a training run).  `warm` is a per-(bond, direction) device buffer of tnml_svd_warm_bytes() bytes, zeroed by
 * the caller before its first use and owned by the caller between calls: every split leaves its short-side rotation
 * there.  fast != 0 (2 Dl = 2 Dr = 128 or the mirrored left-sweep shape, m = 64, refine = 3): the split first tries the
 * deflation path -- one subspace-iteration step from the previous visit's m dominant vectors, CholeskyQR, Rayleigh-Ritz
 * on the 64 x 64 projected Gram matrix (two-sided Jacobi in one CTA), a-posteriori residual / gap / conditioning
 * gates on the device -- and runs the ordinary pipeline (single-CTA form) only when a gate fails.  Results agree with
 * the cold split to rounding.  fast = 0 or warm = NULL: exactly tnml_svd_split_ev (plus the rotation left in `warm`).
 * With fast != 0 svals must hold min(R, C) + 14 values: svals[n] >= 100 marks a split that took the deflation path
 * (100 + Jacobi sweeps), svals[n+2] / svals[n+3] the refusal code and the deciding ratio otherwise, svals[n+4 .. n+13]
 * the phase clocks of the single-CTA kernel (diagnostics).
 * The tail call must receive the same warm / fast arguments. */
int64_t tnml_svd_warm_bytes(int32_t Dl, int32_t Dr, int32_t L, int32_t left_dir);
int tnml_svd_split_warm(const void* Bnew, void* site_p, void* site_q, void* svals, void* ws, void* warm, int32_t Dl,
                        int32_t Dr, int32_t L, int32_t m, int32_t left_dir, int32_t refine, int32_t fast, int32_t dtype,
                        tnml_stream_t stream, void* gram_done_event);
int tnml_svd_split_tail_warm(const void* Bnew, void* svals, void* ws, void* record, void* warm, int32_t Dl, int32_t Dr,
                             int32_t L, int32_t m, int32_t left_dir, int32_t fast, int32_t dtype, tnml_stream_t stream);
int64_t tnml_svd_tail_record_bytes(void);
int tnml_svd_tail_batch(void* recs, int32_t nrec, void* svals, int64_t svals_stride, int32_t dtype, tnml_stream_t stream);

/* General form for Network.tensor_svd on any 2-D Tensor (NC:839-962): Mx [R][C] -> US [R][m] = U sqrt(S),
 * SVh [m][C] = sqrt(S) Vh; svals as above (min(R,C) + 2 values). */
int64_t tnml_svd_workspace_bytes(int32_t R, int32_t C);
int tnml_svd(const void* Mx, void* US, void* SVh, void* svals, void* ws, int32_t R, int32_t C, int32_t m, int32_t refine,
             int32_t dtype, tnml_stream_t stream);

/* Weight operands of the TNML_F32 variant.
 * tnml_convert_f32      : plain FP64 -> FP32 copy (label site for tnml_site_predict).
 * tnml_site_weights_f32 : the W operand of tnml_env_advance(TNML_F32), K-major for the tcgen05 kernel:
 *                         Wt[sigma][m][k] (FP32) with W[k][sigma][m] as defined above, straight from the FP64 site
 *                         [Dl][2][Dr]: right-moving (K = Dl, M = Dr) Wt[sigma][c][a]; left-moving (K = Dr, M = Dl)
 *                         Wt[sigma][a][c].  (kind::tf32 only multiplies K-major shared-memory operands correctly.) */
int tnml_convert_f32(const void* src_f64, void* dst_f32, int64_t n, tnml_stream_t stream);
int tnml_site_weights_f32(const void* site_f64, void* Wt_f32, int32_t Dl, int32_t Dr, int32_t left_moving,
                          tnml_stream_t stream);

/* ---- label-site layout change between sweep directions: [a][s][l][c] <-> [a][l][s][c] ------------------- */
int tnml_label_site_swap(const void* in, void* out, int32_t Dl, int32_t Dr, int32_t L, int32_t to_left_layout,
                         int32_t dtype, tnml_stream_t stream);

/* ---- a3: generic named-axis pair contraction (CLT:10-87) -------------------------------------------------
 * out[u1][u2][c] = sum_k T1[u1][c][k] * T2[u2][c][k]   (operands already permuted to (unique, common, contracted)
 * order, which is what CLT:51-75 does before the broadcast multiply of CLT:81 and the sums of CLT:82-84). */
int tnml_contract(const void* T1, const void* T2, void* out, int64_t U1, int64_t U2, int64_t Cc, int64_t Kc,
                  int32_t dtype, tnml_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TNML_H */
