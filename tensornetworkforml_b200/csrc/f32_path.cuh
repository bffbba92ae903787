// Internal interface of the FP32 / TF32 variant (dtype == TNML_F32).  Per-sample arrays (phi, environments, f, g, pp)
// are FP32; site tensors, bond tensors and everything batch-independent stay FP64 (see tnml.h).  The tile-aligned
// contractions run on tcgen05 tensor cores (kind::tf32, FP32 accumulation in TMEM); ragged shapes fall back to FP32
// FMA kernels.  Every function enqueues on `st` and returns a TNML status.
#pragma once
#include "common.cuh"

namespace tnml {
namespace f32 {

int feature_map(const double* x, float* phi, int64_t Ns, int S, cudaStream_t st);
int pack_features(const double* X, float* phi, int64_t Ns, int S, cudaStream_t st);
int convert(const double* src, float* dst, int64_t n, cudaStream_t st);
int site_weights(const double* site, float* Wt, int Dl, int Dr, int left_moving, cudaStream_t st);
int env_advance(const float* E, const float* phi, const float* W, float* out, int64_t Ns, int K, int M, cudaStream_t st);
int site_predict(const float* Lenv, const float* phi, const float* A, const float* Renv, float* f, int64_t Ns, int Dl,
                 int Dr, int L, cudaStream_t st);
int act_lossder(const float* f, const int32_t* y, const float* phi_p, const float* phi_q, float* g, float* pp,
                double* metrics, double* ws, int64_t Ns, int L, int act, int loss, double T, cudaStream_t st);
int64_t grad_workspace_bytes(int64_t Ns, int Dl, int Dr, int L);
int grad(const float* g, const float* pp, const float* Lenv, const float* Renv, double* dB, void* ws, int64_t Ns, int Dl,
         int Dr, int L, cudaStream_t st);
int64_t project_workspace_bytes(int64_t Ns, int Dl, int Dr, int L);
int project(const double* B, const float* pp, const float* Lenv, const float* Renv, float* f, void* ws, int64_t Ns,
            int Dl, int Dr, int L, int max_ctas, cudaStream_t st);

// TNML_F32_FORCE_SIMT=1 keeps every contraction on the FP32 FMA kernels (A/B knob for the tests)
bool tensor_cores_enabled();

}  // namespace f32
}  // namespace tnml
