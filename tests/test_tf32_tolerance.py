"""CPU: where the stated tolerance of the FP32/TF32 variant (tests/test_gpu_f32.py, TOL_TF32 = 4e-3) comes from.

The tcgen05 kernels multiply operands truncated to TF32 (10 explicit mantissa bits) and accumulate in FP32.  Emulating
exactly that on the oracle's contractions (NumPy; truncation by masking the low 13 mantissa bits, FP32 accumulation)
gives the error level the GPU tests must allow for -- and shows that the bound is neither vacuous nor too tight."""
import numpy as np

from oracle import mps_oracle as O


def tf32(a):
    """Round-toward-zero to TF32, as the tensor core does with FP32 operands in shared memory."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def rel(a, b):
    return float(np.abs(a.astype(np.float64) - b).max() / np.abs(b).max())


def _inputs(Ns, Dl, Dr, nl, seed):
    rng = np.random.default_rng(seed)
    Le, Re = rng.standard_normal((Ns, Dl)), rng.standard_normal((Ns, Dr))
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    g = rng.standard_normal((Ns, nl))
    return Le, Re, pa, pb, g


def test_tf32_emulated_gradient_error_level():
    # dB = sum_b (g L) (x) (pp R): both Khatri-Rao operands are formed in FP32, truncated by the MMA, summed in FP32
    Ns, Dl, Dr, nl = 3000, 64, 64, 4
    Le, Re, pa, pb, g = _inputs(Ns, Dl, Dr, nl, 4)
    want = O.gradient(g, Le, pa, pb, Re)
    f = np.float32
    pp = (pa[:, :, None] * pb[:, None, :]).reshape(Ns, 4).astype(f)
    A = tf32((pp[:, :, None] * Re.astype(f)[:, None, :]).reshape(Ns, 4 * Dr))          # (Ns, (st, c))
    B = tf32((g.astype(f)[:, :, None] * Le.astype(f)[:, None, :]).reshape(Ns, nl * Dl))  # (Ns, (l, a))
    D = (A.T @ B).reshape(2, 2, Dr, nl, Dl)                                              # FP32 accumulate
    got = np.transpose(D, (4, 0, 3, 1, 2))                                               # -> (a, s, l, t, c)
    e = rel(got, want)
    assert 1e-5 < e < 4e-3, e


def test_tf32_emulated_projection_and_environment_error_level():
    Ns, Dl, Dr, nl = 2000, 64, 64, 4
    Le, Re, pa, pb, _ = _inputs(Ns, Dl, Dr, nl, 5)
    rng = np.random.default_rng(6)
    Bt = rng.standard_normal((Dl, 2, nl, 2, Dr))
    want = O.project(Bt, Le, pa, pb, Re)
    f = np.float32
    T = tf32(Le.astype(f)) @ tf32(Bt.astype(f).reshape(Dl, -1))                          # (Ns, (s, l, t, c)), FP32 sums
    T = T.reshape(Ns, 2, nl, 2, Dr)
    pp = (pa[:, :, None] * pb[:, None, :]).astype(f)                                     # (Ns, s, t)
    got = np.einsum("bsltc,bst,bc->bl", T, pp, Re.astype(f))                             # epilogue in FP32
    e = rel(got, want)
    assert 1e-5 < e < 4e-3, e
    A = rng.standard_normal((Dl, 2, Dr))
    want = O.env_advance_right(Le, pa, A)
    G = tf32(Le.astype(f)) @ tf32(A.astype(f).reshape(Dl, 2 * Dr))
    got = pa.astype(f)[:, 0:1] * G[:, :Dr] + pa.astype(f)[:, 1:2] * G[:, Dr:]
    e = rel(got, want)
    assert 1e-5 < e < 4e-3, e
