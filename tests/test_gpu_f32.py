"""GPU: the FP32 / TF32 variant (dtype TNML_F32) against the FP64 oracle.

Stated tolerances (relative to the largest magnitude of the expected result):
  * FP32 FMA kernels (ragged shapes, and every shape under TNML_F32_FORCE_SIMT=1): 2e-5 -- FP32 rounding of the inputs
    plus FP32 accumulation over K <= 8192 terms.
  * tcgen05 kernels (kind::tf32): the tensor core truncates both operands to 10 mantissa bits, so one product carries a
    relative error of up to 2^-10 ~ 1e-3; with random signs the sum over K terms stays at ~1e-3 of the result's scale.
    Bound used here: 4e-3.
  * a whole sweep (S = 20 sites, bond dimension 64): predictions within 2e-2 of the FP64 oracle, singular values within
    1e-2 of sigma_max, accuracy within 1 %.
The FP32 variant keeps site tensors, bond tensors, the gradient sum, clipping and the SVD split in FP64 (tnml.h).
"""
import contextlib
import io
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import mps_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL_FMA = 2e-5
TOL_TF32 = 4e-3


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tensornetworkforml_b200 import _lib
    _lib.lib()
    return _lib


_KEEP = []


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=torch.float32):
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)
    _KEEP.append(t)
    return t


def empty(*shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype, device="cuda")


def st():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a = a.cpu().numpy().astype(np.float64) if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


def ws_for(L, name, *args):
    n = getattr(L.lib(), name)(*args)
    return torch.empty(max(1, (n + 7) // 8), dtype=torch.float64, device="cuda")


def tol_for(*dims_tc):
    """tcgen05 path when every listed (dim, multiple) pair divides; FP32 FMA otherwise."""
    forced = os.environ.get("TNML_F32_FORCE_SIMT", "0") not in ("", "0")
    return TOL_TF32 if (not forced and all(d % m == 0 for d, m in dims_tc)) else TOL_FMA


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Ns,S", [(5, 3), (257, 33)])
def test_feature_map_and_pack_f32(L, Ns, S):
    rng = np.random.default_rng(0)
    x = rng.random((Ns, S))
    want = np.transpose(O.feature_map(x), (1, 0, 2))
    phi = empty(S, Ns, 2)
    L.call("tnml_feature_map", dev(x, torch.float64).data_ptr(), phi.data_ptr(), Ns, S, L.F32, st())
    assert np.array_equal(phi.cpu().numpy(), want.astype(np.float32))       # double evaluation, one rounding
    phi2 = empty(S, Ns, 2)
    L.call("tnml_pack_features", dev(O.feature_map(x), torch.float64).data_ptr(), phi2.data_ptr(), Ns, S, L.F32, st())
    assert np.array_equal(phi2.cpu().numpy(), want.astype(np.float32))


ENV_SHAPES = [(77, 3, 5), (1000, 64, 64), (333, 10, 10), (130, 40, 70), (5000, 64, 32), (300, 32, 64), (2000, 128, 128),
              (700, 256, 256), (129, 64, 64), (4097, 96, 160)]


@pytest.mark.parametrize("Ns,K,M", ENV_SHAPES)
def test_env_advance_f32(L, Ns, K, M):
    rng = np.random.default_rng(1)
    E = rng.standard_normal((Ns, K))
    phi = O.feature_map(rng.random((Ns,)))
    A = rng.standard_normal((K, 2, M))
    out, Wt = empty(Ns, M), empty(2, M, K)
    L.call("tnml_site_weights_f32", dev(A, torch.float64).data_ptr(), Wt.data_ptr(), K, M, 0, st())
    assert np.array_equal(Wt.cpu().numpy(), np.transpose(A, (1, 2, 0)).astype(np.float32))
    L.call("tnml_env_advance", dev(E).data_ptr(), dev(phi).data_ptr(), Wt.data_ptr(), out.data_ptr(), Ns, K, M, L.F32,
           st())
    torch.cuda.synchronize()
    assert rel(out, O.env_advance_right(E, phi, A)) < tol_for((K, 32), (M, 32))
    # left-moving: site (M, 2, K) seen from the right
    A2 = rng.standard_normal((M, 2, K))
    L.call("tnml_site_weights_f32", dev(A2, torch.float64).data_ptr(), Wt.data_ptr(), M, K, 1, st())
    assert np.array_equal(Wt.cpu().numpy(), np.transpose(A2, (1, 0, 2)).astype(np.float32))
    L.call("tnml_env_advance", dev(E).data_ptr(), dev(phi).data_ptr(), Wt.data_ptr(), out.data_ptr(), Ns, K, M, L.F32,
           st())
    torch.cuda.synchronize()
    assert rel(out, O.env_advance_left(E, phi, A2)) < tol_for((K, 32), (M, 32))


def test_convert_and_site_predict_f32(L):
    rng = np.random.default_rng(2)
    Ns, Dl, Dr, nl = 100, 5, 1, 10
    Le, Re = rng.standard_normal((Ns, Dl)), rng.standard_normal((Ns, Dr))
    phi = O.feature_map(rng.random((Ns,)))
    A = rng.standard_normal((Dl, 2, nl, Dr))
    A32 = empty(A.size)
    L.call("tnml_convert_f32", dev(A, torch.float64).data_ptr(), A32.data_ptr(), A.size, st())
    assert np.array_equal(A32.cpu().numpy(), A.astype(np.float32).reshape(-1))
    f = empty(Ns, nl)
    L.call("tnml_site_predict", dev(Le).data_ptr(), dev(phi).data_ptr(), A32.data_ptr(), dev(Re).data_ptr(),
           f.data_ptr(), Ns, Dl, Dr, nl, L.F32, st())
    assert rel(f, O.site_predict(Le, phi, A, Re)) < TOL_FMA


@pytest.mark.parametrize("act,loss", [("linear", "MSE"), ("softmax", "full_cross_ent"), ("sigmoid", "cross_entropy")])
def test_act_lossder_f32(L, act, loss):
    rng = np.random.default_rng(3)
    Ns, nl, T = 1001, 10, 0.1
    f = (rng.standard_normal((Ns, nl)) * 0.2).astype(np.float32).astype(np.float64)
    y = rng.integers(0, nl, Ns)
    y1h = np.eye(nl)[y]
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    g = empty(Ns * nl + 4 * Ns + 4)
    pp, met = empty(Ns, 4), empty(4, dtype=torch.float64)
    ws = ws_for(L, "tnml_act_lossder_workspace_bytes", Ns)
    L.call("tnml_act_lossder", dev(f).data_ptr(), dev(y, torch.int32).data_ptr(), dev(pa).data_ptr(), dev(pb).data_ptr(),
           g.data_ptr(), pp.data_ptr(), met.data_ptr(), ws.data_ptr(), Ns, nl, L.ACT[act], L.LOSS[loss], T, L.F32, st())
    fa = O.apply_act(f, act, T)
    want_g = O.loss_derivative(fa, y1h, act, loss, T)
    w = (pa[:, :, None] * pb[:, None, :]).reshape(Ns, 4)
    assert rel(pp, w) < 1e-6
    assert rel(g[:Ns * nl], want_g.reshape(-1)) < 1e-6
    off = (Ns * nl + 3) // 4 * 4
    assert torch.equal(g[off:off + 4 * Ns], pp.reshape(-1))                 # the appended copy of pp
    acc, mae = O.metrics(fa, y1h)
    m = met.cpu().numpy()
    assert m[0] == round(acc * Ns)
    assert abs(m[1] / (Ns * nl) - mae) < 1e-12


GRAD_SHAPES = [(50, 1, 5, 2), (200, 4, 4, 3), (3000, 64, 64, 10), (777, 10, 2, 2), (100, 70, 3, 2), (4100, 16, 32, 10),
               (9000, 64, 64, 2), (2500, 128, 64, 10), (2000, 64, 128, 3), (6001, 64, 64, 7)]


def _inputs(Ns, Dl, Dr, nl, seed=4):
    rng = np.random.default_rng(seed)
    Le, Re = rng.standard_normal((Ns, Dl)), rng.standard_normal((Ns, Dr))
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    g = rng.standard_normal((Ns, nl))
    w = (pa[:, :, None] * pb[:, None, :]).reshape(Ns, 4)
    return Le, Re, pa, pb, g, w


@pytest.mark.parametrize("Ns,Dl,Dr,nl", GRAD_SHAPES)
def test_gradient_f32(L, Ns, Dl, Dr, nl):
    Le, Re, pa, pb, g, w = _inputs(Ns, Dl, Dr, nl)
    off = (Ns * nl + 3) // 4 * 4
    q = np.zeros(off + 4 * Ns)
    q[:Ns * nl] = g.reshape(-1)
    q[off:] = w.reshape(-1)
    dB = empty(Dl, 2, nl, 2, Dr, dtype=torch.float64)
    ws = ws_for(L, "tnml_grad_workspace_bytes", Ns, Dl, Dr, nl)
    args = (dev(q), dev(Le), dev(Re))
    L.call("tnml_grad", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), dB.data_ptr(), ws.data_ptr(), Ns,
           Dl, Dr, nl, L.F32, st())
    torch.cuda.synchronize()
    want = O.gradient(g, Le, pa, pb, Re)
    assert rel(dB, want) < tol_for((Dl, 64), (Dr, 64))
    dB2 = empty(Dl, 2, nl, 2, Dr, dtype=torch.float64)
    L.call("tnml_grad", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), dB2.data_ptr(), ws.data_ptr(), Ns,
           Dl, Dr, nl, L.F32, st())
    assert torch.equal(dB, dB2)                                             # deterministic split-K


@pytest.mark.parametrize("Ns,Dl,Dr,nl", GRAD_SHAPES)
def test_projection_f32(L, Ns, Dl, Dr, nl):
    Le, Re, pa, pb, g, w = _inputs(Ns, Dl, Dr, nl, seed=5)
    rng = np.random.default_rng(6)
    B = rng.standard_normal((Dl, 2, nl, 2, Dr))
    f = empty(Ns, nl)
    ws = ws_for(L, "tnml_project_workspace_bytes", Ns, Dl, Dr, nl)
    args = (dev(B, torch.float64), dev(w), dev(Le), dev(Re))
    L.call("tnml_project", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(), f.data_ptr(),
           ws.data_ptr(), Ns, Dl, Dr, nl, 0, L.F32, st())
    torch.cuda.synchronize()
    want = O.project(B, Le, pa, pb, Re)
    tol = tol_for((Dl, 64), (Dr, 64))
    assert rel(f, want) < tol
    f2 = empty(Ns, nl)                         # capped grid (fewer sample splits): same result
    L.call("tnml_project", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(), f2.data_ptr(),
           ws.data_ptr(), Ns, Dl, Dr, nl, 37, L.F32, st())
    assert rel(f2, want) < tol


# ---------------------------------------------------------------------------------------------------
def _sweeps(dtype, S, Ns, nl, D, nsweeps, act, loss, seed=11):
    import tensornetworkforml_b200 as tn
    np.random.seed(seed)
    X = O.feature_map(np.random.random((Ns, S)))
    y = np.random.randint(0, nl, Ns)
    state = np.random.get_state()
    orc = O.OracleMPS.from_seed(S, D, nl, calibration_X=X, normalize=True, act_fn=act, loss_fn=loss, rule="fixed",
                                max_bond=D)
    np.random.set_state(state)
    with contextlib.redirect_stdout(io.StringIO()):
        net = tn.Network(N=S, M=D, L=nl, normalize=True, calibration_X=X, act_fn=act, loss_fn=loss, truncation="fixed",
                         max_bond=D, dtype=dtype)
    out = []
    for _ in range(nsweeps):
        fo, f = orc.forward(X), net.forward(X)
        left = orc.l_pos == S - 1
        d_fwd = float(np.abs(f.elem.T - fo).max() / np.abs(fo).max())
        fo = orc.sweep(y, fo, 1e-3, 1e-3, L2_flag=True, left_dir=left)
        f = net.sweep(X, y, f, 1e-3, 1e-3, L2_flag=True, left_dir=left)
        h = net.last_history
        oh = orc.hist[-(S - 1):]
        sv_dev = max(float(np.abs(h["svals"][i][:len(oh[i]["S"])] - oh[i]["S"]).max() / oh[i]["S"].max())
                     for i in range(S - 1))
        out.append(dict(fwd=d_fwd, swp=float(np.abs(f.elem.T - fo).max() / np.abs(fo).max()),
                        acc=float(abs(net.accuracy(X, y, f) - orc.accuracy(X, y, fo))), sv=sv_dev,
                        bonds=(net._eng.bond_dims(), orc.bond_dims())))
    return out


@pytest.mark.parametrize("act,loss", [("linear", "MSE"), ("softmax", "full_cross_ent")])
def test_sweeps_f32_vs_oracle(L, act, loss):
    """Two sweeps (right, left) of a 20-site, 4-label, D = 64 MPS in the FP32/TF32 variant stay within the stated
    tolerance of the FP64 oracle; the interior bonds use the tcgen05 kernels, the chain ends the FP32 FMA ones."""
    res = _sweeps("float32", S=20, Ns=2048, nl=4, D=64, nsweeps=2, act=act, loss=loss)
    for r in res:
        assert r["bonds"][0] == r["bonds"][1]
        assert r["fwd"] < 2e-2 and r["swp"] < 2e-2, res
        assert r["acc"] <= 0.01 and r["sv"] < 1e-2, res


def test_f32_simt_only_matches_fp32_tolerance(L):
    """The same kernels' tests with the tensor cores switched off (TNML_F32_FORCE_SIMT=1): every shape must then meet
    the FP32 FMA tolerance, which cross-checks the tcgen05 kernels' operand layouts against plain code."""
    env = dict(os.environ, TNML_F32_FORCE_SIMT="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", __file__, "-k",
                        "env_advance_f32 or gradient_f32 or projection_f32"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
