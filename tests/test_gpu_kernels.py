"""GPU: every C-ABI entry point of libtnml.so against the oracle, on seeded inputs.  FP64 tolerance: 1e-12
relative to the largest magnitude of the expected result unless a test states otherwise (bitwise where the
operation is a pure data movement)."""
import ctypes

import numpy as np
import pytest

from oracle import mps_oracle as O

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tensornetworkforml_b200 import _lib
    _lib.lib()
    return _lib


_KEEP = []     # device tensors must outlive the asynchronous call that reads them through a raw pointer


@pytest.fixture(autouse=True)
def _release_device_tensors():
    yield
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    _KEEP.clear()


def dev(a, dtype=torch.float64):
    t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)
    _KEEP.append(t)
    return t


def empty(*shape):
    return torch.empty(shape, dtype=torch.float64, device="cuda")


def st():
    return torch.cuda.current_stream().cuda_stream


def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    return float(np.abs(a.reshape(b.shape) - b).max() / max(np.abs(b).max(), 1e-300))


def ws_for(L, name, *args):
    n = getattr(L.lib(), name)(*args)
    return torch.empty(max(1, (n + 7) // 8), dtype=torch.float64, device="cuda")


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Ns,S", [(5, 3), (100, 64), (257, 33)])
def test_feature_map_and_pack(L, Ns, S):
    rng = np.random.default_rng(0)
    x = rng.random((Ns, S))
    phi = empty(S, Ns, 2)
    L.call("tnml_feature_map", dev(x).data_ptr(), phi.data_ptr(), Ns, S, L.F64, st())
    want = np.transpose(O.feature_map(x), (1, 0, 2))
    assert rel(phi, want) < 1e-15
    X = O.feature_map(x)
    phi2 = empty(S, Ns, 2)
    L.call("tnml_pack_features", dev(X).data_ptr(), phi2.data_ptr(), Ns, S, L.F64, st())
    assert np.array_equal(phi2.cpu().numpy(), want)          # pure data movement: bitwise


@pytest.mark.parametrize("Ns,K,M", [(77, 3, 5), (1000, 64, 64), (50, 1, 2), (333, 10, 10), (64, 2, 1), (130, 40, 70),
                                    (5000, 32, 64), (777, 64, 32), (300, 16, 2), (260, 4, 64), (20011, 64, 64),
                                    (4097, 8, 20)])
def test_env_advance_both_directions(L, Ns, K, M):
    rng = np.random.default_rng(1)
    E = rng.standard_normal((Ns, K))
    phi = O.feature_map(rng.random((Ns,)))
    A = rng.standard_normal((K, 2, M))
    out = empty(Ns, M)
    L.call("tnml_env_advance", dev(E).data_ptr(), dev(phi).data_ptr(), dev(A).data_ptr(), out.data_ptr(), Ns, K, M,
           L.F64, st())
    assert rel(out, O.env_advance_right(E, phi, A)) < TOL
    # left-moving: site (M, 2, K) seen from the right
    A2 = rng.standard_normal((M, 2, K))
    wt = empty(K, 2, M)
    L.call("tnml_site_transpose", dev(A2).data_ptr(), wt.data_ptr(), M, K, L.F64, st())
    assert np.array_equal(wt.cpu().numpy(), np.transpose(A2, (2, 1, 0)))
    L.call("tnml_env_advance", dev(E).data_ptr(), dev(phi).data_ptr(), wt.data_ptr(), out.data_ptr(), Ns, K, M, L.F64,
           st())
    assert rel(out, O.env_advance_left(E, phi, A2)) < TOL


@pytest.mark.parametrize("Ns,Dl,Dr,nl", [(40, 1, 7, 2), (100, 5, 1, 10), (33, 3, 4, 3)])
def test_site_predict(L, Ns, Dl, Dr, nl):
    rng = np.random.default_rng(2)
    Le, Re = rng.standard_normal((Ns, Dl)), rng.standard_normal((Ns, Dr))
    phi = O.feature_map(rng.random((Ns,)))
    A = rng.standard_normal((Dl, 2, nl, Dr))
    f = empty(Ns, nl)
    L.call("tnml_site_predict", dev(Le).data_ptr(), dev(phi).data_ptr(), dev(A).data_ptr(), dev(Re).data_ptr(),
           f.data_ptr(), Ns, Dl, Dr, nl, L.F64, st())
    assert rel(f, O.site_predict(Le, phi, A, Re)) < TOL


@pytest.mark.parametrize("act", O.ACT_FNS)
@pytest.mark.parametrize("loss", O.LOSS_FNS)
def test_act_lossder_metrics(L, act, loss):
    rng = np.random.default_rng(3)
    Ns, nl, T = 1000, 10, 0.1
    f = rng.standard_normal((Ns, nl)) * 0.2
    if act == "linear" and loss != "MSE":
        f = np.abs(f) + 0.1
    y = rng.integers(0, nl, Ns)
    y1h = np.eye(nl)[y]
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    q, pp, met = empty(Ns, nl, 4), empty(Ns, 4), empty(4)
    ws = ws_for(L, "tnml_act_lossder_workspace_bytes", Ns)
    L.call("tnml_act_lossder", dev(f).data_ptr(), dev(y, torch.int32).data_ptr(), dev(pa).data_ptr(), dev(pb).data_ptr(),
           q.data_ptr(), pp.data_ptr(), met.data_ptr(), ws.data_ptr(), Ns, nl, L.ACT[act], L.LOSS[loss], T, L.F64, st())
    fa = O.apply_act(f, act, T)
    g = O.loss_derivative(fa, y1h, act, loss, T)
    w = (pa[:, :, None] * pb[:, None, :]).reshape(Ns, 4)
    assert rel(pp, w) < 1e-15
    assert rel(q, g[:, :, None] * w[:, None, :]) < TOL
    acc, mae = O.metrics(fa, y1h)
    m = met.cpu().numpy()
    assert m[0] == round(acc * Ns)
    assert abs(m[1] / (Ns * nl) - mae) < 1e-13
    assert m[2] == Ns and abs(m[3] - np.abs(f).sum()) < 1e-11 * np.abs(f).sum()      # NC:744 (debug history)


GRAD_SHAPES = [(50, 1, 5, 2), (200, 4, 4, 3), (3000, 64, 64, 10), (777, 10, 2, 2), (100, 70, 3, 2), (90, 3, 130, 2),
               (4100, 16, 32, 10)]


def _grad_inputs(Ns, Dl, Dr, nl, seed=4):
    rng = np.random.default_rng(seed)
    Le, Re = rng.standard_normal((Ns, Dl)), rng.standard_normal((Ns, Dr))
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    g = rng.standard_normal((Ns, nl))
    w = (pa[:, :, None] * pb[:, None, :]).reshape(Ns, 4)
    return Le, Re, pa, pb, g, w


@pytest.mark.parametrize("Ns,Dl,Dr,nl", GRAD_SHAPES)
def test_gradient(L, Ns, Dl, Dr, nl):
    Le, Re, pa, pb, g, w = _grad_inputs(Ns, Dl, Dr, nl)
    q = g[:, :, None] * w[:, None, :]
    dB = empty(Dl, 2, nl, 2, Dr)
    ws = ws_for(L, "tnml_grad_workspace_bytes", Ns, Dl, Dr, nl)
    args = (dev(q), dev(Le), dev(Re))
    L.call("tnml_grad", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), dB.data_ptr(), ws.data_ptr(), Ns,
           Dl, Dr, nl, L.F64, st())
    want = O.gradient(g, Le, pa, pb, Re)
    assert rel(dB, want) < TOL
    # determinism: bitwise identical on a second run
    dB2 = empty(Dl, 2, nl, 2, Dr)
    L.call("tnml_grad", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), dB2.data_ptr(), ws.data_ptr(), Ns,
           Dl, Dr, nl, L.F64, st())
    assert torch.equal(dB, dB2)


@pytest.mark.parametrize("Ns,Dl,Dr,nl", GRAD_SHAPES)
def test_projection(L, Ns, Dl, Dr, nl):
    Le, Re, pa, pb, g, w = _grad_inputs(Ns, Dl, Dr, nl, seed=5)
    rng = np.random.default_rng(6)
    B = rng.standard_normal((Dl, 2, nl, 2, Dr))
    f = empty(Ns, nl)
    ws = ws_for(L, "tnml_project_workspace_bytes", Ns, Dl, Dr, nl)
    L.call("tnml_project", dev(B).data_ptr(), dev(w).data_ptr(), dev(Le).data_ptr(), dev(Re).data_ptr(), f.data_ptr(),
           ws.data_ptr(), Ns, Dl, Dr, nl, 0, L.F64, st())
    want = O.project(B, Le, pa, pb, Re)
    assert rel(f, want) < TOL
    f2 = empty(Ns, nl)                         # capped grid (fewer sample splits): same result
    L.call("tnml_project", _KEEP[-4].data_ptr(), _KEEP[-3].data_ptr(), _KEEP[-2].data_ptr(), _KEEP[-1].data_ptr(),
           f2.data_ptr(), ws.data_ptr(), Ns, Dl, Dr, nl, 37, L.F64, st())
    assert rel(f2, want) < TOL


@pytest.mark.parametrize("tA,tB", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(5, 7, 3), (128, 1280, 64), (70, 65, 130)])
def test_gemm(L, tA, tB, M, N, K):
    rng = np.random.default_rng(7)
    A = rng.standard_normal((K, M) if tA else (M, K))
    B = rng.standard_normal((N, K) if tB else (K, N))
    C0 = rng.standard_normal((M, N))
    C = dev(C0)
    L.call("tnml_gemm", tA, tB, M, N, K, 1.5, dev(A).data_ptr(), A.shape[1], dev(B).data_ptr(), B.shape[1], 0.5,
           C.data_ptr(), N, L.F64, st())
    want = 1.5 * (A.T if tA else A) @ (B.T if tB else B) + 0.5 * C0
    assert rel(C, want) < TOL


@pytest.mark.parametrize("Dl,Dr,nl,L2,scale", [(4, 6, 3, 1, 1.0), (4, 6, 3, 0, 1.0), (64, 64, 10, 1, 1.0),
                                               (2, 2, 2, 1, 1e3), (1, 8, 2, 1, 1e3), (8, 1, 2, 0, 1e3)])
def test_bond_update(L, Dl, Dr, nl, L2, scale):
    rng = np.random.default_rng(8)
    B = rng.standard_normal((Dl, 2, nl, 2, Dr))
    dB = rng.standard_normal((Dl, 2, nl, 2, Dr)) * scale        # scale=1e3 forces the clipping branch
    a, c = rng.standard_normal((Dl, Dl)), rng.standard_normal((Dr, Dr))
    EL, ER = a @ a.T, c @ c.T
    lr, wd = 0.05, 0.3
    Bn, stats = empty(Dl, 2, nl, 2, Dr), empty(8)
    ws = ws_for(L, "tnml_bond_update_workspace_bytes", Dl, Dr, nl)
    Bd, Gd, ws2 = dev(B), empty(Dl, 2, nl, 2, Dr), empty(Dl * 4 * nl * Dr)
    if L2:
        L.call("tnml_l2_term", Bd.data_ptr(), dev(EL).data_ptr(), dev(ER).data_ptr(), Gd.data_ptr(), ws2.data_ptr(), Dl,
               Dr, nl, L.F64, st())
        assert rel(Gd, np.einsum("xa,asltc,cy->xslty", EL, B, ER)) < TOL
    L.call("tnml_bond_update", Bd.data_ptr(), dev(dB).data_ptr(), Gd.data_ptr() if L2 else None, Bn.data_ptr(),
           stats.data_ptr(), ws.data_ptr(), Dl, Dr, nl, lr, wd, L2, L.F64, st())
    if L2:
        loss, g = O.l2_term(B, EL, ER, wd)
        d = dB - g
    else:
        loss, d = 0.0, dB - wd * B
    want = O.clip_and_update(B, d, lr)
    assert rel(Bn, want) < TOL
    s = stats.cpu().numpy()
    assert abs(s[0] - np.abs(B).sum()) < 1e-11 * np.abs(B).sum()
    assert abs(s[1] - np.abs(d).sum()) < 1e-11 * np.abs(d).sum()
    assert abs(s[2] - loss) <= 1e-11 * abs(loss)
    assert s[3] == float(np.abs(d).sum() > np.abs(B).sum())
    reg = g if L2 else wd * B
    assert abs(s[6] - np.abs(reg).mean()) < 1e-11 * np.abs(reg).mean()      # NC:747 (debug history)


@pytest.mark.parametrize("Dl,Dr", [(3, 5), (64, 64), (1, 4), (6, 1), (70, 20)])
def test_norm_env_step(L, Dl, Dr):
    rng = np.random.default_rng(9)
    A = rng.standard_normal((Dl, 2, Dr))
    a, c = rng.standard_normal((Dl, Dl)), rng.standard_normal((Dr, Dr))
    EL, ER = a @ a.T, c @ c.T
    ws = empty(2 * Dl * Dr)
    out = empty(Dr, Dr)
    L.call("tnml_norm_env_step", dev(EL).data_ptr(), dev(A).data_ptr(), out.data_ptr(), ws.data_ptr(), Dl, Dr, 0, L.F64,
           st())
    assert rel(out, O.norm_env_right_step(EL, A)) < TOL
    out = empty(Dl, Dl)
    L.call("tnml_norm_env_step", dev(ER).data_ptr(), dev(A).data_ptr(), out.data_ptr(), ws.data_ptr(), Dl, Dr, 1, L.F64,
           st())
    assert rel(out, O.norm_env_left_step(ER, A)) < TOL


def _run_svd(L, B, left_dir, m, refine=2, tail=False, before_tail=None):
    Dl, _, nl, _, Dr = B.shape
    site_p = empty(Dl * 2 * m * (nl if left_dir else 1))
    site_q = empty(m * 2 * Dr * (1 if left_dir else nl))
    sv = torch.full((4 * max(Dl, Dr, nl) * 2,), float("nan"), dtype=torch.float64, device="cuda")
    ws = ws_for(L, "tnml_svd_split_workspace_bytes", Dl, Dr, nl, left_dir)
    Bd = dev(B)
    L.call("tnml_svd_split", Bd.data_ptr(), site_p.data_ptr(), site_q.data_ptr(), sv.data_ptr(), ws.data_ptr(), Dl,
           Dr, nl, m, left_dir, refine, L.F64, st())
    n = min(2 * Dl * (nl if left_dir else 1), 2 * Dr * (1 if left_dir else nl))
    if tail:          # refine = 3: the deferred refinement of the discarded tail's singular values
        if before_tail is not None:
            before_tail.append((sv.cpu().numpy()[:n].copy(), site_p.cpu().numpy().copy(), site_q.cpu().numpy().copy()))
        if tail == "batch":   # record now, solve later (here: a batch of one)
            rec = torch.zeros(L.lib().tnml_svd_tail_record_bytes() // 8, dtype=torch.float64, device="cuda")
            L.call("tnml_svd_split_tail", Bd.data_ptr(), sv.data_ptr(), ws.data_ptr(), rec.data_ptr(), Dl, Dr, nl, m,
                   left_dir, L.F64, st())
            L.call("tnml_svd_tail_batch", rec.data_ptr(), 1, sv.data_ptr(), sv.numel(), L.F64, st())
        else:
            L.call("tnml_svd_split_tail", Bd.data_ptr(), sv.data_ptr(), ws.data_ptr(), None, Dl, Dr, nl, m, left_dir,
                   L.F64, st())
    sp, sq = site_p.cpu().numpy(), site_q.cpu().numpy()
    if not left_dir:
        Ap, Aq = sp.reshape(Dl, 2, m), sq.reshape(m, 2, nl, Dr)
        prod = np.einsum("asm,mtlc->asltc", Ap, Aq)
    else:
        Ap, Aq = sp.reshape(Dl, nl, 2, m), sq.reshape(m, 2, Dr)
        prod = np.einsum("alsm,mtc->asltc", Ap, Aq)
    return sv.cpu().numpy()[:n], prod, Ap, Aq


@pytest.mark.parametrize("refine", [2, 1, 0])
@pytest.mark.parametrize("Dl,Dr,nl,left_dir,m", [(4, 4, 3, 0, 4), (4, 4, 3, 1, 4), (64, 64, 10, 0, 64),
                                                 (64, 64, 10, 1, 64), (1, 5, 2, 0, 2), (5, 1, 2, 1, 2), (4, 1, 2, 0, 4),
                                                 (1, 3, 2, 1, 4), (2, 2, 2, 0, 2), (7, 5, 3, 0, 9), (3, 9, 2, 1, 5),
                                                 (128, 128, 3, 0, 128), (100, 90, 2, 1, 77), (256, 256, 2, 0, 256),
                                                 (200, 256, 2, 1, 130)])
def test_svd_split(L, Dl, Dr, nl, left_dir, m, refine):
    rng = np.random.default_rng(10)
    B = rng.standard_normal((Dl, 2, nl, 2, Dr))
    Mx = B.reshape(Dl * 2, -1) if not left_dir else B.reshape(Dl * 2 * nl, 2 * Dr)
    U, S, Vh = np.linalg.svd(Mx, full_matrices=False)
    m = min(m, len(S))
    if max(Dl, Dr) > 64 and refine != 2:
        pytest.skip("large sizes are exercised with the full two-pass variant only")
    sv, prod, Ap, Aq = _run_svd(L, B, left_dir, m, refine)
    assert np.abs(sv - S).max() / S.max() < 1e-12
    want = ((U[:, :m] * S[:m]) @ Vh[:m]).reshape(B.shape)
    assert np.abs(prod - want).max() / np.abs(want).max() < 1e-11
    # sqrt(S) sits on both factors (NC:912-915): the Gram matrix of each factor over its outer indices is diag(S)
    Gp = np.tensordot(Ap, Ap, axes=(list(range(Ap.ndim - 1)), list(range(Ap.ndim - 1))))
    assert np.abs(Gp - np.diag(S[:m])).max() / S.max() < 1e-10


def test_svd_small_singular_values_need_the_second_pass(L):
    """Graded spectrum 1 .. 1e-9: absolute accuracy eps*sigma_max on every singular value with refine=1."""
    rng = np.random.default_rng(11)
    Dl, Dr, nl = 8, 8, 2
    R, C = 2 * Dl, 2 * nl * Dr
    Q1, _ = np.linalg.qr(rng.standard_normal((R, R)))
    Q2, _ = np.linalg.qr(rng.standard_normal((C, R)))
    S = np.logspace(0, -9, R)
    B = ((Q1 * S) @ Q2.T).reshape(Dl, 2, nl, 2, Dr)
    sv, _, _, _ = _run_svd(L, B, 0, R, refine=1)
    assert np.abs(sv - S).max() < 1e-13
    sv0, _, _, _ = _run_svd(L, B, 0, R, refine=0)
    assert np.abs(sv0 - S).max() < 1e-6          # single Gram pass: sqrt(eps)-level only


@pytest.mark.parametrize("Dl,Dr,nl,left_dir", [(40, 40, 2, 0), (64, 64, 3, 0), (30, 50, 2, 1), (100, 100, 2, 0),
                                               (16, 16, 2, 0), (32, 32, 3, 1), (8, 20, 2, 0)])
def test_svd_two_scale_spectrum_small_block_refinement(L, Dl, Dr, nl, left_dir):
    """The spectrum of a trained bond tensor: half the singular values O(1), the other half 1e-5 .. 1e-10 (what is
    discarded).  refine=1 decomposes only the small block in its second pass; every singular value must still be
    accurate to ~eps * sigma_max, and the kept part of the split exact."""
    rng = np.random.default_rng(15)
    R, C = (2 * Dl, 2 * nl * Dr) if not left_dir else (2 * Dl * nl, 2 * Dr)
    n = min(R, C)
    Q1, _ = np.linalg.qr(rng.standard_normal((R, n)))
    Q2, _ = np.linalg.qr(rng.standard_normal((C, n)))
    S = np.concatenate([np.logspace(0, -1, n // 2), np.logspace(-5, -10, n - n // 2)])
    Mx = (Q1 * S) @ Q2.T
    B = Mx.reshape(Dl, 2, nl, 2, Dr)
    m = n // 2
    want = ((Q1[:, :m] * S[:m]) @ Q2[:, :m].T).reshape(B.shape)
    for refine in (1, 2):
        sv, prod, _, _ = _run_svd(L, B, left_dir, m, refine)
        assert np.abs(sv - S).max() < 2e-13, "refine=%d" % refine
        assert np.abs(prod - want).max() < 1e-11
    # refine = 3: all m kept singular values sit in the accurate leading block, so the second pass is deferred to
    # tnml_svd_split_tail.  The factors are final (and identical) before the tail call; the discarded singular values
    # reach full accuracy after it.
    snap = []
    sv3, prod3, Ap3, Aq3 = _run_svd(L, B, left_dir, m, 3, tail=True, before_tail=snap)
    sv_before, p_before, q_before = snap[0]
    assert np.array_equal(p_before, Ap3.reshape(-1)) and np.array_equal(q_before, Aq3.reshape(-1))
    assert np.abs(prod3 - want).max() < 1e-11
    assert np.abs(sv_before[:m] - S[:m]).max() < 2e-13
    assert np.abs(sv3 - S).max() < 2e-13
    # every path defers (cluster kernels for n > 64, the single-CTA kernel below): before the tail call the discarded
    # values are only as accurate as a single Gram pass leaves them
    assert np.abs(sv_before - S).max() > np.abs(sv3 - S).max()
    # the same through the recorded / batched form of the tail
    sv4, prod4, _, _ = _run_svd(L, B, left_dir, m, 3, tail="batch")
    assert np.abs(prod4 - want).max() < 1e-11
    assert np.abs(sv4 - S).max() < 2e-13


def test_svd_rank_deficient(L):
    rng = np.random.default_rng(12)
    Dl, Dr, nl, r = 6, 6, 2, 5
    R, C = 2 * Dl, 2 * nl * Dr
    Mx = rng.standard_normal((R, r)) @ rng.standard_normal((r, C))
    B = Mx.reshape(Dl, 2, nl, 2, Dr)
    S = np.linalg.svd(Mx, compute_uv=False)
    sv, prod, _, _ = _run_svd(L, B, 0, R, refine=1)
    assert np.abs(sv - S).max() / S.max() < 1e-12
    assert np.isfinite(prod).all()
    assert np.abs(prod - B).max() / np.abs(B).max() < 1e-11


@pytest.mark.parametrize("Dl,Dr,nl", [(3, 4, 5), (1, 2, 2), (8, 1, 10)])
def test_label_site_swap_round_trip(L, Dl, Dr, nl):
    rng = np.random.default_rng(13)
    A = rng.standard_normal((Dl, 2, nl, Dr))
    a, b, c = dev(A), empty(Dl, nl, 2, Dr), empty(Dl, 2, nl, Dr)
    L.call("tnml_label_site_swap", a.data_ptr(), b.data_ptr(), Dl, Dr, nl, 1, L.F64, st())
    assert np.array_equal(b.cpu().numpy(), np.transpose(A, (0, 2, 1, 3)))
    L.call("tnml_label_site_swap", b.data_ptr(), c.data_ptr(), Dl, Dr, nl, 0, L.F64, st())
    assert np.array_equal(c.cpu().numpy(), A)


def test_contract_matches_einsum(L):
    from tensornetworkforml_b200 import Tensor, contract
    rng = np.random.default_rng(14)
    a = Tensor(elem=rng.standard_normal((3, 5, 7)), axes_names=["left", "right", "b"])
    b = Tensor(elem=rng.standard_normal((7, 5, 4)), axes_names=["b", "left", "right"])
    want = np.einsum("lkb,bkr->lrb", a.elem, b.elem)
    out = contract(a, b, "right", "left", common="b")
    assert list(out.axes_names) == ["left", "right", "b"]
    assert np.abs(out.elem - want).max() < 1e-13
    # several contracted axes given as positions (the call pattern of NC:1027-1029)
    c = Tensor(elem=rng.standard_normal((2, 3, 4, 5)), axes_names=["x", "right", "R_2", "y"])
    d = Tensor(elem=rng.standard_normal((3, 4, 6)), axes_names=["left", "L_2", "z"])
    want = np.einsum("xrsy,rsz->xyz", c.elem, d.elem)
    out = contract(c, d, c.ax_to_index(["right", "R_2"]), d.ax_to_index(["left", "L_2"]))
    assert list(out.axes_names) == ["x", "y", "z"] and np.abs(out.elem - want).max() < 1e-13


def test_bad_arguments_are_rejected(L):
    lib = L.lib()
    assert lib.tnml_env_advance(None, None, None, None, 10, 4, 4, L.F64, None) == -1
    x = empty(4)
    # batch-independent entry points are FP64 only (tnml.h): the FP32/TF32 variant is refused there
    assert lib.tnml_site_transpose(x.data_ptr(), x.data_ptr(), 1, 2, L.F32, None) == -2
    assert lib.tnml_feature_map(x.data_ptr(), x.data_ptr(), 2, 2, 7, None) == -1
    assert b"invalid" in lib.tnml_error_string(-1)
