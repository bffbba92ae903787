"""GPU, BASELINE.json full sizes (config 3: Ns = 60 000 samples, bond dimension 64, 10 labels), where the oracle would
take minutes: size-independent properties of the hot path, all through the C ABI.

  * adjointness: the gradient kernel is the transpose of the projection kernel, so for any g and B
        sum_{b,l} g[b,l] * project(B)[b,l]  ==  sum_{a,s,l,t,c} gradient(g)[...] * B[...]
    -- one identity that ties k_grad, k_project and the q / pp operand conventions together at K = Ns = 60 000;
  * linearity of the projection in B, bitwise determinism of the K = Ns reduction;
  * environment advance: composing two sites equals advancing through the contracted pair;
  * SVD split: the two factors reproduce the best rank-m approximation (Eckart-Young: ||Mx - US.SVh||_F^2 equals
    the sum of the discarded sigma^2) and sqrt(S) sits on both;
  * a full sweep of the 196-site chain: finite, deterministic run to run, bond dimensions as the rule prescribes,
    and the running prediction equals a fresh forward() (the last split of a sweep is lossless).
"""
import contextlib
import io

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

NS, D, NL = 60000, 64, 10


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tensornetworkforml_b200 import _lib
    _lib.lib()
    return _lib


def st():
    return torch.cuda.current_stream().cuda_stream


def rnd(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g)


def ws_for(L, name, *args):
    return torch.empty(max(1, (getattr(L.lib(), name)(*args) + 7) // 8), dtype=torch.float64, device="cuda")


def test_gradient_is_the_adjoint_of_the_projection(L):
    Le, Re = rnd(NS, D, seed=1), rnd(NS, D, seed=2)
    ang_p, ang_q = torch.rand(NS, dtype=torch.float64, device="cuda"), torch.rand(NS, dtype=torch.float64, device="cuda")
    phi_p = torch.stack((torch.sin(ang_p), torch.cos(ang_p)), 1).contiguous()
    phi_q = torch.stack((torch.sin(ang_q), torch.cos(ang_q)), 1).contiguous()
    pp = (phi_p[:, :, None] * phi_q[:, None, :]).reshape(NS, 4).contiguous()
    g = rnd(NS, NL, seed=3)
    q = (g[:, :, None] * pp[:, None, :]).contiguous()
    B1, B2 = rnd(D, 2, NL, 2, D, seed=4), rnd(D, 2, NL, 2, D, seed=5)
    dB, dB2 = torch.empty_like(B1), torch.empty_like(B1)
    wg = ws_for(L, "tnml_grad_workspace_bytes", NS, D, D, NL)
    for out in (dB, dB2):
        L.call("tnml_grad", q.data_ptr(), Le.data_ptr(), Re.data_ptr(), out.data_ptr(), wg.data_ptr(), NS, D, D, NL,
               L.F64, st())
    assert torch.equal(dB, dB2)                                   # fixed-order split-K: bitwise reproducible
    wp = ws_for(L, "tnml_project_workspace_bytes", NS, D, D, NL)
    f = {}
    for name, B, cap in (("b1", B1, 0), ("b2", B2, 0), ("sum", B1 + B2, 0), ("b1_capped", B1, 110)):
        out = torch.empty(NS, NL, dtype=torch.float64, device="cuda")
        L.call("tnml_project", B.data_ptr(), pp.data_ptr(), Le.data_ptr(), Re.data_ptr(), out.data_ptr(), wp.data_ptr(),
               NS, D, D, NL, cap, L.F64, st())
        f[name] = out
    lhs = float((g * f["b1"]).sum())
    rhs = float((dB * B1).sum())
    scale = float((g.abs() * f["b1"].abs()).sum())
    assert abs(lhs - rhs) < 1e-12 * scale                         # <g, P B> == <P^T g, B>
    assert float((f["sum"] - f["b1"] - f["b2"]).abs().max()) < 1e-11 * float(f["sum"].abs().max())
    assert float((f["b1_capped"] - f["b1"]).abs().max()) < 1e-12 * float(f["b1"].abs().max())


def test_environment_advance_composes(L):
    E = rnd(NS, D, seed=6)
    ang = torch.rand(NS, 2, dtype=torch.float64, device="cuda")
    phi = [torch.stack((torch.sin(ang[:, i]), torch.cos(ang[:, i])), 1).contiguous() for i in range(2)]
    A1, A2 = rnd(D, 2, D, seed=7) / 8, rnd(D, 2, D, seed=8) / 8
    mid, out = torch.empty(NS, D, dtype=torch.float64, device="cuda"), torch.empty(NS, D, dtype=torch.float64, device="cuda")
    L.call("tnml_env_advance", E.data_ptr(), phi[0].data_ptr(), A1.data_ptr(), mid.data_ptr(), NS, D, D, L.F64, st())
    L.call("tnml_env_advance", mid.data_ptr(), phi[1].data_ptr(), A2.data_ptr(), out.data_ptr(), NS, D, D, L.F64, st())
    # reference: contract the pair first (a,s,t,c), then apply both feature vectors
    pair = torch.einsum("asm,mtc->astc", A1, A2)
    want = torch.einsum("ba,bs,bt,astc->bc", E, phi[0], phi[1], pair)
    assert float((out - want).abs().max()) < 1e-12 * float(want.abs().max())


@pytest.mark.parametrize("left_dir", [0, 1])
def test_svd_split_is_the_best_rank_m_approximation(L, left_dir):
    B = rnd(D, 2, NL, 2, D, seed=9)
    m = D
    site_p = torch.empty(D * 2 * m * (NL if left_dir else 1), dtype=torch.float64, device="cuda")
    site_q = torch.empty(m * 2 * D * (1 if left_dir else NL), dtype=torch.float64, device="cuda")
    sv = torch.zeros(2 * D + 2, dtype=torch.float64, device="cuda")
    ws = ws_for(L, "tnml_svd_split_workspace_bytes", D, D, NL, left_dir)
    L.call("tnml_svd_split", B.data_ptr(), site_p.data_ptr(), site_q.data_ptr(), sv.data_ptr(), ws.data_ptr(), D, D, NL, m,
           left_dir, 1, L.F64, st())
    S = sv[:2 * D]
    assert bool((S[:-1] >= S[1:]).all()) and float(S[-1]) > 0
    if not left_dir:
        prod = torch.einsum("asm,mtlc->asltc", site_p.view(D, 2, m), site_q.view(m, 2, NL, D))
        G = torch.einsum("asm,asn->mn", site_p.view(D, 2, m), site_p.view(D, 2, m))
    else:
        prod = torch.einsum("alsm,mtc->asltc", site_p.view(D, NL, 2, m), site_q.view(m, 2, D))
        G = torch.einsum("mtc,ntc->mn", site_q.view(m, 2, D), site_q.view(m, 2, D))
    err2 = float(((B - prod) ** 2).sum())
    assert abs(err2 - float((S[m:] ** 2).sum())) < 1e-10 * float((S ** 2).sum())      # Eckart-Young
    assert abs(float((B ** 2).sum()) - float((S ** 2).sum())) < 1e-11 * float((S ** 2).sum())
    assert float((G - torch.diag(S[:m])).abs().max()) < 1e-10 * float(S[0])           # sqrt(S) on each factor


def test_full_sweep_config3_is_deterministic_and_consistent():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tensornetworkforml_b200 as tn
    import tensornetworkforml_b200.data_generator as gen
    S, Ns = 196, 20000                      # the full chain; a third of the samples keeps the test at a few seconds
    np.random.seed(2)
    data, labels = gen.create_multiclass_dataset(Ns, 14, NL, 0.7)
    X = gen.psi(data.reshape(Ns, -1))
    finals = []
    for run in range(2):
        np.random.seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            net = tn.Network(N=S, M=D, L=NL, normalize=True, calibration_X=X[:1024], act_fn="linear", loss_fn="MSE",
                             truncation="fixed", max_bond=D)
        f = net.forward(X)
        vh = [[], []]
        f = net.sweep(X, labels, f, 1e-4, 1e-3, left_dir=False, var_hist=vh)
        f = net.forward(X)
        f = net.sweep(X, labels, f, 1e-4, 1e-3, left_dir=True, var_hist=vh)
        finals.append(f.elem.copy())
        assert np.isfinite(f.elem).all() and len(vh[0]) == 2 * (S - 1) and net.l_pos == 0
        bonds = net._eng.bond_dims()
        # after a right and a left sweep: the label end keeps min(2 L, ...) = 20, 40, then the cap; the far end 2^k
        assert bonds[:3] == [20, 40, 64] and max(bonds) == 64 and bonds[-5:] == [32, 16, 8, 4, 2]
        fresh = net.forward(X)
        assert np.abs(fresh.elem - f.elem).max() < 1e-9 * np.abs(f.elem).max()
    assert np.array_equal(finals[0], finals[1])
