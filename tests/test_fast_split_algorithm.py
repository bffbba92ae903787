"""CPU: the warm-started deflation split (csrc/svd_fast.cuh) restated in NumPy -- one subspace step from the previous
visit's basis, CholeskyQR through the LDL^T elimination, Rayleigh-Ritz, Jacobi with SINGLE-PRECISION rotation parameters
renormalised in double, and the a-posteriori gates -- against np.linalg.svd (the replacement of NC:887-925 for a bond
that is split again and again).  The device kernels are tested against the same answers in tests/test_gpu_fast_split.py;
this file pins the ALGORITHM (what the gates accept and refuse, what accuracy an accepted split has) without a GPU."""
import numpy as np
import pytest


def bond_like(rng, R, C, spread=0.5, tail=1e-6):
    n = min(R, C)
    Q1, _ = np.linalg.qr(rng.standard_normal((R, n)))
    Q2, _ = np.linalg.qr(rng.standard_normal((C, n)))
    S = np.concatenate([np.linspace(1.0, spread, n // 2), tail * np.logspace(0, -2, n - n // 2)])
    return (Q1 * S) @ Q2.T


def ldl_orthonormalize(Y):
    """Out = D^-1/2 L^-1 Y with Y Y^T = L D L^T: the elimination applied to S and to the rows of Y at once."""
    Y = Y.copy()
    S = Y @ Y.T
    m = len(S)
    d = np.zeros(m)
    floor_ = np.diag(S).max() * 1e-10
    for k in range(m):
        d[k] = S[k, k]
        if not d[k] > floor_:
            return None
        l = S[k + 1:, k] / d[k]
        Y[k + 1:] -= l[:, None] * Y[k][None, :]
        S[k + 1:, k + 1:] -= np.outer(l, S[k + 1:, k])
    return Y / np.sqrt(d)[:, None]


def jacobi_rows(T, tol=1e-8, max_sweeps=30):
    """One-sided Jacobi on the rows of the symmetric T; rotation angle in float32, cos/sin renormalised in double."""
    X = T.copy()
    n = len(X)
    for sweep in range(max_sweeps):
        big = False
        for r in range(n - 1):
            P = [n - 1] + [(r + k) % (n - 1) for k in range(1, n // 2)]
            Q = [r] + [(r - k) % (n - 1) for k in range(1, n // 2)]
            x, y = X[P], X[Q]
            al, be, ga = (x * x).sum(1), (y * y).sum(1), (x * y).sum(1)
            rot = ga * ga > (n * 4.9e-32) * al * be
            big |= bool((rot & (ga * ga > 1e-16 * al * be)).any())
            sc = 1.0 / (al + be)
            df, tf = ((be - al) * sc).astype(np.float32), ((ga + ga) * sc).astype(np.float32)
            h = np.sqrt(df * df + tf * tf)
            den = df + np.copysign(h, df)
            t = np.where(den != 0, tf / np.where(den != 0, den, 1), np.float32(0))
            cf = (1 / np.sqrt(1 + t * t)).astype(np.float32)
            c, s = cf.astype(np.float64), (cf * t).astype(np.float64)
            e = c * c + s * s - 1
            nu = 1 - e / 2 + 0.375 * e * e
            c, s = np.where(rot, c * nu, 1.0), np.where(rot, s * nu, 0.0)
            X[P], X[Q] = c[:, None] * x - s[:, None] * y, s[:, None] * x + c[:, None] * y
        if not big:
            return X, sweep + 1
    return X, max_sweeps


def fast_split(G, V0, m):
    """-> (U rows, lam, info) or (None, None, reason)"""
    V = V0
    prev = np.inf
    ok = False
    for it in range(4):
        Q = ldl_orthonormalize(V @ G)
        if Q is None:
            return None, None, "cholqr"
        Z = Q @ G
        T = Q @ Z.T
        T = 0.5 * (T + T.T)
        resid2 = float(((Z - T @ Q) ** 2).sum())
        mind = np.diag(T).min()
        ok = mind > 0 and resid2 <= 1e-24 * mind * mind
        if ok:
            break
        if it > 0 and not resid2 < 1e-3 * prev:
            break
        prev, V = resid2, Q
    if not ok:
        return None, None, "residual"
    X, sweeps = jacobi_rows(T)
    lam = np.sqrt((X * X).sum(1))
    order = np.argsort(-lam, kind="stable")
    lam, W = lam[order], (X / np.sqrt((X * X).sum(1))[:, None])[order]
    tau = np.trace(G) - np.trace(T)
    if not (resid2 <= 1e-24 * lam[-1] ** 2 and tau <= 0.25 * lam[-1] and lam[-1] >= 1e-6 * lam[0] and sweeps < 30):
        return None, None, "gates"
    return W @ Q, lam, dict(sweeps=sweeps, steps=it + 1)


@pytest.mark.parametrize("spread", [0.5, 0.997])
def test_accepted_split_matches_svd(spread):
    rng = np.random.default_rng(0)
    A = bond_like(rng, 128, 1280, spread)
    m = 64
    U0 = np.linalg.svd(A, full_matrices=False)[0][:, :m].T
    for visit in range(3):
        A = A + 1e-3 * bond_like(rng, 128, 1280, spread) + 1e-6 * np.abs(A).max() * rng.standard_normal(A.shape)
        G = A @ A.T
        U, lam, info = fast_split(G, U0, m)
        assert U is not None, info
        Ue, Se, Vhe = np.linalg.svd(A, full_matrices=False)
        sv = np.sqrt(lam)
        assert np.abs(sv - Se[:m]).max() / Se[0] < 2e-13
        prod = (U.T * np.sqrt(sv)) @ ((U @ A) / np.sqrt(sv)[:, None])
        want = (Ue[:, :m] * Se[:m]) @ Vhe[:m]
        assert np.abs(prod - want).max() / np.abs(want).max() < 1e-11
        assert np.abs(U @ U.T - np.eye(m)).max() < 1e-13
        U0 = U


def test_gates_refuse_what_the_path_cannot_deliver():
    rng = np.random.default_rng(1)
    m = 64
    A = bond_like(rng, 128, 1280)
    U0 = np.linalg.svd(A, full_matrices=False)[0][:, :m].T
    Ag = rng.standard_normal((128, 1280))                                     # no gap at m
    assert fast_split(Ag @ Ag.T, U0, m)[0] is None
    As = bond_like(rng, 128, 1280, spread=1e-4)                               # kept values down to 1e-4 sigma_max
    Us = np.linalg.svd(As, full_matrices=False)[0][:, :m].T
    assert fast_split(As @ As.T, Us, m)[0] is None
    Ar = A[:, :40] @ rng.standard_normal((40, 1280))                          # rank 40 < m: CholeskyQR must break down
    assert fast_split(Ar @ Ar.T, U0, m)[0] is None


def test_dominance_gate_refuses_an_invariant_subspace_that_is_not_the_dominant_one():
    """What tr G - tr T <= 0.25 lambda_m is there for: a start basis that is EXACTLY invariant but misses one of the m
    dominant directions (here eigenvector 64 replaced by eigenvector 65) stays invariant under the subspace steps, so the
    residual gate passes, and its smallest Ritz value (0.1 lambda_1) passes the range gate -- only the eigenvalue mass
    left outside the subspace (lambda_64 = 0.5) gives it away.  (The device kernel read tr G from the wrong shared-memory
    slot until the end of round 2, which made this gate inert for matrices of order-one scale)"""
    rng = np.random.default_rng(3)
    n, m = 128, 64
    W, _ = np.linalg.qr(rng.standard_normal((n, n)))
    lam = np.concatenate([np.linspace(1.0, 0.5, m), [0.1], 1e-12 * np.ones(n - m - 1)])
    G = (W * lam) @ W.T
    G = 0.5 * (G + G.T)
    dominant = W[:, :m].T
    U, lam_out, info = fast_split(G, dominant, m)
    assert U is not None and np.abs(np.sort(lam_out)[::-1] - lam[:m]).max() < 1e-13, info
    swapped = np.vstack([W[:, :m - 1].T, W[:, m:m + 1].T])
    Q = ldl_orthonormalize(swapped @ G)                      # one subspace step: still invariant, residual at rounding
    Z = Q @ G
    T = Q @ Z.T
    assert np.sqrt(((Z - T @ Q) ** 2).sum()) < 1e-12 * np.linalg.eigvalsh(0.5 * (T + T.T)).min()
    assert np.trace(G) - np.trace(T) > 0.25 * 0.1
    assert fast_split(G, swapped, m) == (None, None, "gates")

