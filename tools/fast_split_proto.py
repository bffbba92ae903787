"""NumPy prototype of the warm-started deflation split (what k_fast_split does on the device), run on the matrices
dumped by tools/warm_split_study.py.  Checks: subspace residual gate, two-sided Jacobi with single-precision rotation
parameters + double renormalisation (sweeps to convergence), final factors against np.linalg.svd."""
import sys, os
import numpy as np

d = np.load(os.environ.get("IN", "/tmp/warm_dumps.npz"))
keys = sorted(d.files, key=lambda k: (int(k.split("_")[0][1:]), int(k.split("_")[1]), int(k.split("s")[-1])))


def round_robin_sets(n):
    """circle method: n-1 sets of n/2 disjoint pairs"""
    sets = []
    for r in range(n - 1):
        P, Q = [n - 1], [r]
        for k in range(1, n // 2):
            P.append((r + k) % (n - 1)); Q.append((r - k) % (n - 1))
        sets.append((np.array(P), np.array(Q)))
    return sets


def jacobi_two_sided(T, tol=1e-8, f32=True, maxs=30):
    n = len(T)
    T = T.copy(); Wt = np.eye(n)
    sets = round_robin_sets(n)
    for sweep in range(maxs):
        mx = 0.0
        for P, Q in sets:
            app, aqq, apq = T[P, P], T[Q, Q], T[P, Q]
            rel = np.abs(apq) / np.sqrt(np.abs(app * aqq))
            mx = max(mx, rel.max())
            rot = rel > 1e-17
            # t = 2 apq / (de + sign(de) sqrt(de^2 + 4 apq^2)), de = aqq - app  (|t| <= 1)
            sc = 1.0 / (np.abs(app) + np.abs(aqq))
            de, tw = (aqq - app) * sc, 2 * apq * sc
            if f32:
                de32, tw32 = de.astype(np.float32), tw.astype(np.float32)
                h = np.sqrt(de32 * de32 + tw32 * tw32)
                den = de32 + np.copysign(h, de32)
                t = np.where(den != 0, tw32 / np.where(den != 0, den, 1), np.float32(0))
                c32 = (1 / np.sqrt(1 + t * t)).astype(np.float32)
                c = c32.astype(np.float64); s = (c32 * t).astype(np.float64)
                e = c * c + s * s - 1
                nu = 1 - e / 2 + 3 * e * e / 8
                c, s = c * nu, s * nu
            else:
                h = np.sqrt(de * de + tw * tw)
                den = de + np.copysign(h, de)
                t = np.where(den != 0, tw / np.where(den != 0, den, 1), 0.0)
                c = 1 / np.sqrt(1 + t * t); s = c * t
            c = np.where(rot, c, 1.0); s = np.where(rot, s, 0.0)
            # rows: new_p = c row_p - s row_q ; new_q = s row_p + c row_q ; then the same on columns
            rp, rq = T[P, :].copy(), T[Q, :].copy()
            T[P, :] = c[:, None] * rp - s[:, None] * rq
            T[Q, :] = s[:, None] * rp + c[:, None] * rq
            cp, cq = T[:, P].copy(), T[:, Q].copy()
            T[:, P] = cp * c[None, :] - cq * s[None, :]
            T[:, Q] = cp * s[None, :] + cq * c[None, :]
            wp, wq = Wt[P, :].copy(), Wt[Q, :].copy()
            Wt[P, :] = c[:, None] * wp - s[:, None] * wq
            Wt[Q, :] = s[:, None] * wp + c[:, None] * wq
        if mx < tol:
            return np.diag(T).copy(), Wt, sweep + 1
    return np.diag(T).copy(), Wt, maxs


prev = {}
worst = dict(sv=0.0, prod=0.0, orth=0.0)
for k in keys:
    bond = k.rsplit("_s", 1)[0]
    Mx = d[k]
    short_rows = Mx.shape[0] <= Mx.shape[1]
    A = Mx if short_rows else Mx.T          # n x Nl, short side first
    n = A.shape[0]; m = n // 2
    G = A @ A.T
    Ue, Se, Vhe = np.linalg.svd(A, full_matrices=False)
    if bond in prev and prev[bond].shape == (m, n):
        V0 = prev[bond]
        Yv = V0 @ G
        S = Yv @ Yv.T
        Lc = np.linalg.cholesky(S)
        Qv = np.linalg.solve(Lc, Yv)
        Zv = Qv @ G
        T = Qv @ Zv.T
        T = 0.5 * (T + T.T)
        Rres = Zv - T @ Qv
        tau = np.trace(G) - np.trace(T)
        lam, Wt, nsw = jacobi_two_sided(T)
        order = np.argsort(-lam)
        lam, Wt = lam[order], Wt[order]
        Uv = Wt @ Qv                         # rows = left singular vectors
        ok = np.linalg.norm(Rres) <= 1e-12 * lam[-1] and tau < 1e-2 * lam[-1]
        sv = np.sqrt(lam)
        e_sv = np.abs(sv - Se[:m]).max() / Se[0]
        short = Uv.T * np.sqrt(sv)[None, :]
        long_ = (Uv @ A) / np.sqrt(sv)[:, None]
        ref = (Ue[:, :m] * Se[:m]) @ Vhe[:m]
        e_prod = np.abs(short @ long_ - ref).max() / np.abs(ref).max()
        e_orth = np.abs(Uv @ Uv.T - np.eye(m)).max()
        worst["sv"] = max(worst["sv"], e_sv); worst["prod"] = max(worst["prod"], e_prod); worst["orth"] = max(worst["orth"], e_orth)
        print("%-12s ok=%d  |R|/lam_m=%.1e tau/lam_m=%.1e  sweeps=%d  err sv=%.1e prod=%.1e orth=%.1e" % (
            k, ok, np.linalg.norm(Rres) / lam[-1], tau / lam[-1], nsw, e_sv, e_prod, e_orth), flush=True)
        prev[bond] = Uv
    else:
        prev[bond] = Ue[:, :m].T.copy()
print("worst", worst)
