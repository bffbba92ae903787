"""Offline study (CPU, NumPy): two-group Jacobi schedule on the dumped bond tensors (tools/dump_bond.py).
Rows of the pivoted Cholesky factor in pivot order; group B = the k rows above the gap, group T = the rest.
Phase A: cyclic sweeps inside each group (the groups run concurrently on different CTAs); phase X: all cross pairs.
Counts sequential rotation SETS against the plain cyclic schedule (n - 1 sets per sweep)."""
import sys
import numpy as np
sys.path.insert(0, "tools")
from jacobi_study import pivoted_cholesky, EPS


def rotate(W, P, Q, tol):
    x, y = W[P], W[Q]
    al = np.einsum("ij,ij->i", x, x); be = np.einsum("ij,ij->i", y, y); ga = np.einsum("ij,ij->i", x, y)
    ab = al * be
    rel = np.where(ab > 0, ga * ga / np.where(ab > 0, ab, 1), 0.0)
    rot = rel > tol * tol
    de = be - al
    h = np.sqrt(de * de + 4 * ga * ga)
    den = de + np.copysign(h, de)
    t = np.where(rot & (den != 0), 2 * ga / np.where(den != 0, den, 1), 0.0)
    c = 1 / np.sqrt(1 + t * t); s = c * t
    W[P] = c[:, None] * x - s[:, None] * y
    W[Q] = s[:, None] * x + c[:, None] * y
    return float(np.sqrt(rel.max())) if len(rel) else 0.0


def round_robin(idx):
    """Sets of disjoint pairs covering all pairs of idx (circle method)."""
    idx = list(idx)
    if len(idx) % 2:
        idx.append(None)
    n = len(idx)
    pos = idx[:]
    for _ in range(n - 1):
        P, Q = [], []
        for i in range(n // 2):
            a, b = pos[i], pos[n - 1 - i]
            if a is not None and b is not None:
                P.append(a); Q.append(b)
        yield np.array(P, int), np.array(Q, int)
        pos = [pos[0]] + [pos[-1]] + pos[1:-1]


def cross_sets(A, B):
    A, B = list(A), list(B)
    if len(A) < len(B):
        A, B = B, A
    for s in range(len(A)):
        P = [A[(i + s) % len(A)] for i in range(len(B))]
        yield np.array(P, int), np.array(B, int)


def plain(W, tol):
    W = W.copy(); n = len(W); sets = 0
    for sweep in range(40):
        mx = 0.0
        for P, Q in round_robin(range(n)):
            mx = max(mx, rotate(W, P, Q, tol)); sets += 1
        if mx < 1e-8:
            break
    return sets, sweep + 1, W


def split(W, k, tol):
    W = W.copy(); n = len(W); sets = 0; log = []
    gB, gT = range(k), range(k, n)
    for cycle in range(20):
        # phase A: both groups sweep concurrently until each has converged
        sa = [0, 0]; mxA = 0.0
        for gi, g in enumerate((gB, gT)):
            for sweep in range(40):
                mx = 0.0
                for P, Q in round_robin(g):
                    mx = max(mx, rotate(W, P, Q, tol)); sa[gi] += 1
                if sweep == 0:
                    mxA = max(mxA, mx)
                if mx < 1e-8:
                    break
        sets += max(sa)
        # phase X: cross pairs
        mxX = 0.0; sx = 0
        for P, Q in cross_sets(gB, gT):
            mxX = max(mxX, rotate(W, P, Q, tol)); sx += 1
        sets += sx
        log.append((max(sa), "%.0e" % mxA, sx, "%.0e" % mxX))
        if mxX < 1e-8 and (cycle > 0 and mxA < 1e-8):
            break
    return sets, log, W


if __name__ == "__main__":
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bond_dumps.npz")
    for name in sorted(d.files):
        M = d[name]
        A = M if M.shape[0] <= M.shape[1] else M.T
        n = A.shape[0]
        G = A @ A.T
        R, piv = pivoted_cholesky(G)
        W = R[:len(piv)]
        if len(piv) < n:
            W = np.vstack([W, np.eye(n)[len(piv):] * np.sqrt(G.diagonal().max() * n * EPS)])
        dg = np.einsum("ij,ij->i", W, W)
        ratio = dg[1:] / dg[:-1]
        k = int(np.argmin(ratio)) + 1
        gap = ratio[k - 1]
        tol = np.sqrt(n) * EPS
        s_plain, sw, Wp = plain(W, tol)
        sv_ref = np.linalg.svd(A, compute_uv=False)
        s_split, log, Ws = split(W, k, tol)
        svp = np.sort(np.sqrt(np.einsum("ij,ij->i", Wp, Wp)))[::-1]
        svs = np.sort(np.sqrt(np.einsum("ij,ij->i", Ws, Ws)))[::-1]
        print("%-12s n %d k %d gap %.1e | plain: %d sets (%d sweeps) err %.1e | split: %d sets err %.1e  %s" % (
            name, n, k, gap, s_plain, sw, np.abs(svp - sv_ref).max() / sv_ref[0], s_split,
            np.abs(svs - sv_ref).max() / sv_ref[0], log))
