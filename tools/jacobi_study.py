"""Offline study (CPU, NumPy) of the one-sided Jacobi sweep count on bond tensors dumped by tools/dump_bond.py.
Emulates the block ordering of k_jacobi_cluster_w (blocks of 4 rows, circle method over block pairs) and counts the
sweeps to convergence for different preconditioners / row orders."""
import sys
import numpy as np

EPS = 2.220446049250313e-16


def pivoted_cholesky(G):
    """Returns R (rows in pivot order, upper-trapezoidal in the permuted basis) and the pivot list: G[piv][:,piv] = R'^T R'."""
    n = G.shape[0]
    S = G.copy()
    R = np.zeros((n, n))
    active = np.ones(n, bool)
    piv = []
    floor = None
    for k in range(n):
        d = np.where(active, np.diag(S), -1.0)
        c = int(np.argmax(d))
        if floor is None:
            floor = d[c] * n * EPS
        if d[c] <= floor:
            break
        r = S[c] / np.sqrt(d[c])
        r[~active] = 0.0
        R[k] = r
        S -= np.outer(r, r)
        active[c] = False
        piv.append(c)
    return R, piv


def sweep_sets(n):
    """Yield index arrays (p, q) of the disjoint rotations of each set, in the order of the cluster kernel."""
    K = 4
    NB = n // K
    pos = list(range(NB))  # circle method: position 0 fixed
    for rnd in range(NB - 1):
        pairs = []
        ring = pos
        for w in range(NB // 2):
            bi = ring[0] if w == 0 else ring[w]
            bj = ring[NB - 1 - w]
            pairs.append((bi, bj))
        if rnd == 0:
            for s in range(3):
                P, Q = [], []
                for bi, bj in pairs:
                    for base in (bi * K, bj * K):
                        for j in range(2):
                            if s == 0:
                                p, q = 2 * j, 2 * j + 1
                            elif s == 1:
                                p, q = j, j + 2
                            else:
                                p, q = j, 3 - j
                            P.append(base + p); Q.append(base + q)
                yield np.array(P), np.array(Q)
        for s in range(4):
            P, Q = [], []
            for bi, bj in pairs:
                for w in range(4):
                    P.append(bi * K + w); Q.append(bj * K + (w + s) % 4)
            yield np.array(P), np.array(Q)
        pos = [pos[0]] + [pos[-1]] + pos[1:-1]


def jacobi(W, max_sweeps=40, tol=None, sort_each_sweep=False):
    W = W.copy()
    n = W.shape[0]
    tol = tol or np.sqrt(n) * EPS
    sets = list(sweep_sets(n))
    hist = []
    for sweep in range(max_sweeps):
        if sort_each_sweep:
            order = np.argsort(-np.einsum("ij,ij->i", W, W), kind="stable")
            W = W[order]
        big = False
        nrot = 0
        mx = 0.0
        for P, Q in sets:
            x, y = W[P], W[Q]
            al = np.einsum("ij,ij->i", x, x); be = np.einsum("ij,ij->i", y, y); ga = np.einsum("ij,ij->i", x, y)
            ab = al * be
            rel = np.where(ab > 0, ga * ga / np.where(ab > 0, ab, 1), 0.0)
            rot = rel > tol * tol
            mx = max(mx, float(np.sqrt(rel.max())))
            big |= bool((rel > 1e-16).any())
            nrot += int(rot.sum())
            de = be - al
            h = np.sqrt(de * de + 4 * ga * ga)
            den = de + np.copysign(h, de)
            t = np.where(rot & (den != 0), 2 * ga / np.where(den != 0, den, 1), 0.0)
            c = 1 / np.sqrt(1 + t * t); s = c * t
            W[P] = c[:, None] * x - s[:, None] * y
            W[Q] = s[:, None] * x + c[:, None] * y
        hist.append((nrot, mx))
        if not big:
            break
    return W, hist


def study(name, M):
    R_, C_ = M.shape
    A = M if R_ <= C_ else M.T
    n = A.shape[0]
    G = A @ A.T
    sv = np.linalg.svd(A, compute_uv=False)
    R, piv = pivoted_cholesky(G)
    k = len(piv)
    rest = [i for i in range(n) if i not in piv]
    perm = piv + rest
    # current kernel: rows scattered by pivot (physical row c holds the k-th factor row), columns in original order
    tiny = np.sqrt(G.diagonal().max() * n * EPS)
    Wscat = np.zeros((n, n))
    for kk, c in enumerate(piv):
        Wscat[c] = R[kk]
    for c in rest:
        Wscat[c, c] = tiny
    Wsort = Wscat[perm]
    res = {}
    for label, W in (("scattered (current)", Wscat), ("pivot order", Wsort)):
        _, h = jacobi(W)
        res[label] = h
    # second factorisation: R2 upper with R2^T R2 = Wsort Wsort^T; Jacobi on the rows of R2
    Gs = Wsort @ Wsort.T
    try:
        R2 = np.linalg.cholesky(Gs).T
        _, h = jacobi(R2)
        res["double (chol of R R^T), rows of R2"] = h
        _, h = jacobi(R2.T.copy())
        res["double, rows of R2^T (= L)"] = h
    except np.linalg.LinAlgError:
        res["double"] = "cholesky failed"
    Q2, R2q = np.linalg.qr(Wsort.T)
    _, h = jacobi(R2q)
    res["double (QR of R^T), rows of R2"] = h
    _, h = jacobi(R2q.T.copy())
    res["double (QR), rows of L=R2^T"] = h
    _, h = jacobi(Wsort, sort_each_sweep=True)
    res["pivot order + resort each sweep"] = h
    print("==", name, M.shape, "n", n, "rank(chol)", k, "sigma max %.3e s64 %.3e s65 %.3e min %.3e" % (
        sv[0], sv[min(63, n - 1)], sv[min(64, n - 1)], sv[-1]))
    for label, h in res.items():
        if isinstance(h, str):
            print("   %-40s %s" % (label, h)); continue
        print("   %-40s sweeps %2d  rotations/sweep %s  max rel %s" % (
            label, len(h), [x[0] for x in h], ["%.0e" % x[1] for x in h]))


if __name__ == "__main__":
    d = np.load(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/bond_dumps.npz")
    for k in sorted(d.files):
        study(k, d[k])
