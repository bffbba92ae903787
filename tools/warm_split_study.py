"""Offline study (CPU, NumPy): warm-started subspace split.

Runs the oracle on the bench workload (config 3 at reduced Ns), records the SVD matrix of selected bonds in every
sweep, and asks for each (bond, direction) revisit:
  * how far the previous visit's top-m left basis Q0 is from the new top-m invariant subspace of the Gram matrix G,
  * the subspace error after ONE multiplication Y = G Q0 + orthonormalisation,
  * how diagonal the Rayleigh-Ritz matrix T = Q^T G Q is, and how many cyclic Jacobi sweeps T needs (warm vs cold).
"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mps_oracle as O
import tensornetworkforml_b200.data_generator as gen

S, L, D = 196, 10, int(os.environ.get("D", 64))
Ns = int(os.environ.get("NS", 3000))
NSWEEP = int(os.environ.get("NSWEEP", 8))
lr, wd = 1e-4, 1e-3
WATCH = set(int(x) for x in os.environ.get("WATCH", "10,40,70,98,130,160,185").split(","))

np.random.seed(2)
side = 14
data, labels = gen.create_multiclass_dataset(Ns, side, L, 0.7)
X = gen.psi(data.reshape(Ns, -1))
np.random.seed(2)
net = O.OracleMPS.from_seed(S, D, L, calibration_X=X[:2048], normalize=True, act_fn="linear", loss_fn="MSE",
                            rule="fixed", max_bond=D)

dumps = {}          # (p, left_dir) -> list of Mx per visit
orig = O.svd_split
cur = dict(sweep=0)


def spy(B, left_dir, m):
    a, _, L_, _, c = B.shape
    # p = position of the left site of the pair
    p = net.l_pos - 1 if left_dir else net.l_pos
    if p in WATCH:
        Mx = B.reshape(a * 2, L_ * 2 * c) if not left_dir else B.reshape(a * 2 * L_, 2 * c)
        dumps.setdefault((p, left_dir), []).append((cur["sweep"], Mx.copy(), m))
    return orig(B, left_dir, m)


O.svd_split = spy
t0 = time.time()
for sw in range(NSWEEP):
    cur["sweep"] = sw
    f = net.forward(X)
    left = net.l_pos == S - 1
    net.sweep(labels, f, lr, wd, True, left)
    print("sweep", sw, "left" if left else "right", "%.1fs" % (time.time() - t0), "mae", net.hist[-1]["mae"], flush=True)
np.savez_compressed(os.environ.get("OUT", "/tmp/warm_dumps.npz"),
                    **{"p%d_%d_s%d" % (p, int(ld), sw): Mx for (p, ld), v in dumps.items() for sw, Mx, m in v})


def jacobi_sweeps(T, tol=1e-8, maxs=30):
    """two-sided cyclic Jacobi (round robin); returns sweeps until a sweep sees max relative offdiag < tol"""
    T = T.copy(); n = len(T)
    for sweep in range(maxs):
        mx = 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = T[p, q]
                d = np.sqrt(abs(T[p, p] * T[q, q]))
                rel = abs(apq) / d if d > 0 else 0.0
                mx = max(mx, rel)
                if rel < 1e-17:
                    continue
                th = (T[q, q] - T[p, p]) / (2 * apq)
                t = np.sign(th) / (abs(th) + np.sqrt(1 + th * th)) if th != 0 else 1.0
                c = 1 / np.sqrt(1 + t * t); s = c * t
                J = np.array([[c, s], [-s, c]])
                T[[p, q], :] = J.T @ T[[p, q], :]
                T[:, [p, q]] = T[:, [p, q]] @ J
        if mx < tol:
            return sweep + 1, mx
    return maxs, mx


for (p, ld), v in sorted(dumps.items()):
    prevU = None
    for sw, Mx, m in v:
        short_rows = Mx.shape[0] <= Mx.shape[1]
        G = Mx @ Mx.T if short_rows else Mx.T @ Mx
        lam, V = np.linalg.eigh(G)
        lam, V = lam[::-1], V[:, ::-1]
        n = len(lam)
        U1 = V[:, :m]
        line = "p=%3d %s sweep %d n=%d m=%d  s_m/s_1=%.2e s_m+1/s_m=%.2e mingap(rel)=%.1e" % (
            p, "L" if ld else "R", sw, n, m, np.sqrt(lam[m - 1] / lam[0]),
            np.sqrt(max(lam[m], 0) / lam[m - 1]) if m < n else 0.0,
            np.min(np.abs(np.diff(np.sqrt(lam[:m]))) / np.sqrt(lam[0])))
        if prevU is not None and prevU.shape == U1.shape and m < n:
            Q0 = prevU
            sv = np.linalg.svd(U1.T @ Q0, compute_uv=False)
            th0 = np.sqrt(max(0.0, 1 - sv.min() ** 2))
            Y = G @ Q0
            Q, _ = np.linalg.qr(Y)
            sv1 = np.linalg.svd(V[:, m:].T @ Q, compute_uv=False)
            T = Q.T @ G @ Q
            dd = np.sqrt(np.abs(np.outer(np.diag(T), np.diag(T))))
            off = np.abs(T - np.diag(np.diag(T))) / dd
            # CholQR conditioning: column-scaled Gram of Y
            Sg = Y.T @ Y
            ds = np.sqrt(np.diag(Sg))
            condS = np.linalg.cond(Sg / np.outer(ds, ds))
            nsw_warm, _ = jacobi_sweeps(T)
            Tc = U1.T @ G @ U1      # exact subspace, random basis -> cold count
            Rq, _ = np.linalg.qr(np.random.default_rng(0).standard_normal((m, m)))
            nsw_cold, _ = jacobi_sweeps(Rq.T @ np.diag(lam[:m]) @ Rq)
            line += "  sin(th0)=%.1e sin(th1)=%.1e offT=%.1e condS=%.1e jac warm=%d cold=%d" % (
                th0, sv1.max(), off.max(), condS, nsw_warm, nsw_cold)
        print(line, flush=True)
        prevU = U1
