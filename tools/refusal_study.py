"""Offline study (CPU, NumPy): which warm-started splits would the device-side gates refuse late in a training run,
and why?

Runs the oracle on the bench workload (config 3; NS / NSWEEP / D from the environment) and, at every split with a short
side of 2D whose bond was split before in the same direction, runs the NumPy restatement of csrc/svd_fast.cuh
(tests/test_fast_split_algorithm.py) from the previous visit's dominant basis -- at EVERY revisit, whatever a backoff
policy would have done (the warm buffer holds the previous visit's basis after either pipeline, so the outcome of an
attempt does not depend on earlier attempts).  One row per (sweep, bond, direction):

    sweep p dir visit code steps tau/lam_m lam_{m+1}/lam_m lam_m/lam_1 sin(theta) resid/lam_m jacobi_sweeps

(jacobi_sweeps: sweeps of the restated one-sided Jacobi on T with single-precision rotation parameters, JACOBI=1 only.)

code: 0 accepted | 2 CholeskyQR breakdown | 3 residual after the subspace steps | 4.1 residual against the true lam_m |
4.2 tau > 0.25 lam_m (no gap: the subspace is not the dominant one) | 4.3 lam_m < 1e-6 lam_1.
The table goes to OUT (.npy); tools/refusal_policy_eval.py replays backoff policies on it.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mps_oracle as O                                   # noqa: E402
import tensornetworkforml_b200.data_generator as gen                 # noqa: E402
from tests.test_fast_split_algorithm import jacobi_rows, ldl_orthonormalize       # noqa: E402

S, L, D = int(os.environ.get("S", 196)), 10, int(os.environ.get("D", 64))
Ns = int(os.environ.get("NS", 60000))
NSWEEP = int(os.environ.get("NSWEEP", 25))
OUT = os.environ.get("OUT", "/tmp/refusal_table.npy")
JACOBI = os.environ.get("JACOBI", "0") != "0"        # also run the restated single-precision-parameter Jacobi on T
lr, wd = 1e-4, 1e-3
LAST = dict(sweeps=np.nan)


def attempt(G, V0, m):
    """The gates of k_fast_split with the reason of a refusal; eigenvalues of T by LAPACK (the Jacobi sweep count gate,
    sweeps < 30, never binds on these matrices)."""
    V, prev, ok, it = V0, np.inf, False, 0
    resid2 = mind = 0.0
    for it in range(4):
        Q = ldl_orthonormalize(V @ G)
        if Q is None:
            return 2.0, it + 1, np.nan
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
        return 3.0, it + 1, (np.sqrt(resid2) / mind if mind > 0 else np.inf)
    lam = np.sort(np.linalg.eigvalsh(T))[::-1]
    if JACOBI:
        LAST["sweeps"] = jacobi_rows(T, max_sweeps=60)[1]
    tau = np.trace(G) - np.trace(T)
    rr = np.sqrt(resid2) / lam[-1]
    if not resid2 <= 1e-24 * lam[-1] ** 2:
        return 4.1, it + 1, rr
    if not tau <= 0.25 * lam[-1]:
        return 4.2, it + 1, rr
    if not lam[-1] >= 1e-6 * lam[0]:
        return 4.3, it + 1, rr
    return 0.0, it + 1, rr


def main():
    np.random.seed(2)
    side = int(round(S ** 0.5))
    data, labels = gen.create_multiclass_dataset(Ns, side, L, 0.7)
    X = gen.psi(data.reshape(Ns, -1))
    np.random.seed(2)
    net = O.OracleMPS.from_seed(S, D, L, calibration_X=X[:2048], normalize=True, act_fn="linear", loss_fn="MSE",
                                rule="fixed", max_bond=D)
    basis, visits, rows = {}, {}, []
    cur = dict(sweep=0)
    orig = O.svd_split

    def spy(B, left_dir, m):
        a, _, L_, _, c = B.shape
        p = net.l_pos - 1 if left_dir else net.l_pos
        Mx = B.reshape(a * 2, L_ * 2 * c) if not left_dir else B.reshape(a * 2 * L_, 2 * c)
        n = min(Mx.shape)
        if n == 2 * D and 2 * m == n:
            G = Mx @ Mx.T if not left_dir else Mx.T @ Mx
            w, Vec = np.linalg.eigh(G)
            w, Vec = w[::-1], Vec[:, ::-1]
            key = (p, bool(left_dir))
            v = visits[key] = visits.get(key, 0) + 1
            if key in basis:
                V0 = basis[key]
                LAST["sweeps"] = np.nan
                code, steps, rr = attempt(G, V0, m)
                # sine of the largest principal angle between the old basis and the new dominant subspace
                sin_t = np.linalg.norm(Vec[:, m:].T @ V0.T, 2)
                rows.append((cur["sweep"], p, int(left_dir), v, code, steps, w[m:].sum() / w[m - 1], w[m] / w[m - 1],
                             w[m - 1] / w[0], sin_t, rr, LAST["sweeps"]))
            basis[key] = np.ascontiguousarray(Vec[:, :m].T)
        return orig(B, left_dir, m)

    O.svd_split = spy
    t0 = time.time()
    for sw in range(NSWEEP):
        cur["sweep"] = sw + 1
        f = net.forward(X)
        left = net.l_pos == S - 1
        net.sweep(labels, f, lr, wd, True, left)
        tab = np.array(rows)
        np.save(OUT, tab)
        mine = tab[tab[:, 0] == sw + 1] if len(tab) else tab
        codes = {c: int((mine[:, 4] == c).sum()) for c in np.unique(mine[:, 4])} if len(mine) else {}
        print("sweep %d %s %.0f s  mae %.5f  attempts by outcome %s" % (sw + 1, "left" if left else "right", time.time() - t0,
                                                                      net.hist[-1]["mae"], codes), flush=True)


if __name__ == "__main__":
    main()
