"""Shared helpers: load a golden case written by tests/golden/make_golden.py and replay it on an engine."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SWEEP_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                     if not os.path.basename(p).startswith("diag_model"))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    case = {k: z[k] for k in z.files}
    S = int(case["S_sites"])
    case["sites0"] = [case["site0_%d" % p] for p in range(S)]
    for k in ("S_sites", "M", "L", "nsweeps", "max_bond", "L2"):
        case[k] = int(case[k])
    for k in ("wd", "lr", "T"):
        case[k] = float(case[k])
    for k in ("act", "loss"):
        case[k] = str(case[k])
    case["rule"] = "reference" if case["max_bond"] < 0 else "fixed"
    return case


def rel(a, b):
    """max |a-b| relative to max |b| (the comparison SURVEY.md section 7 prescribes for f and S)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_sweep_against_golden(case, sw, f_fwd, f_swp, acc, mae, svals, bonds, tol, tol_metric=None):
    """Compare one replayed sweep with the reference's record.  svals: list of 1-D arrays (all singular values)."""
    tol_metric = tol if tol_metric is None else tol_metric
    assert rel(f_fwd, case["f_fwd_%d" % sw]) < tol, "forward f"
    assert rel(f_swp, case["f_swp_%d" % sw]) < tol, "post-sweep f"
    assert np.abs(np.asarray(mae) - case["mae_%d" % sw]).max() < tol_metric, "MAE history"
    assert np.abs(np.asarray(acc) - case["acc_%d" % sw]).max() < 1e-12, "accuracy history"
    sv_ref = case["sv_%d" % sw]
    assert len(svals) == sv_ref.shape[0]
    for i, s in enumerate(svals):
        r = sv_ref[i][~np.isnan(sv_ref[i])]
        s = np.asarray(s)[:len(r)]
        assert np.abs(s - r[:len(s)]).max() / r.max() < tol, "singular values at step %d" % i
    assert list(bonds) == list(case["bonds_%d" % sw]), "bond dimensions"
