"""CPU, build container only: run the reference LIVE (imported from /root/reference) next to the oracle.
Skipped where /root/reference does not exist (the GPU box) -- the committed fixtures cover that case."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from oracle import mps_oracle as O

REF = "/root/reference/TensorNetwork"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # the product package aliases these module names for pickle compatibility; make sure we get the reference
    for m in ("Network_class", "Tensor_class", "custom_linalg_tools"):
        mod = sys.modules.get(m)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REF):
            del sys.modules[m]
    import Network_class
    return Network_class


@pytest.mark.parametrize("act,loss,L2,wd", [("softmax", "full_cross_ent", True, 1.0),
                                           ("linear", "MSE", False, 0.01),
                                           ("sigmoid", "cross_entropy", True, 1e-3),
                                           ("linear", "cross_entropy", True, 0.1),
                                           ("softmax", "cross_entropy", False, 0.0)])
def test_live_reference_three_sweeps(act, loss, L2, wd):
    NC = _ref()
    S, M, L, Ns, lr = 9, 5, 2, 56, 0.02
    np.random.seed(11)
    X = O.feature_map(np.random.random((Ns, S)))
    y = np.random.randint(0, L, Ns)
    st = np.random.get_state()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = NC.Network(N=S, M=M, L=L, normalize=True, calibration_X=X, act_fn=act, loss_fn=loss)
    np.random.set_state(st)
    orc = O.OracleMPS.from_seed(S, M, L, calibration_X=X, normalize=True, act_fn=act, loss_fn=loss)
    for sw in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            fr = ref.forward(X)
        fo = orc.forward(X)
        assert np.abs(fr.elem.T - fo).max() <= 1e-12 * np.abs(fo).max()
        left = ref.l_pos == S - 1
        vh = [[], []]
        with contextlib.redirect_stdout(io.StringIO()):
            fr = ref.sweep(X, y, fr, lr, wd, L2_flag=L2, left_dir=left, var_hist=vh)
        n0 = len(orc.hist)
        fo = orc.sweep(y, fo, lr, wd, L2_flag=L2, left_dir=left)
        assert np.abs(fr.elem.T - fo).max() <= 1e-11 * np.abs(fo).max()
        assert np.allclose([h["acc"] for h in orc.hist[n0:]], vh[0], atol=1e-12)
        assert np.allclose([h["mae"] for h in orc.hist[n0:]], vh[1], atol=1e-12)
        assert ref.l_pos == orc.l_pos


def test_live_reference_debug_history_l2_term():
    """debug var_hist rows (NC:741-747): |B|, |dB|, acc, |f|, MAE, L2 loss term, |L2 gradient|."""
    NC = _ref()
    S, M, L, Ns = 7, 4, 2, 32
    np.random.seed(5)
    X = O.feature_map(np.random.random((Ns, S)))
    y = np.random.randint(0, L, Ns)
    st = np.random.get_state()
    with contextlib.redirect_stdout(io.StringIO()):
        ref = NC.Network(N=S, M=M, L=L, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE")
        fr = ref.forward(X)
        vh = [[] for _ in range(7)]
        ref.sweep(X, y, fr, 0.01, 0.5, L2_flag=True, left_dir=False, var_hist=vh, debug=True)
    np.random.set_state(st)
    orc = O.OracleMPS.from_seed(S, M, L, calibration_X=X, normalize=True, act_fn="linear", loss_fn="MSE")
    orc.sweep(y, orc.forward(X), 0.01, 0.5, L2_flag=True, left_dir=False)
    for row, key in ((0, "absB"), (1, "absdB"), (3, "absf"), (4, "mae"), (5, "l2_loss"), (6, "absreg")):
        a = np.array([h[key] for h in orc.hist]); b = np.array(vh[row], dtype=np.float64).reshape(-1)
        assert np.abs(a - b).max() <= 1e-11 * np.abs(b).max(), key
