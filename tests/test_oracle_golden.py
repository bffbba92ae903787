"""CPU: pin oracle/mps_oracle.py against fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import mps_oracle as O
from tests import _golden as G


@pytest.mark.parametrize("name", G.SWEEP_CASES)
def test_oracle_replays_reference_sweeps(name):
    case = G.load(name)
    net = O.OracleMPS(case["sites0"], case["L"], T=case["T"], act_fn=case["act"], loss_fn=case["loss"],
                      rule=case["rule"], max_bond=case["max_bond"])
    X, y = case["X"], case["y"]
    for sw in range(case["nsweeps"]):
        f = net.forward(X)
        left = net.l_pos == net.S - 1
        assert int(left) == int(case["left_%d" % sw])
        n0 = len(net.hist)
        f2 = net.sweep(y, f, case["lr"], case["wd"], L2_flag=bool(case["L2"]), left_dir=left)
        h = net.hist[n0:]
        G.check_sweep_against_golden(case, sw, f, f2, [r["acc"] for r in h], [r["mae"] for r in h],
                                     [r["S"] for r in h], net.bond_dims(), tol=1e-11)
    assert G.rel(net.forward(X), case["f_final"]) < 1e-11


def test_oracle_known_answer_trained_diag_model():
    """forward of the shipped trained_diag_model.dat on np.random.seed(123) input."""
    z = np.load(G.GOLDEN_DIR + "/diag_model_known_answer.npz")
    S = 64
    sites = [z["site_%d" % p] for p in range(S)]
    net = O.OracleMPS(sites, int(z["L"]), T=float(z["T"]), act_fn=str(z["act"]), loss_fn=str(z["loss"]),
                      l_pos=int(z["l_pos"]))
    f = net.forward(z["X"])
    assert G.rel(f, z["f"]) < 1e-12
    expect = np.array([[0.03107, 0.00638037, -0.00083815, -0.00064346],
                       [-0.03112103, -0.00638817, 0.00083617, 0.00064254]]).T
    assert np.abs(f - expect).max() < 1e-7


def test_feature_map_is_sin_first():
    x = np.array([[0.0, 1.0, 0.5]])
    phi = O.feature_map(x)
    assert np.allclose(phi[0, 0], [0.0, 1.0]) and np.allclose(phi[0, 1], [1.0, 0.0], atol=1e-15)
    assert np.allclose(phi[0, 2], [np.sqrt(0.5)] * 2)


def test_reference_rule_reproduces_L_gt_2_crash():
    """the unmodified rule cannot finish a right sweep for L>2."""
    np.random.seed(0)
    S, M, L, Ns = 6, 4, 3, 16
    X = O.feature_map(np.random.random((Ns, S)))
    y = np.random.randint(0, L, Ns)
    net = O.OracleMPS.from_seed(S, M, L, calibration_X=X, normalize=True, act_fn="linear", loss_fn="MSE")
    f = net.forward(X)
    with pytest.raises(ValueError, match="not aligned"):
        net.sweep(y, f, 0.01, 0.0, L2_flag=False, left_dir=False)
