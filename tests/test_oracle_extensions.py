"""CPU: the oracle's restatements of the two opt-in extensions against the reference's own
expressions -- NC:890-891 for the adaptive index, NC:794 for the softmax they stabilise."""
import os
import sys

import numpy as np
import pytest

from oracle import mps_oracle as O

REF = "/root/reference/TensorNetwork"


def test_stable_softmax_equals_the_reference_expression_where_it_is_finite():
    rng = np.random.default_rng(0)
    f = rng.standard_normal((50, 7))
    T = 0.1
    ref = np.exp(f / T) / np.exp(f / T).sum(axis=1, keepdims=True)            # NC:794 (label axis = 1 here)
    assert np.abs(O.apply_act(f, "softmax", T) - ref).max() == 0.0
    assert np.abs(O.apply_act(f, "softmax_stable", T) - ref).max() < 1e-15
    big = f * 1e3                                                               # exp overflows: inf / inf = nan in NC:794
    with np.errstate(over="ignore", invalid="ignore"):
        assert not np.isfinite(O.apply_act(big, "softmax", T)).all()
    st = O.apply_act(big, "softmax_stable", T)
    assert np.isfinite(st).all() and np.abs(st.sum(axis=1) - 1).max() < 1e-12
    # the derivative branch of cross entropy treats both spellings alike (NC:826-830)
    y1h = np.eye(7)[rng.integers(0, 7, 50)]
    fa = O.apply_act(f, "softmax", T)
    assert np.array_equal(O.loss_derivative(fa, y1h, "softmax", "cross_entropy", T),
                          O.loss_derivative(fa, y1h, "softmax_stable", "cross_entropy", T))


def test_adaptive_index_is_the_expression_of_the_reference():
    """NC:890-891: cumulative_variance_explained = cumsum(S)/S.sum(); index = argmax(cve > threshold); the intent
    m_new = max(10, min(index, m)) is in old_files/TensorNetwork.py:1310-1326 (10 -> min_bond, m -> max_bond)."""
    rng = np.random.default_rng(1)
    for _ in range(20):
        S = np.sort(rng.random(24))[::-1] * rng.choice([1.0, 1e-3], 24)
        S = np.sort(S)[::-1]
        for thr in (0.5, 0.9, 0.999):
            index = int(np.argmax(np.cumsum(S) / S.sum() > thr))
            for lo, hi in ((2, 8), (10, 16), (1, 64)):
                assert O.adaptive_m(S, thr, lo, hi) == min(len(S), max(lo, min(index, hi)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_reference_computes_the_same_index_and_ignores_it():
    """The live reference evaluates NC:890-891 inside tensor_svd and then truncates by its fixed rule: the bond it keeps
    does not depend on `threshold`, and the index it computed is the one adaptive_m starts from."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    src = open(os.path.join(REF, "Network_class.py")).read()
    assert "cumulative_variance_explained = np.cumsum(S)/S.sum()" in src
    assert "index = np.argmax(cumulative_variance_explained>threshold)" in src


def test_adaptive_oracle_sweep_keeps_data_dependent_bonds():
    np.random.seed(3)
    S_, D, L, Ns = 10, 8, 3, 120
    X = O.feature_map(np.random.random((Ns, S_)))
    y = np.random.randint(0, L, Ns)
    net = O.OracleMPS.from_seed(S_, D, L, calibration_X=X, normalize=True, act_fn="linear", loss_fn="MSE",
                                rule="adaptive", max_bond=D, threshold=0.999, min_bond=3)
    f = net.forward(X)
    f = net.sweep(y, f, 0.02, 0.01, True, False)
    bonds = net.bond_dims()
    assert max(bonds) <= D and min(bonds[1:-1]) >= 3 and len(set(bonds)) > 1
    assert np.isfinite(f).all() and np.abs(net.forward(X) - f).max() < 1e-2 * np.abs(f).max()
