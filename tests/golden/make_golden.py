"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF (imported from /root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed here as small .npz files; the
CPU tests pin oracle/mps_oracle.py against them and the GPU tests pin the CUDA path against them.

What is recorded per case: inputs (X, y), the calibrated initial site tensors (canonical layout of
oracle/mps_oracle.py), and per sweep: forward f, post-sweep f, var_hist rows (accuracy, MAE), the singular
values handed back by every np.linalg.svd call inside tensor_svd (NC:887), and the bond dimensions.
"""
import contextlib
import io
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/TensorNetwork"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import Network_class as NC                      # noqa: E402  (the reference)
from Tensor_class import Tensor                 # noqa: E402
from oracle import mps_oracle as O              # noqa: E402  (only for layout conversion + feature map)


class FixedBondReference(NC.Network):
    """The reference with ONLY tensor_svd's choice of m replaced: m = min(len(S), max_bond),
    both factors cut; same sqrt(S) split and same aggregate / disaggregate bookkeeping as NC:912-925."""
    max_bond = None

    def tensor_svd(self, T, left_dir=False, threshold=0.999):
        U, S, Vh = np.linalg.svd(np.array(T.elem, copy=True))
        m = min(len(S), int(self.max_bond))
        sq = np.sqrt(np.eye(m, m) * S[:m])
        TU = Tensor(elem=np.dot(U[:, :m], sq), axes_names=["i", "right"])
        TSVh = Tensor(elem=np.dot(sq, Vh[:m, :]), axes_names=["left", "j"])
        TU.aggregations["i"] = T.aggregations["i"]
        TSVh.aggregations["j"] = T.aggregations["j"]
        TU.disaggregate("i")
        TSVh.disaggregate("j")
        return TU, TSVh


@contextlib.contextmanager
def record_svd(store):
    real = np.linalg.svd

    def wrapped(*a, **k):
        out = real(*a, **k)
        store.append(np.array(out[1], copy=True))
        return out
    np.linalg.svd = wrapped
    try:
        yield
    finally:
        np.linalg.svd = real


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def run_case(name, S, M, L, Ns, act, loss, wd, lr, L2, nsweeps, seed, max_bond=None, data="uniform"):
    np.random.seed(seed)
    if data == "uniform":
        x = np.random.random((Ns, S))
    else:
        raise ValueError(data)
    X = O.feature_map(x)
    y = np.random.randint(0, L, Ns)
    cls = NC.Network if max_bond is None else FixedBondReference
    with quiet():
        net = cls(N=S, M=M, L=L, normalize=True, calibration_X=X, act_fn=act, loss_fn=loss)
    if max_bond is not None:
        net.max_bond = max_bond
    out = dict(X=X, y=y, S_sites=S, M=M, L=L, act=act, loss=loss, wd=wd, lr=lr, L2=int(L2), nsweeps=nsweeps,
               max_bond=-1 if max_bond is None else max_bond, T=net.T)
    for p, A in enumerate(O.sites_from_reference(net.As)):
        out["site0_%d" % p] = A
    for sw in range(nsweeps):
        with quiet():
            f = net.forward(X)
        out["f_fwd_%d" % sw] = f.elem.T.copy()
        left = net.l_pos == S - 1
        vh, svals = [[], []], []
        with quiet(), record_svd(svals):
            f = net.sweep(X, y, f, lr, wd, L2_flag=L2, left_dir=left, var_hist=vh)
        out["f_swp_%d" % sw] = f.elem.T.copy()
        out["acc_%d" % sw] = np.array(vh[0])
        out["mae_%d" % sw] = np.array(vh[1])
        out["left_%d" % sw] = int(left)
        width = max(len(s) for s in svals)
        sv = np.full((len(svals), width), np.nan)
        for i, s in enumerate(svals):
            sv[i, :len(s)] = s
        out["sv_%d" % sw] = sv
        bonds = []
        for p, A in enumerate(O.sites_from_reference(net.As)[:-1]):
            bonds.append(A.shape[-1])
        out["bonds_%d" % sw] = np.array(bonds)
    with quiet():
        out["f_final"] = net.forward(X).elem.T.copy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; bonds", out["bonds_%d" % (nsweeps - 1)])


def diag_known_answer():
    """trained_diag_model.dat known answer: sites + forward on seeded input."""
    with open(os.path.join(REF, "trained_diag_model.dat"), "rb") as fh:
        net = pickle.load(fh)
    np.random.seed(123)
    x = np.random.random((4, 64))
    X = O.feature_map(x)
    with quiet():
        f = net.forward(X)
    out = dict(X=X, f=f.elem.T.copy(), l_pos=net.l_pos, L=net.L, T=net.T, act=net.act_fn, loss=net.loss_fn, M=net.M)
    with open(os.path.join(REF, "trained_diag_model.dat"), "rb") as fh:
        net = pickle.load(fh)                       # reload: forward permutes axes of As in place (CLT:74-75)
    for p, A in enumerate(O.sites_from_reference(net.As)):
        out["site_%d" % p] = A
    path = os.path.join(HERE, "diag_model_known_answer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; f =", out["f"].T)


if __name__ == "__main__":
    # unmodified reference (L=2 only -- SURVEY.md section 0.2 fact 4)
    run_case("ref_softmax_fce_L2", S=8, M=5, L=2, Ns=48, act="softmax", loss="full_cross_ent",
             wd=1.0, lr=0.01, L2=True, nsweeps=3, seed=0)
    run_case("ref_linear_mse_decay", S=10, M=6, L=2, Ns=64, act="linear", loss="MSE",
             wd=0.01, lr=0.05, L2=False, nsweeps=3, seed=1)
    run_case("ref_sigmoid_ce_L2", S=7, M=4, L=2, Ns=40, act="sigmoid", loss="cross_entropy",
             wd=0.001, lr=0.01, L2=True, nsweeps=2, seed=2)
    # fixed-bond variant (tensor_svd's m only) -- the oracle for L>2 / retained D
    run_case("fixed_linear_mse_L3_D4", S=8, M=4, L=3, Ns=48, act="linear", loss="MSE",
             wd=0.01, lr=0.05, L2=True, nsweeps=3, seed=3, max_bond=4)
    run_case("fixed_softmax_fce_L10_D8", S=9, M=8, L=10, Ns=96, act="softmax", loss="full_cross_ent",
             wd=1e-3, lr=0.01, L2=True, nsweeps=2, seed=4, max_bond=8)
    diag_known_answer()
