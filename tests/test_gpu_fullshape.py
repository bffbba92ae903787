"""GPU: CUDA path against the oracle AT THE BASELINE SHAPE (config 3: Ns = 60 000, D = 64, L = 10; NC:440-573), a
fixed-D sweep at D = 128, and the warm-started split inside real sweeps.  Tolerance 1e-10 (FP64, BASELINE.json)."""
import contextlib
import io

import numpy as np
import pytest

from oracle import mps_oracle as O
from tests import _golden as G

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-10


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


@pytest.fixture(scope="module")
def tn():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tensornetworkforml_b200 as pkg
    return pkg


def test_three_interior_bond_updates_at_config3_shape(tn):
    """Three consecutive interior bond updates (left/right bond 64, 10 labels) over all 60 000 samples: prediction
    f_new, singular values, accuracy / MAE sums and the product of the two new site tensors against
    OracleMPS.sweep_step (about 0.5 s of NumPy per update).  Same construction as bench.cpu_bond_updates."""
    from tensornetworkforml_b200.engine import SweepEngine
    Ns, D, L, n, lr, wd = 60000, 64, 10, 3, 1e-4, 1e-3
    rng = np.random.default_rng(0)
    S, p0 = n + 4, 1
    sites = [rng.standard_normal((D, 2, D)) / np.sqrt(2 * D) for _ in range(S)]
    sites[0] = rng.standard_normal((1, 2, D)) / np.sqrt(2)
    sites[-1] = rng.standard_normal((D, 2, 1)) / np.sqrt(2 * D)
    sites[p0] = rng.standard_normal((D, 2, L, D)) / np.sqrt(2 * D)
    X = O.feature_map(rng.random((Ns, S)))
    envs = {p: rng.standard_normal((Ns, D)) / np.sqrt(D) for p in range(p0 + 2, S)}
    y = rng.integers(0, L, Ns)
    f0 = rng.standard_normal((Ns, L)) * 0.1

    orc = O.OracleMPS(sites, L, act_fn="linear", loss_fn="MSE", rule="fixed", max_bond=D, l_pos=p0)
    orc.phi = X
    orc.env = [None] * (S + 1)
    orc.env[0] = np.ones((Ns, 1))
    orc.env[S] = np.ones((Ns, 1))
    for p, e in envs.items():
        orc.env[p] = e.copy()
    orc._build_norm_stack(False)
    # the left norm environment of the starting position (the oracle's stack assumes a sweep that started at site 0)
    orc._norm[p0 - 1] = np.ones((1, 1))
    y1h = np.zeros((Ns, L))
    y1h[np.arange(Ns), y] = 1

    eng = SweepEngine(S, L, 0.1, "linear", "MSE", rule="fixed", max_bond=D)
    eng.set_sites(sites, p0)
    eng.load_input(X)
    for p, e in envs.items():
        eng.env[p, :Ns * D].copy_(torch.from_numpy(e.reshape(-1)))
    eng.f_buf[0].copy_(torch.from_numpy(f0.reshape(-1)))
    eng.f_cur = 0
    eng.begin_sweep(y, False, True)

    fo = f0
    for step in range(n):
        fo = orc.sweep_step(fo, y1h, lr, wd, True, False)
        f = eng.sweep_step(lr, wd, True, False).cpu().numpy()
        assert G.rel(f, fo) < TOL, "f_new, step %d" % step
        p = p0 + step
        prod = np.einsum("asm,mtlc->asltc", eng.sites[p].cpu().numpy().reshape(D, 2, -1),
                         eng.sites[p + 1].cpu().numpy().reshape(-1, 2, L, D))
        want = np.einsum("asm,mtlc->asltc", orc.sites[p], orc.sites[p + 1])
        # the left bond a of this pair carries the sign gauge of the previous split (u_k, v_k -> -u_k, -v_k): align it
        sgn = np.sign(np.einsum("asltc,asltc->a", prod, want))
        assert G.rel(prod * sgn[:, None, None, None, None], want) < TOL, "A_p' A_q', step %d" % step
    h = eng.history()
    for step in range(n):
        ref = orc.hist[step]
        assert h["acc"][step] == ref["acc"] and abs(h["mae"][step] - ref["mae"]) < TOL
        assert np.abs(h["svals"][step] - ref["S"]).max() / ref["S"].max() < TOL
        assert abs(h["stats"][step, 2] - ref["l2_loss"]) < TOL * max(1.0, abs(ref["l2_loss"]))


def test_fixed_bond_128_sweeps_match_oracle(tn):
    """Bond dimension 128 (config 4's): two sweeps of a 16-site chain against the oracle; the interior splits are
    256 x 768 (cluster Jacobi at n = 256), the contractions run in two 64-wide chunks per side."""
    S, D, Lbl, Ns, lr, wd = 16, 128, 3, 192, 0.02, 0.01
    np.random.seed(31)
    X = O.feature_map(np.random.random((Ns, S)))
    y = np.random.randint(0, Lbl, Ns)
    state = np.random.get_state()
    orc = O.OracleMPS.from_seed(S, D, Lbl, calibration_X=X, normalize=True, act_fn="linear", loss_fn="MSE", rule="fixed",
                                max_bond=D)
    np.random.set_state(state)
    with quiet():
        net = tn.Network(N=S, M=D, L=Lbl, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE",
                         truncation="fixed", max_bond=D)
    for sw in range(2):
        fo, f = orc.forward(X), net.forward(X)
        assert G.rel(f.elem.T, fo) < TOL
        left = orc.l_pos == S - 1
        n0 = len(orc.hist)
        fo = orc.sweep(y, fo, lr, wd, L2_flag=True, left_dir=left)
        f = net.sweep(X, y, f, lr, wd, L2_flag=True, left_dir=left)
        assert G.rel(f.elem.T, fo) < TOL, "sweep %d" % sw
        for mine, ref in zip(net.last_history["svals"], [r["S"] for r in orc.hist[n0:]]):
            assert np.abs(mine[:len(ref)] - ref).max() / ref.max() < TOL
        assert net._eng.bond_dims() == orc.bond_dims()
    assert max(net._eng.bond_dims()) == 128


def test_warm_started_split_inside_sweeps(tn):
    """Six sweeps of a 16-site chain with config-3 bond dimensions (D = 64, L = 10) on the bench's stripe data.  From the third sweep on the interior splits
    start from the previous visit's basis (csrc/svd_fast.cuh).  (1) Every sweep with the fast path equals the same sweep without
    it (cold pipeline, restarted from the same tensors) to 1e-10 in f and singular values; (2) both follow the oracle
    (free-running: 1e-7 after six sweeps of a chaotic iteration -- a 1e-15 perturbation reaches 5e-8 after six
    reference sweeps); (3) the fast path was taken."""
    import tensornetworkforml_b200.data_generator as gen
    side, D, Lbl, Ns, lr, wd = 4, 64, 10, 2048, 1e-4, 1e-3          # the bench workload (stripe templates) in small
    S = side * side
    np.random.seed(2)
    data, labels = gen.create_multiclass_dataset(Ns, side, Lbl, 0.7)
    X, y = gen.psi(data.reshape(Ns, -1)), labels.astype(np.int64)
    np.random.seed(2)
    state = np.random.get_state()
    orc = O.OracleMPS.from_seed(S, D, Lbl, calibration_X=X, normalize=True, act_fn="linear", loss_fn="MSE", rule="fixed",
                                max_bond=D)
    nets = []
    for warm in (True, False):
        np.random.set_state(state)
        with quiet():
            net = tn.Network(N=S, M=D, L=Lbl, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE",
                             truncation="fixed", max_bond=D)
        net._engine().warm_split = warm
        nets.append(net)
    fast_taken = 0
    for sw in range(6):
        fo = orc.forward(X)
        left = orc.l_pos == S - 1
        n0 = len(orc.hist)
        fo = orc.sweep(y, fo, lr, wd, L2_flag=True, left_dir=left)
        if sw > 0:      # the cold run restarts every sweep from the fast run's tensors: a per-sweep comparison from identical
            import copy  # states (free-running, the two would drift apart like any two roundings of a chaotic iteration)
            with quiet():
                nets[1].As = copy.deepcopy(nets[0].As)
                nets[1].l_pos = nets[0].l_pos
        outs = []
        for net in nets:
            f = net.forward(X)
            f = net.sweep(X, y, f, lr, wd, L2_flag=True, left_dir=left)
            outs.append((f.elem.T.copy(), net.last_history["svals"]))
        assert G.rel(outs[0][0], outs[1][0]) < TOL, "fast vs cold, sweep %d" % sw
        for a, b in zip(outs[0][1], outs[1][1]):
            assert np.abs(a - b).max() / b.max() < TOL
        assert G.rel(outs[0][0], fo) < 1e-7, "vs oracle, sweep %d" % sw
        for mine, ref in zip(outs[0][1], [r["S"] for r in orc.hist[n0:]]):
            assert np.abs(mine[:len(ref)] - ref).max() / ref.max() < 1e-7
        eng = nets[0]._eng
        sv = eng.hist["svals"][:eng.hist["n"]].cpu().numpy()
        fast_taken += sum(1 for i, n in enumerate(eng.hist["nsv"]) if n == 128 and sv[i, n] >= 100)
    assert fast_taken >= 4, "the deflation path never engaged (%d splits)" % fast_taken
