"""CPU: the C-ABI library loads and exports every symbol include/tnml.h declares (no compute calls), and the
host-side mirror of the reference's bookkeeping types behaves like the reference's."""
import contextlib
import io
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/TensorNetwork"


def test_library_exports_every_declared_symbol():
    from tensornetworkforml_b200 import _lib
    header = open(os.path.join(ROOT, "include", "tnml.h")).read()
    declared = set(re.findall(r"\b(tnml_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 22
    handle = _lib.lib()
    for name in declared:
        assert hasattr(handle, name), "libtnml.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree: %s" % (declared ^ set(_lib.SIGNATURES))
    assert handle.tnml_version() >= 100
    assert handle.tnml_error_string(0) == b"ok"
    # pure host-side queries are safe without a GPU
    # (the query has no dtype argument: it covers the FP64 split-K partials and the FP32 variant's)
    assert handle.tnml_grad_workspace_bytes(60000, 64, 64, 10) >= 14 * 64 * 4 * 10 * 64 * 8
    assert handle.tnml_svd_split_workspace_bytes(64, 64, 10, 0) > 128 * 1280 * 8


def test_compute_paths_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import tensornetworkforml_b200 as tn
    np.random.seed(0)
    net = tn.Network(N=5, M=3, L=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.forward(np.zeros((4, 5, 2)))
    a = tn.Tensor(elem=np.ones((2, 3)), axes_names=["x", "k"])
    b = tn.Tensor(elem=np.ones((3, 2)), axes_names=["k", "y"])
    with pytest.raises(RuntimeError, match="GPU"):
        tn.contract(a, b, contracted="k")


def test_tensor_aggregate_disaggregate_transpose_add():
    from tensornetworkforml_b200 import Tensor
    rng = np.random.default_rng(0)
    e = rng.standard_normal((2, 3, 4, 5))
    T = Tensor(elem=e.copy(), axes_names=["d1", "left", "right", "l"])
    T.aggregate(axes_names=["d1", "left"], new_ax_name="i")
    T.aggregate(axes_names=["right", "l"], new_ax_name="j")
    T.transpose(["i", "j"])
    assert T.shape == (6, 20) and list(T.axes_names) == ["i", "j"]
    assert np.array_equal(T.elem, e.reshape(6, 20))
    assert T.aggregations["i"] == {"d1": 2, "left": 3}
    T.disaggregate("j")
    assert list(T.axes_names) == ["right", "l", "i"] and T.shape == (4, 5, 6)
    T.disaggregate("i")
    assert list(T.axes_names) == ["d1", "left", "right", "l"] and np.array_equal(T.elem, e)
    A = Tensor(elem=e.copy(), axes_names=["a", "b", "c", "d"])
    B = Tensor(elem=np.transpose(e, (3, 2, 1, 0)).copy(), axes_names=["d", "c", "b", "a"])
    assert np.array_equal((A + B).elem, 2 * e) and np.array_equal((A - B).elem, 0 * e)
    assert list(B.axes_names) == ["a", "b", "c", "d"]          # right operand permuted in place, like TC:282
    with pytest.raises(AssertionError):
        A + Tensor(elem=e, axes_names=["a", "b", "c", "z"])
    with pytest.raises(Exception):
        Tensor()
    with pytest.raises(ValueError):
        A.aggregate(axes_names=["a"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted")
def test_tensor_and_data_generator_match_live_reference():
    """Same seeds -> same weights, same data, same bookkeeping as the reference's own classes."""
    import importlib.util
    def load(name):
        spec = importlib.util.spec_from_file_location("_ref_" + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        saved = {k: sys.modules.get(k) for k in ("Tensor_class",)}
        sys.path.insert(0, REF)
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.path.remove(REF)
        return mod
    from tensornetworkforml_b200 import Tensor
    import tensornetworkforml_b200.data_generator as gen
    RT = load("Tensor_class")
    np.random.seed(3)
    a = RT.Tensor(shape=[4, 3, 2], axes_names=["left", "right", "d1"], scale=1.7)
    np.random.seed(3)
    b = Tensor(shape=[4, 3, 2], axes_names=["left", "right", "d1"], scale=1.7)
    assert np.array_equal(a.elem, b.elem)
    for T in (a, b):
        T.aggregate(axes_names=["d1", "left"], new_ax_name="i")
    assert np.array_equal(a.elem, b.elem) and list(a.axes_names) == list(b.axes_names)
    assert {k: int(v) for k, v in a.aggregations["i"].items()} == {k: int(v) for k, v in b.aggregations["i"].items()}
    for T in (a, b):
        T.disaggregate("i")
    assert np.array_equal(a.elem, b.elem) and list(a.axes_names) == list(b.axes_names)
    RG = load("data_generator")
    np.random.seed(0)
    d0, l0 = RG.create_dataset(50, 6, 0.7)
    np.random.seed(0)
    d1, l1 = gen.create_dataset(50, 6, 0.7)
    assert np.array_equal(d0, d1) and np.array_equal(l0, l1)
    import torch
    torch.manual_seed(0)
    r = RG.prepare_dataset(d0, l0, 1, 0.2, 8, 4, 4)
    torch.manual_seed(0)
    m = gen.prepare_dataset(d1, l1, 1, 0.2, 8, 4, 4)
    for lr_, lm in zip(r, m):
        assert len(lr_) == len(lm)
    torch.manual_seed(1)                     # the sampler draws its permutation when the iterator is created
    br = next(iter(r[0]))
    torch.manual_seed(1)
    bm = next(iter(m[0]))
    assert all(np.array_equal(x[0], z[0]) and x[1] == z[1] for x, z in zip(br, bm))


def test_shipped_reference_pickle_loads_into_our_classes():
    """module names Network_class / Tensor_class resolve to this package."""
    path = os.path.join(REF, "trained_diag_model.dat")
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    import pickle
    import tensornetworkforml_b200 as tn
    for m in ("Network_class", "Tensor_class"):
        mod = sys.modules.get(m)
        if mod is not None and getattr(mod, "__file__", "").startswith(REF):
            pytest.skip("reference modules already imported in this process")
    with open(path, "rb") as fh:
        net = pickle.load(fh)
    assert isinstance(net, tn.Network) and net.N == 64 and net.L == 2 and net.l_pos == 63
    assert len(net.As) == 64 and isinstance(net.As[0], tn.Tensor)


def test_reduce_gradient_with_count_written_by_the_kernel():
    """tnml_act_lossder writes [n_correct, sum|y-f|, count, 0] itself; the host must then leave the extras alone."""
    import torch
    from tensornetworkforml_b200.parallel import N_EXTRA, reduce_gradient_and_metrics
    n = 5
    buf = torch.arange(n + N_EXTRA, dtype=torch.float64)
    buf[n + 2], buf[n + 3] = 123.0, 0.0                       # what the kernel wrote
    out = reduce_gradient_and_metrics(buf.clone(), n, 999, world=1, count_written=True)
    assert torch.equal(out, buf)
    out = reduce_gradient_and_metrics(buf.clone(), n, 999, world=1)            # legacy path: host fills count / spare
    assert out[n + 2] == 999.0 and out[n + 3] == 0.0 and torch.equal(out[:n + 2], buf[:n + 2])


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times) prints one JSON line with the contract's keys; it
    needs no GPU.  Two bond updates at the full Ns = 60000 on the host cores."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "bond_updates_per_s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "bond-updates/s" and line["config"]["Ns"] == 60000
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_index_batches_reproduce_the_loader_and_its_rng_consumption():
    """Network.train's resident path asks the loader's sampler for INDICES (DG:187-192): same batches and same torch RNG
    state afterwards as iterating the loader itself, for the samplers prepare_dataset builds and for a generic one."""
    import torch
    from torch.utils.data import DataLoader
    import tensornetworkforml_b200.data_generator as gen
    from tensornetworkforml_b200.Network_class import Network
    np.random.seed(0)
    data, label = gen.create_dataset(100, 4, 0.7)

    def loaders():
        torch.manual_seed(0)
        tl, vl, _ = gen.prepare_dataset(data, label, 1, 0.2, 24, 8, 8)
        sh = DataLoader(tl.dataset, 16, shuffle=True, collate_fn=lambda b: b)      # RandomSampler: generic path
        sq = DataLoader(tl.dataset, 30, collate_fn=lambda b: b)                    # SequentialSampler, ragged tail
        return [tl, vl, sh, sq]

    def position(ds, x):
        return int(np.where((ds.data == x).all(axis=(1, 2)))[0][0])

    for k in range(4):
        ld = loaders()[k]
        real = [[position(ld.dataset, s[0]) for s in b] for b in ld]
        st_real = torch.get_rng_state()
        ld = loaders()[k]
        mine = [[int(i) for i in b] for b in Network._index_batches(ld)]
        assert mine == real and torch.equal(torch.get_rng_state(), st_real), "loader %d" % k


def test_warm_split_backoff_policy():
    """Host logic of the warm-started split (engine.warm_feedback + the wait counter consumed in split_phase): one free
    retry after a refusal; after a second refusal in a row two visits on the cold pipeline (generic form) or 4, 8, 16, 16
    ... visits (single-CTA form, where a refusal is ten times as expensive as sitting out); reset by an accepted attempt."""
    from tensornetworkforml_b200.engine import warm_feedback

    def run(single_cta, outcomes):
        """-> per visit whether it attempts the fast path; outcomes = what the gates say at each attempt"""
        fails, waits, key, tried = {}, {}, (5, 0, 64, 64), []
        outcomes = iter(outcomes)
        while True:
            if waits.get(key, 0) > 0:
                waits[key] -= 1
                tried.append(False)
                continue
            try:
                warm_feedback(fails, waits, key, next(outcomes), single_cta=single_cta)
            except StopIteration:
                return tried, fails
            tried.append(True)

    T, F = True, False
    tried, fails = run(False, [F, F, F, T, F, T, T])
    #               refused, free retry refused, 2 out, refused, 2 out, accepted, single refusal, accepted twice
    assert tried == [T, T, F, F, T, F, F, T, T, T, T] and fails == {}
    tried, _ = run(True, [F, F, F, F, F, F])
    gaps = [len(g) for g in "".join("x" if t else "." for t in tried).split("x")[1:]]
    assert gaps == [0, 4, 8, 16, 16, 16]              # free retry, then exponential, capped
    tried, fails = run(True, [F, F, T, F, F])         # an accepted attempt starts the count again
    assert tried == [T, T] + [F] * 4 + [T, T, T] + [F] * 4 and fails[(5, 0, 64, 64)] == 2


def test_truncation_rule_of_the_engine_equals_the_oracle():
    """engine.choose_m (host logic in front of tnml_svd_split) against the oracle's restatement of NC:898-910 /
    NC:933-945 for every label position, direction and shape class -- including WHERE each of them raises (the
    reference's np.dot failures for L > 2)."""
    from tensornetworkforml_b200.engine import choose_m as mine
    from oracle.mps_oracle import choose_m as theirs
    S = 7

    def outcome(fn, *a):
        try:
            return ("ok", fn(*a))
        except ValueError as e:
            return ("raises", str(e))

    n = 0
    for rule, max_bond in (("reference", None), ("fixed", 3), ("fixed", 6), ("fixed", 64)):
        for left_dir in (False, True):
            for l_pos in range(S):
                for Dl in (1, 2, 3, 5):
                    for Dr in (1, 2, 4):
                        for L in (2, 3, 10):
                            R, C = (2 * Dl, 2 * L * Dr) if not left_dir else (2 * Dl * L, 2 * Dr)
                            a = outcome(mine, rule, max_bond, left_dir, l_pos, S, Dl, R, C)
                            b = outcome(theirs, rule, left_dir, l_pos, S, Dl, min(R, C), R, C, max_bond)
                            assert a == b, (rule, max_bond, left_dir, l_pos, Dl, Dr, L, a, b)
                            n += 1
    assert n == 4 * 2 * S * 4 * 3 * 3
    # 'adaptive' asks the split for the cap; the data-dependent cut follows it (engine.split_phase)
    assert mine("adaptive", 6, False, 2, S, 4, 8, 80) == mine("fixed", 6, False, 2, S, 4, 8, 80) == 6


def test_site_and_bond_layout_conversions_round_trip():
    """Host layout glue between the reference's named-axis Tensors and the canonical device layouts (pure NumPy, no
    device): a site Tensor in ANY axis order maps to (Dl,2,[L,]Dr) and back, the chain ends lose / regain their dummy
    bond, and a bond Tensor B (NC:484) in any axis order maps to (Dl,2,L,2,Dr) and back into the caller's axis order."""
    import contextlib
    import io
    import itertools
    from tensornetworkforml_b200 import Network_class as NCm
    from tensornetworkforml_b200.Tensor_class import Tensor
    rng = np.random.default_rng(5)
    S, Dl, Dr, L = 6, 3, 4, 5
    for p, is_label in itertools.product((0, 2, S - 1), (False, True)):
        shape = (1 if p == 0 else Dl, 2) + ((L,) if is_label else ()) + (1 if p == S - 1 else Dr,)
        A = rng.standard_normal(shape)
        T = NCm._canonical_to_named(A, p, S, is_label)
        want = ([] if p == 0 else ["left"]) + ["d%d" % p] + (["l"] if is_label else []) + ([] if p == S - 1 else ["right"])
        want_shape = shape[1 if p == 0 else 0:len(shape) - (1 if p == S - 1 else 0)]     # dummy edge bonds dropped
        assert [str(n) for n in T.axes_names] == want and T.elem.shape == want_shape
        for perm in itertools.permutations(range(T.elem.ndim)):
            Tp = Tensor(elem=np.transpose(T.elem, perm), axes_names=[want[i] for i in perm])
            assert np.array_equal(NCm._named_to_canonical(Tp, p), A)
    with contextlib.redirect_stdout(io.StringIO()):
        net = NCm.Network(N=S, M=3, L=L)              # no calibration: nothing touches the device
    for p in (0, 2, S - 2):
        names = net._bond_names(p)
        shape = [1 if p == 0 else Dl, 2, L, 2, 1 if p == S - 2 else Dr]
        Bc = rng.standard_normal(shape)
        keep = [i for i, nm in enumerate(names) if not ((nm == "left" and p == 0) or (nm == "right" and p == S - 2))]
        base = Bc.reshape([shape[i] for i in keep])
        for perm in itertools.permutations(range(len(keep))):
            B = Tensor(elem=np.transpose(base, perm), axes_names=[names[keep[i]] for i in perm])
            can = net._bond_to_canonical(B, p)
            assert can.shape == tuple(shape) and np.array_equal(can, Bc)
            back = net._bond_from_canonical(can * 2.0, B, p)
            assert [str(n) for n in back.axes_names] == [str(n) for n in B.axes_names]
            assert np.array_equal(back.elem, 2.0 * B.elem)
    with pytest.raises(AssertionError):
        net._bond_to_canonical(Tensor(elem=np.zeros((2, 2)), axes_names=["d1", "bogus"]), 1)


def test_bench_clock_sampler_parses_nvidia_smi_lines():
    """bench.py's `clocks` block (median SM clock under load, max clock, throttle reasons seen DURING the timed region) from
    the CSV lines `nvidia-smi --query-gpu=... -lms 100` writes; without nvidia-smi the block is empty, not an error."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.ClockSampler(None).stop() == dict(sm_mhz=None, sm_max_mhz=None, reasons=[])

    class _Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    smp = bench.ClockSampler(None)
    smp.proc, smp.fh = _Proc(), open(smp.path, "w")
    smp.fh.write("0, 1965, 1965, 540.1, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active\n"
                 "0, 1950, 1965, 990.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active\n"
                 "0, 1935, 1965, 995.2, 0x0000000000000004, Not Active, Not Active, Not Active, Active\n"
                 "0, [N/A], 1965, 1.0, 0x0, Not Active, Not Active, Not Active, Not Active\n"
                 "garbage line\n")
    smp.fh.flush()
    out = smp.stop()
    assert out == dict(sm_mhz=1950.0, sm_max_mhz=1965.0, reasons=["sw_power_cap"], samples=3)
    assert not os.path.exists(smp.path)
    cfg = bench.workload_config(196, 10, 64, 60000, "f64")
    assert cfg["workload"].startswith("config3:") and (cfg["S"], cfg["L"], cfg["D"], cfg["Ns"]) == (196, 10, 64, 60000)
    assert bench.workload_config(784, 10, 128, 60000, "f64")["workload"].startswith("variant of config3: 28x28")


def test_kept_ratio_of_a_recorded_split():
    """engine.kept_ratio: (sigma_m / sigma_1)^2 from a row of the sweep's singular-value record -- what the opt-in rule
    TNML_FAST_MIN_RATIO compares (a bond graded below it at its previous visit does not attempt the warm-started split)."""
    from tensornetworkforml_b200.engine import kept_ratio
    row = np.full(260, np.nan)
    row[:128] = np.concatenate([np.linspace(2.0, 0.5, 64), 1e-6 * np.ones(64)])
    assert kept_ratio(row, 128) == pytest.approx((0.5 / 2.0) ** 2)
    assert np.isnan(kept_ratio(np.full(260, np.nan), 128))              # nothing recorded
    assert np.isnan(kept_ratio(np.zeros(260), 128))                     # sigma_1 = 0
    assert np.isnan(kept_ratio(row[:10], 128)) and np.isnan(kept_ratio(row, 0))
