"""CPU, world_size 2 over gloo: the host-side logic of the sample-sharded sweep (tensornetworkforml_b200/parallel.py).
The sharded gradient/metric sums must equal the single-process sums; uses the oracle's gradient as the per-shard
payload so the test exercises the same [dB | n_correct | sum|y-f| | count] packing the engine uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mps_oracle as O
from tensornetworkforml_b200 import parallel as P


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    rng = np.random.default_rng(0)
    Ns, a, c, L = 101, 3, 4, 3
    Le, Re = rng.standard_normal((Ns, a)), rng.standard_normal((Ns, c))
    pa, pb = O.feature_map(rng.random(Ns)), O.feature_map(rng.random(Ns))
    f = rng.standard_normal((Ns, L))
    y = rng.integers(0, L, Ns)
    return Ns, L, Le, Re, pa, pb, f, y


def _worker(rank, world, port, out, count_written=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Ns, L, Le, Re, pa, pb, f, y = _problem()
    lo, hi = P.shard_bounds(Ns, rank, world)
    y1h = np.eye(L)[y[lo:hi]]
    fa = O.apply_act(f[lo:hi], "linear", 0.1)
    g = O.loss_derivative(fa, y1h, "linear", "MSE", 0.1)
    dB = O.gradient(g, Le[lo:hi], pa[lo:hi], pb[lo:hi], Re[lo:hi]).reshape(-1)
    n = dB.size
    buf = torch.zeros(n + P.N_EXTRA + 3, dtype=torch.float64)
    buf[:n] = torch.from_numpy(dB)
    buf[n] = float((np.argmax(fa, 1) == y[lo:hi]).sum())
    buf[n + 1] = float(np.abs(y1h - fa).sum())
    buf[n + P.N_EXTRA:] = 7.0                                   # bytes past the payload must not be touched
    if count_written:                                           # what tnml_act_lossder writes on the device
        buf[n + 2], buf[n + 3] = float(hi - lo), 0.0
        P.reduce_gradient_and_metrics(buf, n, -1, world=world, count_written=True)
    else:
        P.reduce_gradient_and_metrics(buf, n, hi - lo, world=world)
    fmax = P.global_abs_max(np.abs(f[lo:hi]).max(), "cpu", world=world)
    out[rank] = (buf.numpy().copy(), fmax, (lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("count_written", [False, True])
def test_sharded_sums_equal_single_process_sums(count_written):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, count_written), nprocs=world, join=True)
    Ns, L, Le, Re, pa, pb, f, y = _problem()
    y1h = np.eye(L)[y]
    g = O.loss_derivative(f, y1h, "linear", "MSE", 0.1)
    dB = O.gradient(g, Le, pa, pb, Re).reshape(-1)
    n = dB.size
    acc, mae = O.metrics(f, y1h)
    bounds = sorted(out[r][2] for r in range(world))
    assert bounds[0][0] == 0 and bounds[-1][1] == Ns and bounds[0][1] == bounds[1][0]
    for r in range(world):
        buf, fmax, _ = out[r]
        assert np.abs(buf[:n] - dB).max() < 1e-12 * np.abs(dB).max()
        a, m = P.metrics_from_sums(buf[n], buf[n + 1], buf[n + 2], L)
        assert buf[n + 2] == Ns and abs(a - acc) < 1e-15 and abs(m - mae) < 1e-14
        assert np.all(buf[n + P.N_EXTRA:] == 7.0)
        assert fmax == np.abs(f).max()
    assert np.array_equal(out[0][0], out[1][0])                  # every rank ends with the same bits


@pytest.mark.parametrize("Ns,world", [(10, 3), (60000, 8), (7, 8), (1, 2)])
def test_shard_bounds_partition(Ns, world):
    edges = [P.shard_bounds(Ns, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == Ns
    assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_bounds(Ns, world, world)
