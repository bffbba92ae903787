"""Multi-GPU parity check (run under torchrun): the sample-sharded sweep (dB + metrics all-reduced over NCCL) must
reproduce the single-GPU sweep over the whole batch.  Prints max relative deviations on rank 0.

Default = teacher forcing: before every sweep the sharded network is reset to the solo network's weights, so each
sweep is compared from an identical state and the bar is 1e-10 on f, singular values and MAE (BASELINE.json).
--free: the two runs evolve independently; the summation-order difference of the all-reduce (1e-16) is amplified by
the chaotic iteration (~30x per sweep), the bar is then 1e-8 after three sweeps."""
import contextlib, io, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensornetworkforml_b200 as tn

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", lr_))
S, D, L, Ns = 16, 16, 10, 4096
np.random.seed(1)
x = np.random.random((Ns, S)); X = np.stack((np.sin(np.pi * x / 2), np.cos(np.pi * x / 2)), -1)
y = np.random.randint(0, L, Ns)
lo, hi = rank * Ns // world, (rank + 1) * Ns // world
kw = dict(N=S, M=D, L=L, normalize=True, act_fn="linear", loss_fn="MSE", truncation="fixed", max_bond=D)
state = np.random.get_state()
with contextlib.redirect_stdout(io.StringIO()):
    dist_net = tn.Network(calibration_X=X[lo:hi], **kw)                  # sharded: calibration max is all-reduced
np.random.set_state(state)
with contextlib.redirect_stdout(io.StringIO()):
    solo_net = tn.Network(calibration_X=X, process_group=False, **kw)    # whole batch on this GPU, no communication
np.random.set_state(state)
with contextlib.redirect_stdout(io.StringIO()):
    solo2 = tn.Network(calibration_X=X, process_group=False, **kw)      # determinism probe: must match solo_net bitwise
FREE = "--free" in sys.argv
TOL = 1e-8 if FREE else 1e-10
worst = 0.0
for sweep in range(3):
    if not FREE and sweep > 0:
        import copy
        with contextlib.redirect_stdout(io.StringIO()):
            dist_net.As = copy.deepcopy(solo_net.As)                     # teacher forcing: same state on both sides
            dist_net.l_pos = solo_net.l_pos
    f2 = solo2.forward(X)
    f2 = solo2.sweep(X, y, f2, 0.005, 1e-2, left_dir=(solo2.l_pos == S - 1))
    fd, fs = dist_net.forward(X[lo:hi]), solo_net.forward(X)
    left = solo_net.l_pos == S - 1
    vd, vs = [[], []], [[], []]
    fd = dist_net.sweep(X[lo:hi], y[lo:hi], fd, 0.005, 1e-2, left_dir=left, var_hist=vd)
    fs = solo_net.sweep(X, y, fs, 0.005, 1e-2, left_dir=left, var_hist=vs)
    e_f = np.abs(fd.elem - fs.elem[:, lo:hi]).max() / np.abs(fs.elem).max()
    e_m = max(np.abs(np.array(vd[0]) - np.array(vs[0])).max(), np.abs(np.array(vd[1]) - np.array(vs[1])).max())
    sv_d, sv_s = dist_net.last_history["svals"], solo_net.last_history["svals"]
    e_s = max(np.abs(a - b).max() / b.max() for a, b in zip(sv_d, sv_s))
    t = torch.tensor([e_f, e_m, e_s], device="cuda", dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    worst = max(worst, float(t.max()))
    if rank == 0:
        print("  solo run-to-run bitwise identical:", bool(np.array_equal(f2.elem, fs.elem) and np.isfinite(fs.elem).all()))
        print("sweep %d %s: max rel dev f %.2e, metrics %.2e, singular values %.2e" % (sweep, "L" if left else "R", *t.tolist()))
# replicas must hold bitwise identical tensors (replicated SVD, no broadcast)
sites = torch.cat([s.reshape(-1) for s in dist_net._eng.sites])
ref = sites.clone()
torch.distributed.broadcast(ref, 0)
same = bool(torch.equal(sites, ref))
flag = torch.tensor([1.0 if same else 0.0], device="cuda", dtype=torch.float64)
torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
if rank == 0:
    print("replicas bitwise identical:", bool(flag.item()), "| worst deviation", worst)
    print("mode:", "free-running" if FREE else "teacher forcing", "| tolerance", TOL, "| world", world)
    assert worst < TOL and flag.item() == 1.0
    print("DIST_CHECK_OK")
torch.distributed.destroy_process_group()
