"""Why do the device-resident arm and the API arm of bench.py differ?  Times one sweep (forward + 195 bond updates) of
config 3 three ways: engine loop without host syncs, engine loop with one sync per sweep, Network.forward/sweep."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tensornetworkforml_b200 as tn
dt = os.environ.get("PDT", "float64")
c = bench.CFG
S, L, D, Ns = c["S"], c["L"], c["D"], c["Ns"]
X, y = bench.synthetic_data(Ns, S, L, c["seed"])
np.random.seed(c["seed"])
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:2048], act_fn=c["act"], loss_fn=c["loss"],
                     truncation="fixed", max_bond=D, dtype=dt)
eng = net._engine()
eng.load_input(X)
yd = torch.from_numpy(y.astype(np.int32)).to(eng.device)
def dev_step():
    eng.forward(); left = eng.l_pos == S - 1
    eng.begin_sweep(yd, left, True)
    t0 = time.perf_counter()
    for _ in range(S - 1): eng.sweep_step(c["lr"], c["wd"], True, left)
    return time.perf_counter() - t0
for _ in range(3): dev_step()
torch.cuda.synchronize()
for sync in (False, True, False, True):
    t0 = time.perf_counter(); enq = 0.0
    for _ in range(3):
        enq += dev_step()
        if sync: torch.cuda.synchronize()
    torch.cuda.synchronize(); dtot = (time.perf_counter() - t0) / 3
    print("%s engine loop, sync per sweep=%s: %.1f ms/sweep (python enqueue of the 195 steps %.1f ms)" % (dt, sync, dtot * 1e3, enq / 3 * 1e3))
net.l_pos, net._host_fresh = eng.l_pos, False
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f = net.forward(X); t1 = time.perf_counter()
    f = net.sweep(X, y, f, c["lr"], c["wd"], L2_flag=True, left_dir=(net.l_pos == S - 1)); t2 = time.perf_counter()
    print("%s API: forward %.1f ms, sweep %.1f ms" % (dt, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
