"""Pure host cost of enqueueing one sweep: the C-ABI calls are replaced by no-ops, so nothing blocks on the GPU."""
import contextlib, io, os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tensornetworkforml_b200 as tn
from tensornetworkforml_b200 import engine as E
c = bench.CFG
S, L, D, Ns = c["S"], c["L"], c["D"], c["Ns"]
X, y = bench.synthetic_data(Ns, S, L, c["seed"])
np.random.seed(c["seed"])
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:2048], act_fn=c["act"], loss_fn=c["loss"],
                     truncation="fixed", max_bond=D)
eng = net._engine()
eng.load_input(X)
yd = torch.from_numpy(y.astype(np.int32)).to(eng.device)
def sweep():
    eng.forward(); left = eng.l_pos == S - 1
    eng.begin_sweep(yd, left, True)
    for _ in range(S - 1): eng.sweep_step(c["lr"], c["wd"], True, left)
for _ in range(3): sweep()
torch.cuda.synchronize()
real_call = E.call
ncalls = [0]
def fake(name, *a): ncalls[0] += 1
E.call = fake
t0 = time.perf_counter(); sweep(); t1 = time.perf_counter()
print("host-only enqueue of one sweep (no-op C calls): %.1f ms for %d C calls" % ((t1 - t0) * 1e3, ncalls[0]))
lib = E._lib.lib()
def cheap(name, *a): ncalls[0] += 1; lib.tnml_version()
E.call = cheap
t0 = time.perf_counter(); sweep(); t1 = time.perf_counter()
print("with a trivial ctypes call each: %.1f ms" % ((t1 - t0) * 1e3))
E.call = fake
pr = cProfile.Profile(); pr.enable(); sweep(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
