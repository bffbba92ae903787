"""How long does the deferred tail refinement take on its stream, and does it lag behind the sweep?"""
import contextlib, io, os, sys, time
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
# wrap the tail call with events
orig_call = E.call
rec = []
def call(name, *a):
    if name == "tnml_svd_split_tail":
        tail = eng._tail_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(tail); orig_call(name, *a); e1.record(tail); rec.append((e0, e1))
    else:
        orig_call(name, *a)
E.call = call
for sw in range(int(os.environ.get("NSW", 10))):
    rec.clear()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s0 = torch.cuda.Event(enable_timing=True); s0.record()
    eng.forward(); left = eng.l_pos == S - 1
    eng.begin_sweep(yd, left, True)
    for _ in range(S - 1): eng.sweep_step(c["lr"], c["wd"], True, left)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    dur = np.array([a.elapsed_time(b) for a, b in rec])
    end = np.array([s0.elapsed_time(b) for a, b in rec])
    ms = torch.cuda.memory_stats()
    print("sweep %2d %.1f ms | tail: mean %.3f ms max %.3f ms, sum %.1f ms, last tail ends at %.1f ms | device allocs %d, reserved %.2f GB" % (
        sw, t * 1e3, dur.mean(), dur.max(), dur.sum(), end[-1], ms.get("num_device_alloc", -1), ms["reserved_bytes.all.current"] / 1e9), flush=True)
