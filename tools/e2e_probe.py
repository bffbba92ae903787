"""Where does the end-to-end time of Network.forward(X_host) + Network.sweep go?  (config 3 shapes, reduced S)"""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensornetworkforml_b200 as tn
S, D, L, Ns = 196, 64, 10, 60000
np.random.seed(0)
X = np.random.random((Ns, S)); X = np.ascontiguousarray(np.stack((np.sin(np.pi * X / 2), np.cos(np.pi * X / 2)), -1))
y = np.random.randint(0, L, Ns)
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:2048], act_fn="linear", loss_fn="MSE",
                     truncation="fixed", max_bond=D)
eng = net._engine()
def t(label, fn, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); print("%-34s %8.2f ms" % (label, (time.perf_counter() - t0) / n * 1e3)); return r
t("load_input (H2D + pack) first", lambda: eng.load_input(X), 1)
from tensornetworkforml_b200 import _lib
print("registered:", _lib._registered[:1], "pinned staging:", None if eng._pinned is None else eng._pinned.numel())
t("load_input (H2D + pack)", lambda: eng.load_input(X))
t("engine.forward", lambda: eng.forward())
t("Network.forward(X)", lambda: net.forward(X))
for sweep in range(3):
    f = net.forward(X)
    left = net.l_pos == S - 1
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.begin_sweep(y, left, True)
    t1 = time.perf_counter()
    for _ in range(S - 1): eng.sweep_step(1e-4, 1e-3, True, left)
    t2 = time.perf_counter(); torch.cuda.synchronize(); t3 = time.perf_counter()
    net.l_pos = eng.l_pos; net._host_fresh = False
    h = eng.history(); t4 = time.perf_counter()
    print("sweep %d: begin %.1f ms, python enqueue of 195 steps %.1f ms, wait for GPU %.1f ms, history %.1f ms" %
          (sweep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
print("--- through the public API")
for sweep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f = net.forward(X); t1 = time.perf_counter()
    left = net.l_pos == S - 1
    f = net.sweep(X, y, f, 1e-4, 1e-3, L2_flag=True, left_dir=left); t2 = time.perf_counter()
    print("api sweep %d: forward %.1f ms, sweep %.1f ms" % (sweep, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
f = net.forward(X); f = net.sweep(X, y, f, 1e-4, 1e-3, L2_flag=True, left_dir=(net.l_pos == S - 1))
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
