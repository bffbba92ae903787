"""Probe: a short chain with config-3 bond/label dimensions (the batch-independent kernels do not depend on S or
Ns), printing the Jacobi sweep counters.  Used under ncu for the per-launch time list."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensornetworkforml_b200 as tn

S, D, L, Ns = int(os.environ.get("PS", 20)), int(os.environ.get("PD", 64)), 10, int(os.environ.get("PNS", 8192))
np.random.seed(0)
X = np.random.random((Ns, S)); X = np.stack((np.sin(np.pi * X / 2), np.cos(np.pi * X / 2)), -1)
y = np.random.randint(0, L, Ns)
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE",
                     truncation="fixed", max_bond=D, dtype=os.environ.get("PDT", "float64"))
for sw in range(int(os.environ.get("PSW", 3))):
    f = net.forward(X)
    f = net.sweep(X, y, f, 1e-4, 1e-3, left_dir=(net.l_pos == S - 1))
    eng = net._eng
    sv = eng.hist["svals"].cpu().numpy()
    info = [(eng.hist["nsv"][i], int(sv[i, eng.hist["nsv"][i]]), int(sv[i, eng.hist["nsv"][i] + 1])) for i in range(S - 1)]
    print("sweep", sw, "bonds", eng.bond_dims(), "\n (n, sweeps pass1, sweeps pass2):", info)
torch.cuda.synchronize()
