"""Per-sweep time and Jacobi iteration counts over many sweeps of config 3 (does the SVD split slow down as training goes on?)."""
import contextlib, io, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import tensornetworkforml_b200 as tn
c = bench.CFG
S, L, D, Ns = c["S"], c["L"], c["D"], c["Ns"]
X, y = bench.synthetic_data(Ns, S, L, c["seed"])
np.random.seed(c["seed"])
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:2048], act_fn=c["act"], loss_fn=c["loss"],
                     truncation="fixed", max_bond=D, dtype=os.environ.get("PDT", "float64"))
eng = net._engine()
eng.load_input(X)
yd = torch.from_numpy(y.astype(np.int32)).to(eng.device)
for sw in range(14):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eng.forward(); left = eng.l_pos == S - 1
    eng.begin_sweep(yd, left, True)
    for _ in range(S - 1): eng.sweep_step(c["lr"], c["wd"], True, left)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    h = eng.history()
    sv = eng.hist["svals"][:eng.hist["n"]].cpu().numpy(); nsv = eng.hist["nsv"]
    it = np.array([[sv[i, nsv[i]], sv[i, nsv[i] + 1]] for i in range(len(nsv)) if nsv[i] == 2 * D])
    gap = np.array([h["svals"][i][D] / h["svals"][i][D - 1] for i in range(len(nsv)) if nsv[i] == 2 * D])
    spread = np.array([h["svals"][i][D - 1] / h["svals"][i][0] for i in range(len(nsv)) if nsv[i] == 2 * D])
    print("sweep %2d %s %.1f ms | pass1 iters mean %.1f min %d max %d | pass2 mean %.1f | gap s65/s64 median %.1e max %.1e | s64/s1 median %.4f | acc %.3f" % (
        sw, "L" if left else "R", t * 1e3, it[:, 0].mean(), it[:, 0].min(), it[:, 0].max(), it[:, 1].mean(), np.median(gap), gap.max(), np.median(spread), h["acc"][-1]), flush=True)
