"""Dump a few updated bond tensors B' of the bench workload (config 3) to gpurun_out/bond_dumps.npz, for offline
study of the Jacobi sweep counts (tools/jacobi_study.py)."""
import contextlib, io, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import tensornetworkforml_b200 as tn

c = dict(bench.CFG, Ns=int(os.environ.get("PNS", 60000)))
S, L, D, Ns = c["S"], c["L"], c["D"], c["Ns"]
X, y = bench.synthetic_data(Ns, S, L, c["seed"])
np.random.seed(c["seed"])
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=L, normalize=True, calibration_X=X[:2048], act_fn=c["act"], loss_fn=c["loss"],
                     truncation="fixed", max_bond=D, device="cuda:0")
eng = net._engine()
eng.load_input(X)
y_dev = torch.from_numpy(y.astype(np.int32)).to(eng.device)
want = {(0, 60), (0, 150), (1, 100), (3, 30), (3, 100), (3, 170), (5, 100)}
out = {}
for sweep in range(6):
    eng.forward()
    left = eng.l_pos == S - 1
    eng.begin_sweep(y_dev, left, c["L2"])
    for step in range(S - 1):
        ctx = eng.update_phase(c["lr"], c["wd"], c["L2"], left)
        if (sweep, step) in want:
            torch.cuda.synchronize()
            R, Cc = (2 * ctx["Dl"], 2 * L * ctx["Dr"]) if not left else (2 * ctx["Dl"] * L, 2 * ctx["Dr"])
            out["B_%d_%d_%d" % (sweep, step, int(left))] = ctx["Bn"].cpu().numpy().reshape(R, Cc)
        eng.split_phase(ctx)
    h = eng.history()
    sv = eng.hist["svals"][:eng.hist["n"]].cpu().numpy()
    nsv = eng.hist["nsv"]
    print("sweep", sweep, "pass1/pass2 sweeps at steps 30,100,170:",
          [(int(sv[i, nsv[i]]), int(sv[i, nsv[i] + 1])) for i in (30, 100, 170)], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "bond_dumps.npz"), **out)
print("saved", sorted(out))
