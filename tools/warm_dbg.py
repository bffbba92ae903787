"""GPU probe: which interior splits of a small chain take the warm-started path, and why the others are refused
(code / deciding ratio left by k_fast_split), plus the spectral gap at the truncation point."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensornetworkforml_b200 as tn
import tensornetworkforml_b200.data_generator as gen
side, D, Lbl, Ns = int(os.environ.get("SIDE", 4)), 64, 10, int(os.environ.get("NS", 2048))
lr, wd = float(os.environ.get("LR", 1e-4)), 1e-3
S = side * side
np.random.seed(2)
data, labels = gen.create_multiclass_dataset(Ns, side, Lbl, 0.7)
X = gen.psi(data.reshape(Ns, -1)); y = labels.astype(np.int64)
np.random.seed(2)
with contextlib.redirect_stdout(io.StringIO()):
    net = tn.Network(N=S, M=D, L=Lbl, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE", truncation="fixed", max_bond=D)
for sw in range(int(os.environ.get("NSW", 6))):
    f = net.forward(X); f = net.sweep(X, y, f, lr, wd, L2_flag=True, left_dir=(net.l_pos == S - 1))
    eng = net._eng; sv = eng.hist["svals"][:eng.hist["n"]].cpu().numpy()
    print(sw, [(int(sv[i, n]), sv[i, n + 2], float("%.3g" % sv[i, n + 3]), float("%.2g" % (sv[i, n // 2] / sv[i, n // 2 - 1])))
               for i, n in enumerate(eng.hist["nsv"]) if n == 128])
