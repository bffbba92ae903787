"""Build container only: the VERBATIM reference (imported from /root/reference, only tensor_svd's m replaced by the
fixed-D rule of SURVEY.md section 8c) timed on interior bond updates with config-3 bond dimensions (D = 64, L = 10) at
the largest Ns that is practical here, with and without the L2 term; linear-in-Ns extrapolation to Ns = 60 000 for
BASELINE.md.  (The reference cannot run config 3 itself: 78 GB intermediate in update_B)"""
import contextlib, io, os, sys, time, copy
import numpy as np
REF = "/root/reference/TensorNetwork"
sys.path.insert(0, REF)
import Network_class as NC
from Tensor_class import Tensor

D, L, S = 64, 10, int(os.environ.get("S", 16))


class FixedD(NC.Network):
    Dmax = D

    def tensor_svd(self, T, left_dir=False, threshold=0.999):
        U, Sv, Vh = np.linalg.svd(copy.deepcopy(T.elem))
        m = min(len(Sv), self.Dmax)
        U, Vh = U[:, :m], Vh[:m, :]
        Sq = np.sqrt(np.eye(m, m) * Sv[:m])
        TU = Tensor(elem=np.dot(U, Sq), axes_names=['i', 'right'])
        TSVh = Tensor(elem=np.dot(Sq, Vh), axes_names=['left', 'j'])
        TU.aggregations['i'] = T.aggregations['i']
        TSVh.aggregations['j'] = T.aggregations['j']
        TU.disaggregate('i')
        TSVh.disaggregate('j')
        return TU, TSVh


for Ns in [int(v) for v in os.environ.get("NS", "64,128").split(",")]:
    for L2 in (False, True):
        np.random.seed(0)
        x = np.random.random((Ns, S))
        X = np.stack((np.sin(np.pi * x / 2), np.cos(np.pi * x / 2)), -1)
        y = np.random.randint(0, L, Ns)
        with contextlib.redirect_stdout(io.StringIO()):
            net = FixedD(N=S, M=D, L=L, normalize=True, calibration_X=X, act_fn="linear", loss_fn="MSE")
            t0 = time.perf_counter()
            f = net.forward(X)
            t_fwd = time.perf_counter() - t0
            yy = np.zeros((L, Ns)); yy[y, np.arange(Ns)] = 1
            net.l_cum_contraction = []
            times = []
            for i in range(S - 1):
                t0 = time.perf_counter()
                f = net.sweep_step(f, yy, 1e-4, Ns, 1e-3, L2_flag=L2, left_dir=False)
                times.append(time.perf_counter() - t0)
        bonds = [a.elem.shape for a in net.As]
        interior = times[6:S - 4]
        print("S=%d Ns=%d L2=%s forward %.2fs; per bond update (interior, D=64 both sides): %s -> mean %.3f s"
              % (S, Ns, L2, t_fwd, ["%.2f" % t for t in interior], float(np.mean(interior))), flush=True)
