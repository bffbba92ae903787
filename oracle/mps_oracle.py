"""CPU oracle for the sweeping MPS-classifier hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm in
``/root/reference/TensorNetwork/Network_class.py`` (``NC``), ``custom_linalg_tools.py`` (``CLT``),
``Tensor_class.py`` (``TC``) and ``data_generator.py`` (``DG``).  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product package ``tensornetworkforml_b200`` never does.

Pinning: the reference ships no unit tests or golden vectors for this path, so the oracle is
pinned by (a) live re-execution of the reference imported from ``/root/reference`` (``tests/test_oracle_vs_reference.py``,
run in the build container) and (b) fixtures produced by the reference itself with
``tests/golden/make_golden.py`` and committed under ``tests/golden/`` (these travel to the GPU box).

Layout used here (the "canonical" layout of DESIGN.md):
  non-label site  A[p] : (Dl, 2, Dr)       axes (left bond a, physical sigma, right bond c)
  label site      A[p] : (Dl, 2, L, Dr)    axes (a, sigma, label l, c)
  bond tensor     B    : (Dl, 2, L, 2, Dr) axes (a, sigma, l, tau, c)
  environments    E    : (Ns, D)           per-sample rows; left env of sites < p, right env of sites >= p
  outputs         f    : (Ns, L)           (the reference holds (L, Ns); transpose on export)
Sites 0 and S-1 carry a dummy bond of size 1, so no edge special-casing is needed in the arithmetic.
"""
from __future__ import annotations

import numpy as np

ACT_FNS = ("linear", "sigmoid", "softmax")            # NC:127
EXT_ACT_FNS = ("softmax_stable",)                      # opt-in extension, not in the reference
LOSS_FNS = ("MSE", "cross_entropy", "full_cross_ent")  # NC:132


# --------------------------------------------------------------------------------------------------
# a1: feature map                                                                      DG:165-167
# --------------------------------------------------------------------------------------------------
def feature_map(x: np.ndarray) -> np.ndarray:
    """phi(x) = [sin(pi x / 2), cos(pi x / 2)] -- sin FIRST, as in the code (DG:165-167, NC:152-155)."""
    x = np.asarray(x, dtype=np.float64)
    return np.stack((np.sin(np.pi * x / 2), np.cos(np.pi * x / 2)), axis=-1)


# --------------------------------------------------------------------------------------------------
# reference <-> canonical layout                                         NC:145-148, TC:97-199
# --------------------------------------------------------------------------------------------------
def site_from_named(elem: np.ndarray, axes_names, p: int) -> np.ndarray:
    """Bring a reference site tensor (any axis order, named axes) to canonical layout.

    Names used by the reference: 'left', 'right', 'd<p>', 'l' (NC:145-148, NC:918-925)."""
    names = [str(n) for n in axes_names]
    order, shape_extra = [], []
    want = ["left", "d%d" % p, "l", "right"]
    arr = np.asarray(elem, dtype=np.float64)
    present = [w for w in want if w in names]
    arr = np.transpose(arr, [names.index(w) for w in present])
    # insert singleton bonds where the edge sites have none
    if "left" not in names:
        arr = arr[np.newaxis, ...]
    if "right" not in names:
        arr = arr[..., np.newaxis]
    return np.ascontiguousarray(arr)


def sites_from_reference(As) -> list:
    """list[reference Tensor] -> list of canonical arrays."""
    return [site_from_named(T.elem, T.axes_names, p) for p, T in enumerate(As)]


def draw_initial_sites(S: int, M: int, L: int, d: int = 2, scale: float = 1.0) -> list:
    """Draw the initial weights with the reference's RNG calls, in the reference's order and shapes
    (TC:63-64 ``np.random.random(size=shape) / scale``; NC:145-148 / NC:186-189), then re-lay them out."""
    raw0 = np.random.random(size=[L, M, d]) / scale           # axes l, right, d0
    sites = [np.ascontiguousarray(np.transpose(raw0, (2, 0, 1))[np.newaxis])]   # (1, d, L, M)
    for _ in range(1, S - 1):
        raw = np.random.random(size=[M, M, d]) / scale        # axes left, right, d
        sites.append(np.ascontiguousarray(np.transpose(raw, (0, 2, 1))))
    rawN = np.random.random(size=[M, d]) / scale               # axes left, d
    sites.append(np.ascontiguousarray(rawN[..., np.newaxis]))
    return sites


# --------------------------------------------------------------------------------------------------
# a5 / a10: environments                                       NC:227-255, NC:637-642, NC:669-674
# --------------------------------------------------------------------------------------------------
def env_advance_right(E: np.ndarray, phi_p: np.ndarray, A: np.ndarray) -> np.ndarray:
    """L_{p+1}[b,m] = sum_{a,s} L_p[b,a] phi_p[b,s] A_p[a,s,m]        (NC:637-642, NC:244-251)."""
    Dl, d, Dr = A.shape
    U = (E[:, :, None] * phi_p[:, None, :]).reshape(E.shape[0], Dl * d)
    return U @ A.reshape(Dl * d, Dr)


def env_advance_left(E: np.ndarray, phi_p: np.ndarray, A: np.ndarray) -> np.ndarray:
    """R_p[b,a] = sum_{s,c} A_p[a,s,c] phi_p[b,s] R_{p+1}[b,c]          (NC:669-674, NC:234-238)."""
    Dl, d, Dr = A.shape
    V = (phi_p[:, :, None] * E[:, None, :]).reshape(E.shape[0], d * Dr)
    return V @ A.reshape(Dl, d * Dr).T


def site_predict(Lenv, phi_p, A_label, Renv) -> np.ndarray:
    """f[b,l] = sum L[b,a] phi[b,s] A[a,s,l,c] R[b,c]  -- last contraction of forward (NC:242, NC:255)."""
    return np.einsum("ba,bs,aslc,bc->bl", Lenv, phi_p, A_label, Renv, optimize=True)


# --------------------------------------------------------------------------------------------------
# a11 / a12: activation and loss derivative                                  NC:767-796, NC:800-835
# --------------------------------------------------------------------------------------------------
def apply_act(f: np.ndarray, act_fn: str, T: float) -> np.ndarray:
    """f is (Ns, L).  Softmax is NOT max-stabilised, exactly as NC:794."""
    if act_fn == "linear":
        return f.copy()
    if act_fn == "sigmoid":
        return 1.0 / (1.0 + np.exp(-f / T))                               # NC:791
    if act_fn == "softmax":
        e = np.exp(f / T)
        return e / e.sum(axis=1, keepdims=True)                            # NC:794
    if act_fn == "softmax_stable":      # opt-in: NC:794 with the largest logit subtracted (same value, no overflow)
        e = np.exp((f - f.max(axis=1, keepdims=True)) / T)
        return e / e.sum(axis=1, keepdims=True)
    raise AssertionError(act_fn)


def loss_derivative(fa: np.ndarray, y1h: np.ndarray, act_fn: str, loss_fn: str, T: float) -> np.ndarray:
    """fa = activated output (Ns, L); y1h = one-hot (Ns, L)."""
    if loss_fn == "MSE":
        return y1h - fa                                                    # NC:824
    if loss_fn == "cross_entropy":
        if act_fn in ("softmax", "softmax_stable"):
            return (y1h - y1h * fa) / T                                    # NC:828
        return y1h / fa                                                    # NC:830
    if loss_fn == "full_cross_ent":
        g = fa.copy()
        g[y1h == 0] -= 1.0                                                 # NC:832
        return 1.0 / (g + 1e-4)                                            # NC:833
    raise AssertionError(loss_fn)


def metrics(fa: np.ndarray, y1h: np.ndarray):
    """accuracy and MAE exactly as NC:697-702 (pre-update values)."""
    y_pred = np.argmax(fa, axis=1)
    y_tgt = np.argmax(y1h, axis=1)
    acc = (len(y_pred) - (y_tgt != y_pred).sum()) / len(y_pred)
    mae = np.abs(y1h - fa).mean()
    return float(acc), float(mae)


# --------------------------------------------------------------------------------------------------
# a10: gradient, a9: projection                                    NC:625-646, NC:710, NC:494-523
# --------------------------------------------------------------------------------------------------
def _uv(Lenv, phi_p, phi_q, Renv):
    Ns = Lenv.shape[0]
    U = (Lenv[:, :, None] * phi_p[:, None, :]).reshape(Ns, -1)          # (Ns, a*2)
    V = (phi_q[:, :, None] * Renv[:, None, :]).reshape(Ns, -1)          # (Ns, 2*c)
    return U, V


def gradient(g, Lenv, phi_p, phi_q, Renv) -> np.ndarray:
    """dB[a,s,l,t,c] = sum_b g[b,l] L[b,a] phi_p[b,s] phi_q[b,t] R[b,c]      (NC:625-646, NC:710)."""
    Ns, a = Lenv.shape
    c = Renv.shape[1]
    L = g.shape[1]
    U, V = _uv(Lenv, phi_p, phi_q, Renv)
    W = (g[:, :, None] * V[:, None, :]).reshape(Ns, L * 2 * c)
    return (U.T @ W).reshape(a, 2, L, 2, c)


def project(B, Lenv, phi_p, phi_q, Renv) -> np.ndarray:
    """f[b,l] = sum B[a,s,l,t,c] L[b,a] phi_p[b,s] phi_q[b,t] R[b,c]          (NC:494-523)."""
    a, _, L, _, c = B.shape
    U, V = _uv(Lenv, phi_p, phi_q, Renv)
    Tm = (U @ B.reshape(a * 2, L * 2 * c)).reshape(-1, L, 2 * c)
    return np.einsum("blk,bk->bl", Tm, V)


# --------------------------------------------------------------------------------------------------
# a14: L2 term through norm environments                                           NC:966-1179
# --------------------------------------------------------------------------------------------------
def norm_env_right_step(E: np.ndarray, A: np.ndarray) -> np.ndarray:
    """E_L^{(p+1)}[m,m'] = sum_{a,a',s} E_L^{(p)}[a,a'] A[a,s,m] A[a',s,m']        (NC:1004-1029)."""
    return np.einsum("ax,asm,xsn->mn", E, A, A, optimize=True)


def norm_env_left_step(E: np.ndarray, A: np.ndarray) -> np.ndarray:
    """E_R^{(p)}[a,a'] = sum_{c,c',s} A[a,s,c] A[a',s,c'] E_R^{(p+1)}[c,c']          (NC:1035-1061)."""
    return np.einsum("asc,xsd,cd->ax", A, A, E, optimize=True)


def l2_term(B, EL, ER, wd):
    """derivate = E_L . B . E_R ; loss = wd <B, derivate> ; grad = 2 wd derivate   (NC:1129-1177)."""
    G = np.einsum("xa,asltc,cy->xslty", EL, B, ER, optimize=True)
    return float(wd * np.sum(B * G)), 2.0 * wd * G


def clip_and_update(B, dB, lr):
    """NC:755-761: clip on sum|dB| vs sum|B| (division by the ratio, as the reference), then B + lr dB."""
    bm = np.abs(B).sum()
    sd = np.abs(dB).sum()
    if sd > bm:
        dB = dB / (sd / bm)
    dB = dB * lr
    return B + dB


# --------------------------------------------------------------------------------------------------
# a13: SVD split                                                          NC:528-563, NC:839-962
# --------------------------------------------------------------------------------------------------
def choose_m(rule: str, left_dir: bool, pos_l: int, S_sites: int, Dl: int, nS: int, R: int, C: int, max_bond):
    """Bond kept after the split.  'reference' = rule of NC:898-910 / NC:933-945 (copies the left bond, or
    len(S) at the chain ends; raises where the reference's np.dot would).  'fixed' = min(len(S), max_bond)."""
    if rule == "fixed":
        return min(nS, int(max_bond))
    if rule != "reference":
        raise ValueError("unknown truncation rule %r" % (rule,))
    if not left_dir:
        if pos_l == 0:
            return nS                                   # NC:898-901 (only Vh cut; U is R x R with R = nS)
        if pos_l < S_sites - 2:
            m = Dl                                      # NC:902-906
            if m > nS:
                raise ValueError("reference rule: left bond %d exceeds len(S)=%d" % (m, nS))
            return m
        if C != nS:                                     # NC:907-910 -> np.dot(Sqrt, Vh) misaligned
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (nS, nS, C, C))
        return nS
    else:
        if pos_l == S_sites - 1:
            if C != nS:                                 # NC:933-936: Vh uncut
                raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (nS, nS, C, C))
            return nS
        if pos_l > 1:
            m = Dl                                      # NC:937-941
            if m > nS:
                raise ValueError("reference rule: left bond %d exceeds len(S)=%d" % (m, nS))
            return m
        if R != nS:                                     # NC:942-945: U uncut -> np.dot(U, Sqrt) misaligned
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (R, R, nS, nS))
        return nS


def adaptive_m(S: np.ndarray, threshold: float, min_bond: int, max_bond: int) -> int:
    """Opt-in adaptive bond (the authors' unfinished feature): NC:890-891 computes
    ``index = argmax(cumsum(S)/S.sum() > threshold)`` and never uses it; the intent is in the reference's
    old_files/TensorNetwork.py:1310-1326, ``m_new = max(10, min(index, m))``.  Restated with the two constants as
    parameters: m = max(min_bond, min(index, max_bond)), never more than len(S)."""
    cve = np.cumsum(S) / S.sum()
    index = int(np.argmax(cve > threshold))
    return int(min(len(S), max(int(min_bond), min(index, int(max_bond)))))


def svd_split(B, left_dir: bool, m: int):
    """Split B[a,s,l,t,c] into the two new sites with sqrt(S) on both factors (NC:887, NC:912-925, NC:947-960).

    right sweep: rows (a,s)   cols (l,t,c) -> left site (a,s,m) plain, right site (m,t,l,c) carries the label
    left  sweep: rows (a,s,l) cols (t,c)   -> left site (a,s,l,m) carries the label, right site (m,t,c) plain
    Row/column ORDER inside a group does not change the split (the SVD is permutation-covariant)."""
    a, _, L, _, c = B.shape
    if not left_dir:
        Mx = B.reshape(a * 2, L * 2 * c)
    else:
        Mx = B.reshape(a * 2 * L, 2 * c)
    U, S, Vh = np.linalg.svd(Mx, full_matrices=False)                      # NC:887 (economy = same leading part)
    sq = np.sqrt(S[:m])
    US = U[:, :m] * sq[None, :]
    SVh = sq[:, None] * Vh[:m, :]
    if not left_dir:
        A_left = US.reshape(a, 2, m)
        A_right = np.ascontiguousarray(np.transpose(SVh.reshape(m, L, 2, c), (0, 2, 1, 3)))   # (m,t,l,c)
    else:
        A_left = US.reshape(a, 2, L, m)
        A_right = SVh.reshape(m, 2, c)
    return A_left, A_right, S


# --------------------------------------------------------------------------------------------------
# The network: forward / sweep / sweep_step                      NC:195-258, NC:384-436, NC:440-573
# --------------------------------------------------------------------------------------------------
class OracleMPS:
    """State machine mirroring ``Network`` (NC:10) on canonical arrays.  Records per-step history."""

    def __init__(self, sites, L, T=0.1, act_fn="linear", loss_fn="cross_entropy",
                 rule="reference", max_bond=None, l_pos=0, threshold=0.999, min_bond=2):
        assert act_fn in ACT_FNS + EXT_ACT_FNS and loss_fn in LOSS_FNS
        self.sites = [np.array(s, dtype=np.float64) for s in sites]
        self.S = len(sites)
        self.L = L
        self.T = T
        self.act_fn, self.loss_fn = act_fn, loss_fn
        self.rule, self.max_bond = rule, max_bond
        self.threshold, self.min_bond = threshold, min_bond
        self.l_pos = l_pos
        self.phi = None
        self.env = None          # env[p]: left env of sites < p (valid p <= l_pos side) or right env of sites >= p
        self.hist = []           # one dict per bond update

    # -- construction the way the reference does it (NC:137-189) --------------------------------
    @classmethod
    def from_seed(cls, S, M, L, calibration_X=None, normalize=False, **kw):
        d = 2
        scale = float(M) * 0.5 * 0.64 * d if normalize else 1.0             # NC:142
        sites = draw_initial_sites(S, M, L, d, scale)
        net = cls(sites, L, **kw)
        if normalize:
            if calibration_X is None:
                calibration_X = feature_map(np.random.random((16, S)))       # NC:157-159
            f = net.forward(calibration_X)
            f_max = np.abs(f).max()
            F2 = f_max ** (1.0 / S)                                          # NC:170
            net.sites = [A / F2 for A in net.sites]                          # NC:175-176
            net.calibration_factor = F2
            net.forward(calibration_X)                                       # NC:179 (leaves caches like the reference)
        return net

    # -- forward (NC:195-258) ----------------------------------------------------------------------
    def forward(self, X):
        X = np.asarray(X, dtype=np.float64)
        assert X.shape[1] == self.S
        Ns, S = X.shape[0], self.S
        self.phi = X
        env = [None] * (S + 1)
        if self.l_pos == 0:
            env[S] = np.ones((Ns, 1))
            for p in range(S - 1, 0, -1):
                env[p] = env_advance_left(env[p + 1], X[:, p, :], self.sites[p])
            env[0] = np.ones((Ns, 1))
            f = site_predict(env[0], X[:, 0, :], self.sites[0], env[1])
        elif self.l_pos == S - 1:
            env[0] = np.ones((Ns, 1))
            for p in range(0, S - 1):
                env[p + 1] = env_advance_right(env[p], X[:, p, :], self.sites[p])
            env[S] = np.ones((Ns, 1))
            f = site_predict(env[S - 1], X[:, S - 1, :], self.sites[S - 1], env[S])
        else:
            raise Exception("forward should not be called if l has an intermediate position")   # NC:258
        self.env = env
        return f

    def accuracy(self, X, y, f=None):                                        # NC:354-380
        if f is None:
            f = self.forward(X)
        y_pred = np.argmax(f, axis=1)
        return (len(y_pred) - (y != y_pred).sum()) / len(y_pred)

    # -- sweep (NC:384-436) ------------------------------------------------------------------------
    def sweep(self, y, f, lr, weight_dec, L2_flag=True, left_dir=False):
        y1h = np.zeros((y.size, self.L))
        y1h[np.arange(y.size), y] = 1                                        # NC:421-423
        # norm environments on the far side of the sweep are fixed during it; build the stack once
        self._norm = None
        if L2_flag:
            self._build_norm_stack(left_dir)
        for _ in range(self.S - 1):
            f = self.sweep_step(f, y1h, lr, weight_dec, L2_flag, left_dir)
        return f

    def _build_norm_stack(self, left_dir):
        S = self.S
        nrm = [None] * (S + 1)
        if not left_dir:
            nrm[S] = np.ones((1, 1))
            for p in range(S - 1, 1, -1):
                nrm[p] = norm_env_left_step(nrm[p + 1], self.sites[p])
            nrm[0] = np.ones((1, 1))
        else:
            nrm[0] = np.ones((1, 1))
            for p in range(0, S - 2):
                nrm[p + 1] = norm_env_right_step(nrm[p], self.sites[p])
            nrm[S] = np.ones((1, 1))
        self._norm = nrm

    # -- one bond update (NC:440-573 with update_B NC:577-763 inlined) -------------------------------
    def sweep_step(self, f, y1h, lr, weight_dec, L2_flag, left_dir):
        S, l = self.S, self.l_pos
        p = l - 1 if left_dir else l                 # left site of the pair (p, p+1)
        q = p + 1
        X, env = self.phi, self.env
        # environment advance with the site fixed by the previous step (NC:637-642 / NC:669-674)
        if not left_dir and p > 0:
            env[p] = env_advance_right(env[p - 1], X[:, p - 1, :], self.sites[p - 1])
            if L2_flag:
                self._norm[p] = norm_env_right_step(self._norm[p - 1], self.sites[p - 1])
        if left_dir and q < S - 1:
            env[q + 1] = env_advance_left(env[q + 2], X[:, q + 1, :], self.sites[q + 1])
            if L2_flag:
                self._norm[q + 1] = norm_env_left_step(self._norm[q + 2], self.sites[q + 1])
        Lenv, Renv = env[p], env[q + 1]
        # B = A_p . A_q                                                        NC:484
        if not left_dir:
            B = np.einsum("aslm,mtc->asltc", self.sites[p], self.sites[q])
        else:
            B = np.einsum("asm,mtlc->asltc", self.sites[p], self.sites[q])
        fa = apply_act(f, self.act_fn, self.T)                                # NC:694
        acc, mae = metrics(fa, y1h)                                           # NC:697-702
        g = loss_derivative(fa, y1h, self.act_fn, self.loss_fn, self.T)       # NC:707
        dB = gradient(g, Lenv, X[:, p, :], X[:, q, :], Renv)                   # NC:710
        l2_loss = None
        if L2_flag:                                                           # NC:728-734
            l2_loss, l2_grad = l2_term(B, self._norm[p], self._norm[q + 1], weight_dec)
            dB = dB - l2_grad
        else:
            l2_grad = weight_dec * B                                          # NC:731-734
            dB = dB - l2_grad
        absreg = float(np.abs(l2_grad).mean())                                # NC:747 (debug history)
        absB, absdB = float(np.abs(B).mean()), float(np.abs(dB).mean())       # NC:741-742 (debug history)
        Bn = clip_and_update(B, dB, lr)                                       # NC:755-761
        f_new = project(Bn, Lenv, X[:, p, :], X[:, q, :], Renv)                # NC:494-523 (UN-truncated B')
        a, c = B.shape[0], B.shape[4]
        R_, C_ = (a * 2, self.L * 2 * c) if not left_dir else (a * 2 * self.L, 2 * c)
        nS = min(R_, C_)
        if self.rule == "adaptive":
            Mx = Bn.reshape(R_, C_)
            m = adaptive_m(np.linalg.svd(Mx, compute_uv=False), self.threshold, self.min_bond, self.max_bond)
        else:
            m = choose_m(self.rule, left_dir, l, S, a, nS, R_, C_, self.max_bond)
        A_left, A_right, Svals = svd_split(Bn, left_dir, m)                   # NC:563
        if not left_dir:
            self.sites[p] = A_left                                            # (a,s,m)
            self.sites[q] = A_right                                           # (m,t,l,c) label site
            self.l_pos += 1                                                   # NC:568-569
        else:
            self.sites[p] = A_left                                            # (a,s,l,m) label site
            self.sites[q] = A_right
            self.l_pos -= 1                                                   # NC:570-571
        self.hist.append(dict(acc=acc, mae=mae, S=Svals, m=m, l2_loss=l2_loss, absB=absB, absdB=absdB, absreg=absreg,
                              absf=float(np.abs(f).mean())))
        return f_new

    def bond_dims(self):
        return [A.shape[-1] for A in self.sites[:-1]]
