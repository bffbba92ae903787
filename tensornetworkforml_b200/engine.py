"""Device-resident state and kernel sequencing for the MPS sweep (host side of the C ABI).

PyTorch is used for plumbing only: device memory, streams, ``torch.distributed``.  Every arithmetic step of the
hot path is a call into libtnml.so (``_lib.call``); there is no CPU or eager-PyTorch fallback.

Index conventions:
  bonds[p]  = dimension of the bond between sites p-1 and p   (bonds[0] = bonds[S] = 1)
  env[p]    = (Ns, bonds[p]) per-sample environment living on that bond: the LEFT environment of sites < p while
              the label is at or right of p, the RIGHT environment of sites >= p otherwise.  One arena holds both
              families because at any time each bond needs only one of them.
  nrmL[p] / nrmR[p] = (bonds[p], bonds[p]) norm environments of sites < p / >= p   (NC:966-1179)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from ._lib import F64, F32, ACT, LOSS, call
from .parallel import reduce_gradient_and_metrics


def _ptr(t):
    return t.data_ptr() if t is not None else None


def choose_m(rule, max_bond, left_dir, l_pos, S, Dl, R, C):
    """Bond kept by the SVD split.  'fixed': min(len(S), max_bond).  'reference': NC:898-910 / NC:933-945 -- the left
    bond of the pair in the interior, len(S) at the chain ends; raises where the reference's np.dot would."""
    nS = min(R, C)
    if rule in ("fixed", "adaptive"):          # adaptive: the cap; the data-dependent cut follows the split
        return min(nS, max_bond)
    if not left_dir:
        if l_pos == 0:
            return nS                                    # NC:898-901
        if l_pos < S - 2:
            if Dl > nS:
                raise ValueError("reference rule: left bond %d exceeds len(S)=%d" % (Dl, nS))
            return Dl                                    # NC:902-906
        if C != nS:                                      # NC:907-910: Vh stays (C,C), np.dot(Sqrt, Vh) fails
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (nS, nS, C, C))
        return nS
    if l_pos == S - 1:
        if C != nS:                                      # NC:933-936
            raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (nS, nS, C, C))
        return nS
    if l_pos > 1:
        if Dl > nS:
            raise ValueError("reference rule: left bond %d exceeds len(S)=%d" % (Dl, nS))
        return Dl                                        # NC:937-941
    if R != nS:                                          # NC:942-945: U stays (R,R), np.dot(U, Sqrt) fails
        raise ValueError("shapes (%d,%d) and (%d,%d) not aligned" % (R, R, nS, nS))
    return nS


def warm_feedback(fails, waits, key, accepted, single_cta=False):
    """Host side of the warm-started split's backoff (SweepEngine.history -> split_phase).  ``accepted``: the device-side
    gates took the attempt of bond ``key``.  In the first sweeps of a training run the tensors still change a lot between
    visits (the basis of the previous visit is then a poor start: the subspace steps do not converge, or the
    orthonormalisation does not): a refused bond tries again at its next visit.  A second refusal in a row (no gap at m --
    typical for the bonds next to the chain ends, and for more bonds the longer the training runs) sends it back to the
    cold pipeline, then it tries again; an accepted attempt clears the count.
    How long it sits out is a matter of what the two outcomes cost.  ``single_cta`` (n = 128, k_fast_split): a refused
    attempt is followed by the single-CTA form of the cold pipeline (~2.5 ms more than an accepted one) while a visit on
    the cold cluster pipeline costs 0.1 - 0.3 ms more, so a bond that keeps refusing backs off exponentially: 4, 8, 16
    visits (with 2 every time, 20 sweeps into the bench run a third of those bonds refused in every sweep: 324 -> 345 ms).
    Generic form (n = 256 / 512): the cold pipeline is 2.5x slower than the warm-started one, sitting out is nearly as
    expensive as a refusal: two visits, as before."""
    if accepted:
        fails.pop(key, None)
        return
    n = fails.get(key, 0) + 1
    fails[key] = n
    if n < 2:
        waits[key] = 0
    elif not single_cta:
        waits[key] = 2
    else:
        waits[key] = min(4 << (n - 2), 16)


def kept_ratio(svals_row, nshort):
    """(sigma_m / sigma_1)^2 of a recorded split with short side ``nshort`` = 2 m: how far the KEPT singular values of
    that bond are graded (nan when the record is incomplete).  The residual gate of the warm-started split (1e-12
    lambda_m) sits at the floor of a Gram-based pass, ~10 eps lambda_1, once this ratio falls below ~0.1: such bonds --
    the ones next to the chain ends -- flip between accepted and refused (tools/refusal_study.py)."""
    m = nshort // 2
    if m < 1 or len(svals_row) < m:
        return float("nan")
    s1, sm = float(svals_row[0]), float(svals_row[m - 1])
    if not (np.isfinite(s1) and np.isfinite(sm) and s1 > 0.0):
        return float("nan")
    return (sm / s1) ** 2


class _Timed:
    """Optional CUDA-event bracket around one C-ABI call (bench.py's live per-kernel timing)."""

    def __init__(self, eng, name, flops, stream=None):
        self.eng, self.name, self.flops, self.stream = eng, name, flops, stream

    def __enter__(self):
        if self.eng.timers is not None:
            if self.stream is None:
                self.stream = torch.cuda.current_stream(self.eng.device)
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(self.stream)

    def __exit__(self, *exc):
        if self.eng.timers is not None:
            self.e1.record(self.stream)
            self.eng.timers.setdefault(self.name, []).append((self.e0, self.e1, self.flops))
        return False


class SweepEngine:
    def __init__(self, S, L, T=0.1, act_fn="linear", loss_fn="cross_entropy", rule="reference", max_bond=None,
                 device=None, group=None, svd_refine=True, dtype="float64", threshold=0.999, min_bond=2):
        if not torch.cuda.is_available():
            raise RuntimeError("tensornetworkforml_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _lib.lib()
        if rule not in ("reference", "fixed", "adaptive"):
            raise ValueError("truncation rule must be 'reference', 'fixed' or 'adaptive'")
        if rule in ("fixed", "adaptive") and not max_bond:
            raise ValueError("rule='%s' needs max_bond" % rule)
        # 'adaptive' (opt-in, NC:890-891 + old_files/TensorNetwork.py:1310-1326): the bond kept depends on the singular
        # values, so every split reads them back (one host synchronisation per bond update)
        self.threshold, self.min_bond = float(threshold), int(min_bond)
        self.S, self.L, self.T = int(S), int(L), float(T)
        self.act, self.loss = ACT[act_fn], LOSS[loss_fn]
        self.rule, self.max_bond = rule, (int(max_bond) if max_bond else None)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        # group: a torch.distributed process group, None (= the default group when torch.distributed is initialised)
        # or False (= never communicate: this replica owns the whole batch)
        self.group = None if group is False else group
        self.world = 1
        if group is not False and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(self.group)
        self.svd_refine = 1 if svd_refine else 0
        # dtype of the per-sample arrays (phi, environments, f, loss derivative): 'float64' = parity path (DMMA);
        # 'float32' = FP32 storage with TF32 tensor-core products (tcgen05) and FP32 accumulation.  Site tensors,
        # bond tensors, the gradient sum, clipping and the SVD split are FP64 in both modes.
        if dtype not in ("float64", "float32"):
            raise ValueError("dtype must be 'float64' or 'float32'")
        self.dtype = dtype
        self.DT = F64 if dtype == "float64" else F32
        self.tdtype = torch.float64 if dtype == "float64" else torch.float32
        self.esz = 8 if dtype == "float64" else 4
        self.sites = [None] * self.S
        self.bonds = [1] * (self.S + 1)
        self.l_pos = 0
        self.label_layout = "R"
        self.Ns = 0
        self.phi = None
        self.env = None
        self.env_cap = 0
        self.y_dev = None
        self.hist = None
        self._pinned = None
        self._pin_evt = None
        self._ws = {}
        self.timers = None            # dict name -> [(event0, event1, flops)] when bench.py switches timing on
        self.overlap_svd = True       # SVD split on a side stream, concurrent with the projection
        self.project_ctas = int(os.environ.get("TNML_PROJECT_CTAS", "140"))   # grid cap of the projection while the SVD
                                                                          # split runs beside it (0 = no cap)
        self._side = None
        self._inflight = None
        # the second Jacobi pass that only refines the reported singular values of the DISCARDED tail runs on a third,
        # normal-priority stream after the split (tnml_svd_split refine = 3 + tnml_svd_split_tail); the workspaces
        # rotate so that the following splits do not wait for it
        self.defer_tail = True
        self._gram_evt = None
        self.project_delay_ns = int(os.environ.get("TNML_PROJECT_DELAY_NS", "10000"))
        self._st = None
        self._ws_cache = {}
        self._split_evt = [None] * 4
        self._ald_for = None          # (p, left_dir) whose activation / loss derivative was computed ahead (split_phase)
        self._tail = None
        self.n_svd_ws = 4             # SVD workspaces in rotation: a tail refinement may lag this many splits - 1
        self._tail_evt = [None] * self.n_svd_ws
        # warm-started split (tnml_svd_split_warm): per (bond, direction) the short-side rotation of the previous visit;
        # the second and later visits of a bond try the deflation path first (svd_fast.cuh), verified on the device
        self.warm_split = os.environ.get("TNML_FAST_SPLIT", "1") != "0"
        self._warm = {}
        # per (bond, direction): visits to sit out before the deflation path is tried again.  history() sees which
        # attempts the device-side gates refused (a refused attempt costs the attempt plus the single-CTA form of the
        # cold pipeline, which is slower than the cluster form): a refused bond goes back to the cold cluster pipeline
        # for the next two visits.
        self._warm_wait = {}
        self._warm_fail = {}
        # opt-in (TNML_FAST_MIN_RATIO, e.g. 0.1; default 0 = off, not yet measured on hardware): a bond whose kept singular
        # values were graded below this ratio at its previous visit (kept_ratio) does not attempt the single-CTA fast
        # split at all -- its residual gate is marginal there and every refusal costs 2.5 - 4 ms
        self.warm_min_ratio = float(os.environ.get("TNML_FAST_MIN_RATIO", "0"))
        self._warm_ratio = {}
        # Beside the single-CTA fast split the projection leaves two SMs free when it is the longer of the two (large
        # per-GPU batches: the split's small multi-CTA kernels then queue for those two SMs, which does not matter), and
        # sixteen when the split is the critical path (small per-GPU batches, i.e. many GPUs)
        self.project_ctas_fast = int(os.environ.get("TNML_PROJECT_CTAS_FAST", "146"))
        self.project_ctas_fast_small = int(os.environ.get("TNML_PROJECT_CTAS_FAST_SMALL", "132"))
        self.project_bound_flops = float(os.environ.get("TNML_PROJECT_BOUND_GFLOP", "10")) * 1e9

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        """Raw handle of the stream the current phase enqueues on (refreshed by _enter at every public entry point: a
        torch.cuda.current_stream() lookup per kernel launch was a measurable part of the host time per bond update)."""
        return self._st if self._st is not None else torch.cuda.current_stream(self.device).cuda_stream

    def _enter(self):
        main = torch.cuda.current_stream(self.device)
        self._st = main.cuda_stream
        return main

    def _ws_bytes(self, name, *dims):
        """Cached tnml_*_workspace_bytes query."""
        key = (name,) + dims
        v = self._ws_cache.get(key)
        if v is None:
            v = getattr(_lib.lib(), name)(*dims)
            self._ws_cache[key] = v
        return v

    def register_input(self, X):
        """Opt-in: page-lock the caller's (Ns, S, 2) float64 C-contiguous array in place.  load_input() then copies
        straight from it (no staging copy); the registry keeps the array alive until it is evicted or unregistered.  The
        caller must not modify X while a load_input(X) is in flight."""
        if not (isinstance(X, np.ndarray) and X.dtype == np.float64 and X.flags["C_CONTIGUOUS"]):
            raise ValueError("register_input needs a C-contiguous float64 ndarray")
        # registering may evict (cudaHostUnregister) the oldest registration: no copy out of it may still be in flight
        torch.cuda.synchronize(self.device)
        return _lib.register_host_array(X)

    def _side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device, priority=-1)   # the SVD is on the critical path
        return self._side

    def _tail_stream(self):
        if self._tail is None:
            self._tail = torch.cuda.Stream(device=self.device)
        return self._tail

    def _join_tail(self):
        """Order the current stream after every deferred tail refinement enqueued so far."""
        if self._tail is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._tail)

    def _empty(self, n):
        return torch.empty(int(n), dtype=torch.float64, device=self.device)

    def _sample_empty(self, n):
        return torch.empty(int(n), dtype=self.tdtype, device=self.device)

    def _weights(self, t, key):
        """Device pointer of a weight operand in the per-sample dtype (an FP32 copy in a workspace for 'float32')."""
        if self.DT == F64:
            return _ptr(t)
        n = t.numel()
        w = self._workspace(key, n * 4)
        call("tnml_convert_f32", _ptr(t), _ptr(w), n, self._stream())
        return _ptr(w)

    def _site_weights(self, p, left_moving):
        """FP32, K-major weight operand of the environment advance over plain site p (float32 variant)."""
        Dl, Dr = self.bonds[p], self.bonds[p + 1]
        w = self._workspace("w32", Dl * 2 * Dr * 4)
        call("tnml_site_weights_f32", _ptr(self.sites[p]), _ptr(w), Dl, Dr, left_moving, self._stream())
        return _ptr(w)

    def _workspace(self, key, nbytes):
        n = max(1, (int(nbytes) + 7) // 8)
        t = self._ws.get(key)
        if t is None or t.numel() < n:
            t = self._empty(n)
            self._ws[key] = t
        return t

    # ------------------------------------------------------------------ site tensors
    def set_sites(self, host_sites, l_pos):
        """host_sites: canonical arrays -- plain (Dl,2,Dr), label site (Dl,2,L,Dr)."""
        assert len(host_sites) == self.S
        self.l_pos = int(l_pos)
        self.label_layout = "R"
        self._warm = {}               # new weights: the rotations of earlier visits say nothing about them
        self._warm_wait, self._warm_fail, self._warm_ratio = {}, {}, {}
        for p, A in enumerate(host_sites):
            A = np.ascontiguousarray(A, dtype=np.float64)
            if p == self.l_pos:
                assert A.ndim == 4 and A.shape[2] == self.L, "label site must be (Dl,2,L,Dr)"
            else:
                assert A.ndim == 3
            self.bonds[p], self.bonds[p + 1] = A.shape[0], A.shape[-1]
            self.sites[p] = torch.from_numpy(A.reshape(-1)).to(self.device)
        assert self.bonds[0] == 1 and self.bonds[self.S] == 1

    def get_sites(self):
        self._label_to("R")
        out = []
        for p, t in enumerate(self.sites):
            a = t.cpu().numpy()
            Dl, Dr = self.bonds[p], self.bonds[p + 1]
            out.append(a.reshape(Dl, 2, self.L, Dr) if p == self.l_pos else a.reshape(Dl, 2, Dr))
        return out

    def scale_sites(self, factor):
        """As[i] /= factor for every site (the calibration rescale of NC:175-176)."""
        for p in range(self.S):
            self.sites[p] = self.sites[p] / factor

    def _label_to(self, layout):
        if layout == self.label_layout:
            return
        p = self.l_pos
        src = self.sites[p]
        dst = torch.empty_like(src)
        call("tnml_label_site_swap", _ptr(src), _ptr(dst), self.bonds[p], self.bonds[p + 1], self.L,
             1 if layout == "L" else 0, F64, self._stream())
        self.sites[p] = dst
        self.label_layout = layout

    # ------------------------------------------------------------------ input
    def _dcap(self):
        mb = max(self.bonds)
        cap = max(mb, min(2 * self.L, 2 * mb))
        if self.max_bond:
            cap = max(cap, self.max_bond)
        return cap

    def _alloc_batch(self, Ns):
        cap = self._dcap()
        if self.env is None or self.Ns != Ns or self.env_cap < cap:
            self.env = None
            self.env = torch.empty((self.S + 1, Ns * cap), dtype=self.tdtype, device=self.device)
            self.env_cap = cap
            self.env[0, :Ns].fill_(1.0)
            self.env[self.S, :Ns].fill_(1.0)
            self.f_buf = [self._sample_empty(Ns * self.L), self._sample_empty(Ns * self.L)]
            # float64: q = lossder * pp, (Ns, L, 4); float32: lossder (Ns, L) followed by an aligned copy of pp
            self.q_buf = self._sample_empty(Ns * self.L * 4 if self.DT == F64 else Ns * self.L + 4 * Ns + 4)
            self.pp_buf = self._sample_empty(Ns * 4)
        if self.phi is None or self.phi.numel() != self.S * Ns * 2:
            self.phi = self._sample_empty(self.S * Ns * 2)
        self.Ns = Ns

    def load_input(self, X):
        """X: (Ns, S, 2) feature-mapped input (what the reference API takes), NumPy on the host or a CUDA tensor."""
        if isinstance(X, torch.Tensor) and X.is_cuda:
            Xd = X.to(torch.float64).contiguous()
            Ns = Xd.shape[0]
        else:
            direct = isinstance(X, np.ndarray) and _lib.is_registered(X)   # register_input(X): page-locked in place
            Xc = np.ascontiguousarray(X, dtype=np.float64)                  # (X itself when it already has this form)
            Ns = Xc.shape[0]
            Xh = torch.from_numpy(Xc).reshape(-1)
            Xd = self._workspace("xstage", Xh.numel() * 8)[:Xh.numel()]
            if direct:
                # the DMA engine reads the caller's buffer directly: no pageable -> pinned staging copy (which cost
                # ~100 ms for the 188 MB batch of config 3)
                Xd.copy_(Xh, non_blocking=True)
            else:
                # owned page-locked staging buffer; the previous transfer out of it must have finished before the host
                # overwrites it
                if self._pin_evt is not None:
                    self._pin_evt.synchronize()
                if self._pinned is None or self._pinned.numel() < Xh.numel():
                    self._pinned = torch.empty(Xh.numel(), dtype=torch.float64).pin_memory()
                pin = self._pinned[:Xh.numel()]
                pin.copy_(Xh)
                Xd.copy_(pin, non_blocking=True)
                if self._pin_evt is None:
                    self._pin_evt = torch.cuda.Event()
                self._pin_evt.record(torch.cuda.current_stream(self.device))
        assert Xd.numel() == Ns * self.S * 2, "input must have shape (Ns, S, 2)"
        self._ald_for = None
        self._alloc_batch(Ns)
        call("tnml_pack_features", _ptr(Xd), _ptr(self.phi), Ns, self.S, self.DT, self._stream())
        self.h2d_bytes = Ns * self.S * 2 * 8

    def load_raw(self, x):
        """x: (Ns, S) raw pixels; the feature map phi = [sin, cos](pi x / 2) runs on the device (DG:165-167)."""
        xd = x.to(torch.float64) if (isinstance(x, torch.Tensor) and x.is_cuda) else torch.from_numpy(
            np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        Ns = xd.shape[0]
        self._ald_for = None
        self._alloc_batch(Ns)
        call("tnml_feature_map", _ptr(xd.contiguous()), _ptr(self.phi), Ns, self.S, self.DT, self._stream())

    def _phi(self, p):
        return self.phi.data_ptr() + p * self.Ns * 2 * self.esz

    def _env(self, p):
        return self.env.data_ptr() + p * self.env.shape[1] * self.esz

    # ------------------------------------------------------------------ forward  (NC:195-258)
    def _advance_right(self, p):
        """env[p+1] = left env of sites <= p   (needs env[p], plain site p)."""
        Dl, Dr = self.bonds[p], self.bonds[p + 1]
        w = _ptr(self.sites[p]) if self.DT == F64 else self._site_weights(p, left_moving=0)
        with _Timed(self, "env_advance", 4.0 * self.Ns * Dl * Dr):
            call("tnml_env_advance", self._env(p), self._phi(p), w, self._env(p + 1), self.Ns, Dl, Dr, self.DT,
                 self._stream())

    def _advance_left(self, p):
        """env[p] = right env of sites >= p   (needs env[p+1], plain site p)."""
        Dl, Dr = self.bonds[p], self.bonds[p + 1]
        if self.DT == F64:
            wt = self._workspace("wt", Dl * 2 * Dr * 8)
            call("tnml_site_transpose", _ptr(self.sites[p]), _ptr(wt), Dl, Dr, F64, self._stream())
            w = _ptr(wt)
        else:
            w = self._site_weights(p, left_moving=1)
        with _Timed(self, "env_advance", 4.0 * self.Ns * Dl * Dr):
            call("tnml_env_advance", self._env(p + 1), self._phi(p), w, self._env(p), self.Ns, Dr, Dl, self.DT,
                 self._stream())

    def forward(self):
        S, l = self.S, self.l_pos
        self._ald_for = None
        self._enter()
        if l == 0:
            for p in range(S - 1, 0, -1):
                self._advance_left(p)
        elif l == S - 1:
            for p in range(0, S - 1):
                self._advance_right(p)
        else:
            raise Exception("forward should not be called if l has an intermediate position (l_pos=%d)" % l)
        self._label_to("R")
        f = self.f_buf[0]
        call("tnml_site_predict", self._env(l), self._phi(l), self._weights(self.sites[l], "w32"), self._env(l + 1),
             _ptr(f), self.Ns, self.bonds[l], self.bonds[l + 1], self.L, self.DT, self._stream())
        self.f_cur = 0
        self._st = None
        return f.view(self.Ns, self.L)

    # ------------------------------------------------------------------ sweep  (NC:384-436)
    def set_labels(self, y):
        self._ald_for = None
        if isinstance(y, torch.Tensor) and y.is_cuda:
            self.y_dev = y.to(torch.int32).contiguous()
        else:
            self.y_dev = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).to(self.device)
        assert self.y_dev.numel() == self.Ns

    def _choose_m(self, left_dir, Dl, R, C):
        return choose_m(self.rule, self.max_bond, left_dir, self.l_pos, self.S, Dl, R, C)

    def _build_norm_stack(self, left_dir):
        S = self.S
        one = torch.ones(1, dtype=torch.float64, device=self.device)
        self.nrmL = [None] * (S + 1)
        self.nrmR = [None] * (S + 1)
        self.nrmL[0] = one
        self.nrmR[S] = one
        if not left_dir:
            for p in range(S - 1, 1, -1):
                self._norm_step(p, left_moving=True)
        else:
            for p in range(0, S - 2):
                self._norm_step(p, left_moving=False)

    def _norm_step(self, p, left_moving, st=None):
        st = self._stream() if st is None else st
        Dl, Dr = self.bonds[p], self.bonds[p + 1]
        ws = self._workspace("nrm", 2 * Dl * Dr * 8)
        if left_moving:
            out = self._empty(Dl * Dl)
            call("tnml_norm_env_step", _ptr(self.nrmR[p + 1]), _ptr(self.sites[p]), _ptr(out), _ptr(ws), Dl, Dr, 1, F64,
                 st)
            self.nrmR[p] = out
        else:
            out = self._empty(Dr * Dr)
            call("tnml_norm_env_step", _ptr(self.nrmL[p]), _ptr(self.sites[p]), _ptr(out), _ptr(ws), Dl, Dr, 0, F64,
                 st)
            self.nrmL[p + 1] = out

    def begin_sweep(self, y, left_dir, L2_flag, nsteps=None):
        S = self.S
        self._enter()
        self.set_labels(y)
        nsteps = S - 1 if nsteps is None else nsteps
        self._join_tail()                             # the previous record may still receive tail singular values
        # both SVD workspaces at their largest size of the chain up front: they are read by the tail stream, so they must
        # not be re-allocated in the middle of a sweep
        cap = self._dcap()
        nb = max(_lib.lib().tnml_svd_split_workspace_bytes(cap, cap, self.L, d) for d in (0, 1))
        for i in range(self.n_svd_ws):
            self._workspace("svd%d" % i, nb)
        self._workspace("gbuf", (cap * 4 * self.L * cap + 4) * 8)   # largest [dB | metrics] of the chain, never re-sized
        self._label_to("L" if left_dir else "R")
        if L2_flag:
            self._build_norm_stack(left_dir)
        nmax = 2 * max(self._dcap(), self.L) * 2
        rec_doubles = _lib.lib().tnml_svd_tail_record_bytes() // 8
        self.hist = dict(tail_recs=torch.zeros((nsteps, rec_doubles), dtype=torch.float64, device=self.device),
                         tail_solved=0,
                         metrics=torch.zeros((nsteps, 4), dtype=torch.float64, device=self.device),
                         stats=torch.zeros((nsteps, 8), dtype=torch.float64, device=self.device),
                         svals=torch.full((nsteps, nmax), float("nan"), dtype=torch.float64, device=self.device),
                         nsv=[], m=[], n=0, fast_keys=[], fast_seen=0, warm_keys=[], warm_seen=0)
        self.hist["tail_recs"][:, 1] = 1.0            # "nothing recorded" until a tail call writes the header
        self._st = None

    def _metrics_now(self, p, q, keep_for=None):
        """Accuracy and MAE of the current prediction against the labels (NC:354-380, NC:697-702): the activation /
        loss-derivative kernel's metric sums, all-reduced over the sample shards; one 32-byte device->host read."""
        self._enter()
        gb = self._workspace("gbuf", 8 * 8)
        met = gb[gb.numel() - 4:]
        if keep_for is not None:
            # the first bond update of the sweep reuses q, pp and these sums (they sit where update_phase expects them)
            Dl, Dr = self.bonds[p], self.bonds[q + 1]
            nB = Dl * 4 * self.L * Dr
            gb = self._workspace("gbuf", (nB + 4) * 8)
            met = gb[nB:nB + 4]
        self._act_lossder(p, q, met)
        self._ald_for = keep_for
        self._st = None
        m = met.clone()
        if self.world > 1:
            torch.distributed.all_reduce(m[:3], group=self.group)
        m = m.cpu().numpy()
        return float(m[0] / m[2]), float(m[1] / (m[2] * self.L))

    def eval_metrics(self):
        """(accuracy, MAE) of the prediction left by forward() for the labels given to set_labels()."""
        p = min(self.l_pos, self.S - 2)
        return self._metrics_now(p, p + 1)

    def first_step_accuracy(self, left_dir):
        """Accuracy of the prediction a sweep starts from (what Network.train prints before the batch optimisation,
        NC:328); call between begin_sweep and the first sweep_step."""
        p = self.l_pos - 1 if left_dir else self.l_pos
        return self._metrics_now(p, p + 1, keep_for=(p, left_dir))[0]

    def sweep_step(self, lr, weight_dec, L2_flag, left_dir):
        """One bond update (NC:440-573 with update_B NC:577-763); everything stays on the device."""
        ctx = self.update_phase(lr, weight_dec, L2_flag, left_dir)
        return self.split_phase(ctx)

    def _act_lossder(self, p, q, met):
        """q, pp and the metric sums of the pair (p, q) from the current prediction f (NC:694-707)."""
        ws = self._workspace("al", self._ws_bytes("tnml_act_lossder_workspace_bytes", self.Ns))
        call("tnml_act_lossder", _ptr(self.f_buf[self.f_cur]), _ptr(self.y_dev), self._phi(p), self._phi(q),
             _ptr(self.q_buf), _ptr(self.pp_buf), _ptr(met), _ptr(ws), self.Ns, self.L, self.act, self.loss, self.T,
             self.DT, self._stream())

    def update_phase(self, lr, weight_dec, L2_flag, left_dir, B_override=None):
        """update_B (NC:577-763): environment advance, loss derivative + metrics, gradient, regularisation, clipping,
        update.  Returns the context the split phase needs; ctx["Bn"] is the updated bond tensor B'."""
        main = self._enter()
        S, L, Ns, st = self.S, self.L, self.Ns, self._st
        l = self.l_pos
        p = l - 1 if left_dir else l
        q = p + 1
        if (not left_dir and not (0 <= l <= S - 2)) or (left_dir and not (1 <= l <= S - 1)):
            raise Exception("l = %d -> position not allowed for %s sweep step" % (l, "left" if left_dir else "right"))
        # the label site may have been switched to the other layout since begin_sweep (get_sites(), i.e. reading
        # Network.As or pickling between two steps of the public per-step API)
        self._label_to("L" if left_dir else "R")
        step = self.hist["n"]
        side = self._side_stream() if self.overlap_svd else main
        Dl, Dm, Dr = self.bonds[p], self.bonds[q], self.bonds[q + 1]
        nB = Dl * 4 * L * Dr
        B, G, Bn = self._empty(nB), (self._empty(nB) if L2_flag else None), self._empty(nB)
        # ---- batch-independent preparation on the side stream (off the critical path: none of it needs the gradient):
        #      norm-environment advance NC:1004-1061, B = A_p . A_q NC:484, L2 derivative E_L.B.E_R NC:1129-1135
        if side is not main:
            side.wait_stream(main)
        sst = side.cuda_stream                          # explicit stream handles: no torch stream context needed
        if L2_flag:
            if not left_dir and p > 0:
                self._norm_step(p - 1, left_moving=False, st=sst)
            if left_dir and q < S - 1:
                self._norm_step(q + 1, left_moving=True, st=sst)
        if B_override is not None:
            with torch.cuda.stream(side):
                B.copy_(B_override.reshape(-1))
        elif not left_dir:
            call("tnml_gemm", 0, 0, Dl * 2 * L, 2 * Dr, Dm, 1.0, _ptr(self.sites[p]), Dm, _ptr(self.sites[q]),
                 2 * Dr, 0.0, _ptr(B), 2 * Dr, F64, sst)
        else:
            call("tnml_gemm", 0, 0, Dl * 2, L * 2 * Dr, Dm, 1.0, _ptr(self.sites[p]), Dm, _ptr(self.sites[q]),
                 L * 2 * Dr, 0.0, _ptr(B), L * 2 * Dr, F64, sst)
        if L2_flag:
            ws_l2 = self._workspace("l2", nB * 8)
            call("tnml_l2_term", _ptr(B), _ptr(self.nrmL[p]), _ptr(self.nrmR[q + 1]), _ptr(G), _ptr(ws_l2), Dl, Dr, L,
                 F64, sst)
        # ---- critical path on the main stream
        # environment advance over the site fixed by the previous step                       NC:637-642 / NC:669-674
        if not left_dir and p > 0:
            self._advance_right(p - 1)
        if left_dir and q < S - 1:
            self._advance_left(q + 1)
        # activation, loss derivative, metrics                                              NC:694-707
        gbuf = self._workspace("gbuf", (nB + 4) * 8)
        dB, met = gbuf[:nB], gbuf[nB:nB + 4]
        if self._ald_for == (p, left_dir):
            self._ald_for = None          # computed ahead by the previous split_phase, behind the SVD split
        else:
            self._act_lossder(p, q, met)
        # gradient: K = Ns tensor-core reduction                                             NC:625-646, NC:710
        ws = self._workspace("grad", self._ws_bytes("tnml_grad_workspace_bytes", Ns, Dl, Dr, L))
        with _Timed(self, "grad", 8.0 * Ns * L * Dl * Dr):
            call("tnml_grad", _ptr(self.q_buf), self._env(p), self._env(q + 1), _ptr(dB), _ptr(ws), Ns, Dl, Dr, L,
                 self.DT, st)
        if self.world > 1:
            with _Timed(self, "allreduce", 0.0):
                reduce_gradient_and_metrics(gbuf, nB, Ns, group=self.group, world=self.world,   # one collective per update
                                            count_written=True)
        call("tnml_copy", self.hist["metrics"].data_ptr() + step * 32, _ptr(met), 32, st)
        # regularisation, clipping, update                                                   NC:728-761
        if side is not main:
            main.wait_stream(side)                      # B and G are ready
        ws = self._workspace("bu", self._ws_bytes("tnml_bond_update_workspace_bytes", Dl, Dr, L))
        call("tnml_bond_update", _ptr(B), _ptr(dB), _ptr(G), _ptr(Bn), self.hist["stats"].data_ptr() + step * 8 * 8,
             _ptr(ws), Dl, Dr, L, float(lr), float(weight_dec), 1 if L2_flag else 0, F64, st)
        self._inflight = (B, Bn, dB, G)
        return dict(B=B, Bn=Bn, G=G, p=p, q=q, Dl=Dl, Dr=Dr, nB=nB, left_dir=left_dir, step=step, main=main, side=side)

    def split_phase(self, ctx):
        """New prediction from the UN-truncated B' (NC:494-523) on the main stream, concurrently with the SVD split +
        truncation + label move (NC:528-563, NC:839-962) on a side stream: neither depends on the other."""
        S, L, Ns, st = self.S, self.L, self.Ns, self._stream()
        Bn, p, q, Dl, Dr, left_dir, step = (ctx[k] for k in ("Bn", "p", "q", "Dl", "Dr", "left_dir", "step"))
        main, side = ctx["main"], ctx["side"]
        R, Cc = (2 * Dl, 2 * L * Dr) if not left_dir else (2 * Dl * L, 2 * Dr)
        m = self._choose_m(left_dir, Dl, R, Cc)
        new_p = self._empty(Dl * 2 * m * (L if left_dir else 1))
        new_q = self._empty(m * 2 * Dr * (1 if left_dir else L))
        adaptive = self.rule == "adaptive"
        defer = bool(self.defer_tail and self.svd_refine == 1 and side is not main and not adaptive)
        par = step % self.n_svd_ws
        ws_svd = self._workspace("svd%d" % par, self._ws_bytes("tnml_svd_split_workspace_bytes", Dl, Dr, L,
                                                               1 if left_dir else 0))
        f_out = self.f_buf[1 - self.f_cur]
        ws = self._workspace("proj", self._ws_bytes("tnml_project_workspace_bytes", Ns, Dl, Dr, L))
        if side is not main:
            side.wait_stream(main)                      # B' is ready
        sv_ptr = self.hist["svals"].data_ptr() + step * self.hist["svals"].shape[1] * 8
        ldir = 1 if left_dir else 0
        # the SVD is issued first: its kernels are short or small and should get SMs before the projection fills the GPU
        if self._tail_evt[par] is not None:
            side.wait_event(self._tail_evt[par])        # this workspace's previous tail refinement has finished
        gram_done = None
        if side is not main:
            if self._gram_evt is None:
                self._gram_evt = torch.cuda.Event()
                self._gram_evt.record(side)              # creates the CUDA event (in the recorded state)
            gram_done = self._gram_evt
        # warm buffer of this (bond, direction); `fast` = the bond was split before in this direction with these shapes
        warm, fast = None, 0
        nshort = min(R, Cc)
        if defer and self.warm_split and nshort in (128, 256, 512) and 2 * m == nshort:
            key = (p, ldir, Dl, Dr)
            warm = self._warm.get(key)
            if warm is None:
                nw = self._ws_bytes("tnml_svd_warm_bytes", Dl, Dr, L, ldir) // 8
                warm = self._warm[key] = torch.zeros(nw, dtype=torch.float64, device=self.device)
                if side is not main:
                    side.wait_stream(main)              # the zero fill ran on the main stream
            elif self._warm_wait.get(key, 0) > 0:
                self._warm_wait[key] -= 1
            else:
                fast = 1
            if self.warm_min_ratio > 0.0:
                if fast and nshort == 128 and self._warm_ratio.get(key, 1.0) < self.warm_min_ratio:
                    fast = 0                            # graded at the previous visit: cold pipeline, no attempt
                self.hist["warm_keys"].append((step, key))
            self.hist["fast_keys"].append((step, key) if fast else None)
        with _Timed(self, "svd_split", 0.0, side):
            call("tnml_svd_split_warm", _ptr(Bn), _ptr(new_p), _ptr(new_q), sv_ptr, _ptr(ws_svd), _ptr(warm), Dl, Dr, L,
                 m, ldir, 3 if defer else self.svd_refine, fast, F64, side.cuda_stream,
                 gram_done.cuda_event if gram_done is not None else None)
        if defer:
            split_done = self._split_evt[par]
            if split_done is None:
                split_done = self._split_evt[par] = torch.cuda.Event()
            split_done.record(side)
        # fast split: is the projection (8 Ns L Dl Dr flops at ~25 TFLOP/s) longer than the ~0.35 ms split?
        project_bound = 8.0 * Ns * L * Dl * Dr > self.project_bound_flops
        fast128 = bool(fast and nshort == 128)        # single-CTA form; the generic form (n = 256 / 512) still uses clusters
        if gram_done is not None and not (fast128 and project_bound):
            # the split's SM-holding Cholesky cluster must be placed before the projection fills the GPU: the projection
            # becomes eligible a few microseconds after the event (without the pause the order was a race that the
            # first process on a fresh box lost: 454 instead of 345 ms per sweep).  With the fast split the event only
            # lets the Gram kernels run unhindered; a projection that is the critical path does not wait for it.
            main.wait_event(gram_done)
            if not fast128:
                call("tnml_delay", self.project_delay_ns, st)
        with _Timed(self, "project", 8.0 * Ns * L * Dl * Dr):
            # beside a cluster-parallel SVD the projection leaves the 8 SMs of the split's cluster free; beside the
            # single-CTA fast split two SMs
            cap = self.project_ctas if (side is not main and min(R, Cc) > 64) else 0
            if fast128 and cap:
                cap = self.project_ctas_fast if project_bound else self.project_ctas_fast_small
            call("tnml_project", _ptr(Bn), _ptr(self.pp_buf), self._env(p), self._env(q + 1), _ptr(f_out), _ptr(ws), Ns,
                 Dl, Dr, L, cap, self.DT, st)
        if defer:
            tail = self._tail_stream()
            tail.wait_event(split_done)
            # these are used by the tail stream after the main stream is done with them: tell the allocator, so that
            # dropping the engine (or re-sizing a workspace) cannot hand their memory out while a tail is pending
            Bn.record_stream(tail)
            ws_svd.record_stream(tail)
            self.hist["svals"].record_stream(tail)
            self.hist["tail_recs"].record_stream(tail)
            # only the small block's Gram matrix is recorded here; history() solves all records of the sweep at once
            rec_ptr = self.hist["tail_recs"].data_ptr() + step * self.hist["tail_recs"].shape[1] * 8
            if warm is not None:
                warm.record_stream(tail)
            call("tnml_svd_split_tail_warm", _ptr(Bn), sv_ptr, _ptr(ws_svd), rec_ptr, _ptr(warm), Dl, Dr, L, m, ldir,
                 fast, F64, tail.cuda_stream)
            evt = self._tail_evt[par]
            if evt is None:
                evt = torch.cuda.Event()
            evt.record(tail)
            self._tail_evt[par] = evt
        self.f_cur = 1 - self.f_cur
        # the next bond update's activation / loss derivative / metrics only need the new prediction: enqueue them now,
        # so that they run while the SVD split is still busy on the side stream
        pn = p - 1 if left_dir else p + 1
        if side is not main and 0 <= pn and pn + 1 <= S - 1:
            # bonds of the next pair (pn, pn + 1); the bond between p and q is about to be set to m
            Dln, Drn = (m, self.bonds[pn + 2]) if not left_dir else (self.bonds[pn], m)
            nBn = Dln * 4 * L * Drn
            gb = self._workspace("gbuf", (nBn + 4) * 8)
            self._act_lossder(pn, pn + 1, gb[nBn:nBn + 4])
            self._ald_for = (pn, left_dir)
        if side is not main:
            main.wait_stream(side)                      # the next step needs the new site tensors
            self._inflight = ctx                        # keep B, B', G alive until the main stream has passed the wait
        if adaptive:
            # m = max(min_bond, min(index, max_bond)) with index = argmax(cumsum(S)/sum(S) > threshold)  (NC:890-891):
            # the factors were computed for the cap, their leading m columns / rows are the truncated factors
            main.synchronize()
            nS = min(R, Cc)
            sv = self.hist["svals"][step, :nS].cpu().numpy()
            index = int(np.argmax(np.cumsum(sv) / sv.sum() > self.threshold))
            m_keep = int(min(nS, max(self.min_bond, min(index, self.max_bond))))
            if m_keep < m:
                new_p = new_p.view(-1, m)[:, :m_keep].contiguous().view(-1)
                new_q = new_q.view(m, -1)[:m_keep].contiguous().view(-1)
                m = m_keep
                self._ald_for = None      # the metric sums computed ahead sit where the cap's bond tensor would end
        self.sites[p], self.sites[q] = new_p, new_q
        self.bonds[q] = m
        self.l_pos += -1 if left_dir else 1                                                 # NC:568-571
        self.hist["nsv"].append(min(R, Cc))
        self.hist["m"].append(m)
        self.hist["n"] = step + 1
        self._st = None
        return f_out.view(Ns, L)

    def sweep(self, y, lr, weight_dec, L2_flag=True, left_dir=False):
        self.begin_sweep(y, left_dir, L2_flag)
        f = None
        for _ in range(self.S - 1):
            f = self.sweep_step(lr, weight_dec, L2_flag, left_dir)
        return f

    def history(self):
        """Fetch the per-step record of the last sweep (one device->host copy)."""
        n = self.hist["n"]
        if self._tail is not None and n > self.hist["tail_solved"]:
            # the recorded small blocks of this sweep: one batched launch (one CTA per bond update) on the tail stream
            k0 = self.hist["tail_solved"]
            with torch.cuda.stream(self._tail):
                call("tnml_svd_tail_batch", self.hist["tail_recs"].data_ptr() + k0 * self.hist["tail_recs"].shape[1] * 8,
                     n - k0, self.hist["svals"].data_ptr() + k0 * self.hist["svals"].shape[1] * 8,
                     self.hist["svals"].shape[1], F64, self._tail.cuda_stream)
            self.hist["tail_solved"] = n
        self._join_tail()
        met = self.hist["metrics"][:n].cpu().numpy()
        stats = self.hist["stats"][:n].cpu().numpy()
        sv = self.hist["svals"][:n].cpu().numpy()
        total = met[:, 2]
        with np.errstate(invalid="ignore", divide="ignore"):      # (a step without recorded metrics reads as nan)
            acc = met[:, 0] / total                               # NC:700
            mae = met[:, 1] / (total * self.L)                    # NC:702
            absf = met[:, 3] / (total * self.L)                   # NC:744 (debug history: mean |f_orig|)
        svals = [sv[i, :self.hist["nsv"][i]] for i in range(n)]
        fk = self.hist["fast_keys"]
        for ent in fk[self.hist["fast_seen"]:]:       # feedback for the next visits of each bond (see _warm_wait)
            if ent is not None:
                nshort = self.hist["nsv"][ent[0]]
                warm_feedback(self._warm_fail, self._warm_wait, ent[1], sv[ent[0], nshort] >= 100, single_cta=nshort == 128)
        self.hist["fast_seen"] = len(fk)
        wk = self.hist["warm_keys"]                   # (only filled when the opt-in spectrum rule is on)
        for step_i, key in wk[self.hist["warm_seen"]:]:
            r = kept_ratio(sv[step_i], self.hist["nsv"][step_i])
            if r == r:
                self._warm_ratio[key] = r
        self.hist["warm_seen"] = len(wk)
        return dict(acc=acc, mae=mae, absf=absf, stats=stats, svals=svals, m=list(self.hist["m"]))

    def bond_dims(self):
        return list(self.bonds[1:self.S])
