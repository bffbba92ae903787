"""MPS classifier with sweeping (DMRG-style) training: drop-in for the reference's ``Network`` (NC:10-1179).

Same constructor, methods, attributes and pickle layout as the reference; the arithmetic runs device-resident on a
B200 through libtnml.so (see ``engine.SweepEngine``).  Additive keyword-only extensions:

  truncation='reference' | 'fixed', max_bond=D   bond kept by the SVD split.  'reference' is the rule of NC:898-910 /
                                                 NC:933-945 (copies the left bond, so interior bonds collapse to 2 and
                                                 L>2 cannot finish a sweep -- SURVEY.md section 0.2); 'fixed' keeps
                                                 min(len(S), max_bond) and cuts both factors.
  device, process_group                          CUDA device; torch.distributed group for sample sharding (each rank
                                                 feeds its own shard of the batch, dB and the metrics are all-reduced).
  truncation='adaptive', threshold=, min_bond=   opt-in: the authors' unfinished adaptive bond (NC:890-891, intent in
                                                 old_files/TensorNetwork.py:1310-1326): m = max(min_bond, min(index,
                                                 max_bond)), index = argmax(cumsum(S)/sum(S) > threshold)
  stable_softmax=True                            opt-in: act_fn='softmax' with the largest logit subtracted before exp
                                                 (NC:794 overflows to nan for |f|/T > 709); same values where finite
  svd_refine                                     second Jacobi pass (default on; full accuracy for small sing. values)
  dtype='float64' | 'float32'                    per-sample arrays and contractions: FP64 (DMMA, the parity path) or FP32
                                                 storage with TF32 tensor-core products on tcgen05 (bond-tensor algebra
                                                 and the SVD split stay FP64)

There is no CPU fallback: without a CUDA device or without libtnml.so the compute methods raise.
"""
from __future__ import annotations

import numpy as np

from .Tensor_class import Tensor

new = np.newaxis

_ACTS = ['linear', 'sigmoid', 'softmax']
_LOSSES = ['MSE', 'cross_entropy', 'full_cross_ent']


def _named_to_canonical(T, p):
    """Reference-style site Tensor (any axis order, names 'left','right','d<p>','l') -> (Dl,2,[L,]Dr) array."""
    names = [str(n) for n in T.axes_names]
    want = [w for w in ("left", "d%d" % p, "l", "right") if w in names]
    arr = np.transpose(np.asarray(T.elem, dtype=np.float64), [names.index(w) for w in want])
    if "left" not in names:
        arr = arr[new, ...]
    if "right" not in names:
        arr = arr[..., new]
    return np.ascontiguousarray(arr)


def _canonical_to_named(A, p, S, is_label):
    """(Dl,2,[L,]Dr) array -> Tensor with the reference's axis names; the dummy edge bonds are dropped."""
    names = ["left", "d%d" % p] + (["l"] if is_label else []) + ["right"]
    if p == S - 1:
        A, names = A[..., 0], names[:-1]
    if p == 0:
        A, names = A[0], names[1:]
    return Tensor(elem=np.array(A, copy=True), axes_names=names)


class Network():
    """Matrix Product State classifier (see the reference docstring NC:11-82 for the attribute list)."""

    def __init__(self, N, M, D=2, L=10, T=0.1, normalize=False, calibration_X=None, act_fn='linear',
                 loss_fn='cross_entropy', check=False, *, truncation='reference', max_bond=None, device=None,
                 process_group=None, svd_refine=True, dtype='float64', threshold=0.999, min_bond=2,
                 stable_softmax=False):
        self.N, self.D, self.L, self.M, self.T = N, D, L, M, T
        if D != 2:
            raise NotImplementedError("the device kernels are written for the 2-component feature map (D=2)")
        assert act_fn in _ACTS, "Please select an activation function between 'linear', 'sigmoid', 'softmax'"
        assert loss_fn in _LOSSES, "Please select a loss function between 'MSE', 'cross_entropy', 'full_cross_ent'"
        self.act_fn, self.loss_fn = act_fn, loss_fn
        self.l_pos = 0
        if dtype not in ('float64', 'float32'):
            raise ValueError("dtype must be 'float64' or 'float32'")
        self._opts = dict(truncation=truncation, max_bond=max_bond, svd_refine=bool(svd_refine), dtype=dtype,
                          threshold=threshold, min_bond=min_bond, stable_softmax=bool(stable_softmax))
        self._device, self._group = device, process_group
        self._eng = None
        self._host_As = None          # list[Tensor]; authoritative when _host_fresh
        self._host_fresh = True
        self._last_f = None

        # weights: same RNG calls, order and shapes as NC:145-148 / NC:186-189 (Tensor draws np.random.random)
        scale = float(M) * 0.5 * 0.64 * D if normalize else 1.
        if normalize:
            print('Normalizing weights...')
            print('Scaling factor: %.2f' % scale)
        As = [Tensor(shape=[L, M, D], axes_names=['l', 'right', 'd0'], scale=scale)]
        for i in range(1, N - 1):
            As.append(Tensor(shape=[M, M, D], axes_names=['left', 'right', 'd' + str(i)], scale=scale))
        As.append(Tensor(shape=[M, D], axes_names=['left', 'd' + str(N - 1)], scale=scale))
        self._host_As = As

        if normalize:
            if calibration_X is None:
                X = np.random.random((16, N))                                             # NC:157-159
                X = np.stack((np.sin(np.pi * X / 2), np.cos(np.pi * X / 2)), axis=-1)
            else:
                X = calibration_X
            B = X.shape[0]
            print('\nCalibrating weights on dataset...')
            f = self.forward(X)
            f_max = self._abs_max(f)
            F2 = f_max ** (1. / N)                                                        # NC:170
            if check:
                print('f_max for random input of %d samples : ' % B, f_max)
            print("Rescaling factor for calibration: ", F2)
            self._engine().scale_sites(F2)                                                # NC:175-176
            self._host_fresh = False
            self.calibration_factor = F2
            f = self.forward(X)                                                           # NC:179
            if check:
                print('f_max for random input of %d samples (after): ' % B, self._abs_max(f))

    # ------------------------------------------------------------------ device plumbing
    def _abs_max(self, f):
        m = float(np.abs(f.elem).max())
        eng = self._eng
        if eng is not None and eng.world > 1:
            from .parallel import global_abs_max
            m = global_abs_max(m, eng.device, group=eng.group, world=eng.world)
        return m

    def _act_name(self):
        return 'softmax_stable' if (self.act_fn == 'softmax' and self._opts.get("stable_softmax", False)) else self.act_fn

    def _engine(self):
        if self._eng is None:
            from .engine import SweepEngine
            self._eng = SweepEngine(self.N, self.L, self.T, self._act_name(), self.loss_fn,
                                    rule=self._opts["truncation"], max_bond=self._opts["max_bond"],
                                    device=self._device, group=self._group, svd_refine=self._opts["svd_refine"],
                                    dtype=self._opts.get("dtype", "float64"),
                                    threshold=self._opts.get("threshold", 0.999), min_bond=self._opts.get("min_bond", 2))
            self._upload()
        elif self._host_fresh and self._host_dirty:
            self._upload()
        return self._eng

    _host_dirty = True

    def _upload(self):
        sites = [_named_to_canonical(T, p) for p, T in enumerate(self._host_As)]
        lab = sites[self.l_pos]
        if lab.ndim != 4:
            raise ValueError("site %d should carry the label axis 'l' (l_pos=%d)" % (self.l_pos, self.l_pos))
        self._eng.set_sites(sites, self.l_pos)
        self._host_dirty = False

    @property
    def As(self):
        """Site tensors as reference-style named Tensors (downloaded from the device when it holds newer values)."""
        if not self._host_fresh:
            eng = self._eng
            sites = eng.get_sites()
            self._host_As = [_canonical_to_named(A, p, self.N, p == eng.l_pos) for p, A in enumerate(sites)]
            self._host_fresh, self._host_dirty = True, False
        return self._host_As

    @As.setter
    def As(self, value):
        self._host_As = list(value)
        self._host_fresh, self._host_dirty = True, True

    def sync_to_device(self):
        """Call after editing ``net.As[i].elem`` in place on the host."""
        self._host_fresh, self._host_dirty = True, True
        self._engine()

    # ------------------------------------------------------------------ forward (NC:195-258)
    def forward(self, X):
        """Predictions (no activation) for X of shape (batch, N, D); returns a Tensor with axes ('l','b')."""
        assert self.N == X.shape[1], "The 1 dimension of the input data must be the flattened number of pixels"
        eng = self._engine()
        eng.load_input(X)
        f_dev = eng.forward()
        out = Tensor(elem=f_dev.t().cpu().numpy().astype(np.float64), axes_names=['l', 'b'])
        self._last_f = out
        return out

    def register_input(self, X):
        """Opt-in: page-lock the caller's (batch, N, 2) float64 array in place; forward(X) then copies straight from it.
        The array is kept alive by the registry; do not modify it while a forward(X) is in flight."""
        return self._engine().register_input(X)

    # ------------------------------------------------------------------ evaluation without host round trips
    def evaluate(self, X, y):
        """Accuracy and MAE of the network on (X, y) -- forward + apply_act_func + accuracy of the reference's test
        scripts (test_diagonals.py:69-78) in one device-resident pass: only three sums come back to the host.  X:
        (batch, N, 2) host array or CUDA tensor; y: integer labels.  MAE is against the one-hot target (NC:702)."""
        eng = self._engine()
        eng.load_input(X)
        eng.forward()
        eng.set_labels(y)
        self._last_f = None
        return eng.eval_metrics()

    def forward_raw(self, x):
        """forward() for RAW pixels x of shape (batch, N) or (batch, d, d) -- host array or CUDA tensor (e.g. from
        data_generator.create_dataset_device): the feature map psi (DG:165-167) runs on the device."""
        eng = self._engine()
        xr = x.reshape(x.shape[0], -1)
        assert self.N == xr.shape[1], "The 1 dimension of the input data must be the flattened number of pixels"
        eng.load_raw(xr)
        f_dev = eng.forward()
        out = Tensor(elem=f_dev.t().cpu().numpy().astype(np.float64), axes_names=['l', 'b'])
        self._last_f = out
        return out

    # ------------------------------------------------------------------ train (NC:261-350)
    def _resident(self, loader):
        """(data, labels) of the loader's dataset on the device, uploaded ONCE, when the dataset is array-backed
        (``.data`` of shape (n, N, 2) and ``.label``, like data_generator.NumpyDataset); None otherwise."""
        ds = getattr(loader, "dataset", None)
        data, label = getattr(ds, "data", None), getattr(ds, "label", None)
        if (data is None or label is None or getattr(loader, "batch_sampler", None) is None
                or getattr(loader, "num_workers", 0) != 0):
            return None
        data = np.asarray(data)
        if data.ndim != 3 or data.shape[1] != self.N or data.shape[2] != 2:
            return None
        import torch
        cache = self.__dict__.setdefault("_resident_cache", {})   # at most two datasets (train + validation), LRU
        hit = cache.pop(id(ds), None)
        if hit is None or hit[0] is not ds:
            dev = self._engine().device
            hit = (ds, torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)).to(dev),
                   torch.from_numpy(np.ascontiguousarray(np.asarray(label), dtype=np.int32)).to(dev))
            while len(cache) >= 2:
                cache.pop(next(iter(cache)))
        cache[id(ds)] = hit
        return hit[1], hit[2]

    @staticmethod
    def _index_batches(loader):
        """The loader's batches as index arrays, with exactly the RNG consumption of iterating the loader itself (the
        iterator's base seed first, then the sampler's draws), without touching the samples.  SubsetRandomSampler and
        SequentialSampler under a plain BatchSampler (what data_generator.prepare_dataset builds, DG:187-192) are
        evaluated vectorised -- iterating a 60 000-index sampler element by element costs more than a sweep; any
        other sampler is iterated through an index-only DataLoader."""
        import torch
        from torch.utils.data import BatchSampler, DataLoader, SequentialSampler, SubsetRandomSampler
        bs, smp = loader.batch_sampler, getattr(loader, "sampler", None)
        if type(bs) is BatchSampler and type(smp) in (SubsetRandomSampler, SequentialSampler):
            def fast():
                torch.empty((), dtype=torch.int64).random_(generator=loader.generator)   # the iterator's base seed
                if type(smp) is SubsetRandomSampler:
                    perm = torch.randperm(len(smp.indices), generator=smp.generator).numpy()
                    order = np.asarray(smp.indices)[perm]
                else:
                    order = np.arange(len(smp))
                n, b = len(order), bs.batch_size
                stop = (n // b) * b if bs.drop_last else n
                for lo in range(0, stop, b):
                    yield order[lo:min(lo + b, stop)]
            return fast()

        class _Indices:
            def __init__(self, n):
                self.n = n

            def __len__(self):
                return self.n

            def __getitem__(self, i):
                return i

        return DataLoader(_Indices(len(loader.dataset)), batch_sampler=bs, collate_fn=list, generator=loader.generator)

    def train(self, train_loader, val_loader, lr, n_epochs=10, weight_dec=0.001, L2_flag=True, debug=False):
        """NC:261-350.  Loaders over an array-backed dataset (what data_generator.prepare_dataset builds) take the
        resident path: the dataset is uploaded once, every batch is gathered on the device from the sampler's indices,
        and nothing but the per-step metrics returns to the host.  Any other loader is collated on the host like the
        reference does (NC:324-325)."""
        import torch
        val_acc, var_hist = [], []
        print("\n --- TRAINING PROCEDURE ---")
        res_train, res_val = self._resident(train_loader), self._resident(val_loader)
        for epoch in range(n_epochs):
            epoch_train_acc = np.zeros(len(train_loader))
            var_hist.append([[] for _ in range(7)] if debug else [[], []])
            batches = self._index_batches(train_loader) if res_train is not None else train_loader
            for i, data in enumerate(batches, 0):
                if res_train is not None:
                    eng = self._engine()
                    idx = torch.as_tensor(np.asarray(data), dtype=torch.long, device=eng.device)
                    eng.load_input(res_train[0].index_select(0, idx))
                    eng.forward()
                    self._last_f = None
                    left_dir = (self.l_pos == self.N - 1)
                    eng.begin_sweep(res_train[1].index_select(0, idx), left_dir, L2_flag)
                    epoch_train_acc[i] = eng.first_step_accuracy(left_dir)   # accuracy before the batch optimisation
                    for _ in range(self.N - 1):
                        eng.sweep_step(lr, weight_dec, L2_flag, left_dir)
                    self._after_steps(var_hist[epoch], debug, L2_flag)
                else:
                    x = np.array([s[0] for s in data])
                    y = np.array([s[1] for s in data])
                    f = self.forward(x)
                    epoch_train_acc[i] = self.accuracy(x, y, f)          # accuracy before the batch optimisation
                    left_dir = (self.l_pos == self.N - 1)
                    f = self.sweep(x, y, f, lr, weight_dec, L2_flag=L2_flag, left_dir=left_dir, var_hist=var_hist[epoch],
                                   debug=debug)
                print('\r' + "Epoch %d/%d - train accuracy : %.4f - completed : %.2f " %
                      (epoch, n_epochs, epoch_train_acc[i], (i + 1) * 100 / len(train_loader)) + '%', end=' ')
            epoch_val_acc = np.zeros(len(val_loader))
            batches = self._index_batches(val_loader) if res_val is not None else val_loader
            for i, data in enumerate(batches, 0):
                if res_val is not None:
                    idx = torch.as_tensor(np.asarray(data), dtype=torch.long, device=self._engine().device)
                    epoch_val_acc[i] = self.evaluate(res_val[0].index_select(0, idx), res_val[1].index_select(0, idx))[0]
                else:
                    x = np.array([s[0] for s in data])
                    y = np.array([s[1] for s in data])
                    epoch_val_acc[i] = self.accuracy(x, y)
            val_acc.append(epoch_val_acc.mean())
            print('\r' + "Epoch %d/%d - train accuracy : %.4f - val accuracy: %.4f" %
                  (epoch, n_epochs, epoch_train_acc.mean(), val_acc[-1]))
        return val_acc, np.array(var_hist)

    # ------------------------------------------------------------------ accuracy (NC:354-380)
    def accuracy(self, X, y, f=None):
        if f is None:
            f = self.forward(X)
        y_pred = np.argmax(f.elem, axis=0)
        errors = (np.asarray(y) != y_pred).sum()
        return (len(y_pred) - errors) / len(y_pred)

    # ------------------------------------------------------------------ sweep (NC:384-436)
    def sweep(self, X, y, f, lr, weight_dec, L2_flag=True, left_dir=False, var_hist=None, debug=False):
        """One optimisation sweep over the batch loaded by the last ``forward`` (like the reference, X itself is not
        re-read).  Runs entirely on the device; per-step metrics are fetched once at the end."""
        eng = self._engine()
        if eng.phi is None:
            raise Exception("call forward(X) before sweep: the environments of the batch are built there")
        self._push_f(f)
        y = np.asarray(y)
        eng.begin_sweep(y, left_dir, L2_flag)
        f_dev = None
        for _ in range(self.N - 1):
            f_dev = eng.sweep_step(lr, weight_dec, L2_flag, left_dir)
        self._after_steps(var_hist, debug, L2_flag)
        out = Tensor(elem=f_dev.t().cpu().numpy().astype(np.float64), axes_names=['l', 'b'])
        self._last_f = out
        return out

    def _push_f(self, f):
        """Use the caller's f if it is not the Tensor we returned last (then the device copy is already current)."""
        eng = self._eng
        if f is not None and f is not self._last_f:
            import torch
            host = np.ascontiguousarray(np.asarray(f.elem, dtype=np.float64).T)
            eng.f_buf[eng.f_cur].copy_(torch.from_numpy(host.reshape(-1)))      # converts to the per-sample dtype
            eng._ald_for = None                                                # computed ahead from the old prediction

    def _after_steps(self, var_hist, debug, L2_flag):
        eng = self._eng
        self.l_pos = eng.l_pos
        self._host_fresh = False
        h = eng.history()
        self.last_history = h
        if var_hist is not None:
            for i in range(len(h["acc"])):
                if debug:                                               # NC:741-747
                    if not L2_flag:
                        raise NameError("name 'L2_loss_term' is not defined")   # the reference's latent bug, NC:746
                    var_hist[0].append(h["stats"][i, 4])
                    var_hist[1].append(h["stats"][i, 5])
                    var_hist[2].append(h["acc"][i])
                    var_hist[3].append(h["absf"][i])
                    var_hist[4].append(h["mae"][i])
                    var_hist[5].append(h["stats"][i, 2])
                    var_hist[6].append(h["stats"][i, 6])
                else:                                                   # NC:749-750
                    var_hist[0].append(h["acc"][i])
                    var_hist[1].append(h["mae"][i])

    def sweep_step(self, f, y, lr, batch_size, weight_dec, L2_flag=True, left_dir=False, var_hist=None, debug=False):
        """One bond update (NC:440-573).  ``y`` is the one-hot target (L, batch) like in the reference."""
        eng = self._engine()
        self._push_f(f)
        y = np.asarray(y)
        labels = np.argmax(y, axis=0) if y.ndim == 2 else y
        first = (self.l_pos == (self.N - 1 if left_dir else 0))
        if first or eng.hist is None:
            eng.begin_sweep(labels, left_dir, L2_flag)
        else:                                   # keep the norm-environment stacks, restart the per-call record
            eng.set_labels(labels)
            eng.hist["n"], eng.hist["nsv"], eng.hist["m"], eng.hist["tail_solved"] = 0, [], [], 0
        f_dev = eng.sweep_step(lr, weight_dec, L2_flag, left_dir)
        self._after_steps(var_hist, debug, L2_flag)
        out = Tensor(elem=f_dev.t().cpu().numpy().astype(np.float64), axes_names=['l', 'b'])
        self._last_f = out
        return out

    # ------------------------------------------------------------------ update_B (NC:577-763)
    def update_B(self, B, f_orig, y, lr, weight_dec, L2_flag=True, ldf=0, var_hist=None, debug=False):
        """Gradient step on the bond tensor at the current label position: returns the updated B' as a Tensor with
        B's axis names.  Like the reference it advances the environment cache of the sweep as a side effect and
        appends the pre-update metrics to ``var_hist``.  ``y`` is the one-hot target (L, batch)."""
        eng = self._engine()
        if eng.phi is None:
            raise Exception("call forward(X) before update_B: the environments of the batch are built there")
        left_dir = bool(ldf)
        self._push_f(f_orig)
        y = np.asarray(y)
        labels = np.argmax(y, axis=0) if y.ndim == 2 else y
        first = (self.l_pos == (self.N - 1 if left_dir else 0))
        if first or eng.hist is None:
            eng.begin_sweep(labels, left_dir, L2_flag)
        else:
            eng.set_labels(labels)
            eng.hist["n"], eng.hist["nsv"], eng.hist["m"], eng.hist["tail_solved"] = 0, [], [], 0
        p = self.l_pos - ldf
        import torch
        Bc = self._bond_to_canonical(B, p)
        ctx = eng.update_phase(lr, weight_dec, L2_flag, left_dir,
                               B_override=torch.from_numpy(Bc).to(eng.device))
        torch.cuda.current_stream(eng.device).wait_stream(ctx["side"])
        eng._st = None                       # update_phase without split_phase: drop the cached stream handle
        eng.hist["n"] = 1
        eng.hist["nsv"], eng.hist["m"] = [0], [0]
        h = eng.history()
        if var_hist is not None:
            if debug:
                if not L2_flag:
                    raise NameError("name 'L2_loss_term' is not defined")     # the reference's latent bug, NC:746
                for row, val in enumerate((h["stats"][0, 4], h["stats"][0, 5], h["acc"][0], h["absf"][0], h["mae"][0],
                                           h["stats"][0, 2], h["stats"][0, 6])):
                    var_hist[row].append(val)
            else:
                var_hist[0].append(h["acc"][0])
                var_hist[1].append(h["mae"][0])
        eng.hist["n"], eng.hist["nsv"], eng.hist["m"], eng.hist["tail_solved"] = 0, [], [], 0
        Bn = ctx["Bn"].cpu().numpy().reshape(Bc.shape)
        return self._bond_from_canonical(Bn, B, p)

    def _bond_names(self, p):
        return ["left", "d%d" % p, "l", "d%d" % (p + 1), "right"]

    def _bond_to_canonical(self, B, p):
        """Bond Tensor (names 'left','d<p>','l','d<p+1>','right', any order, edge bonds absent) -> (Dl,2,L,2,Dr)."""
        names = [str(n) for n in B.axes_names]
        want = [w for w in self._bond_names(p) if w in names]
        assert len(want) == len(names), "unexpected axes %s for the bond tensor of sites %d,%d" % (names, p, p + 1)
        arr = np.transpose(np.asarray(B.elem, dtype=np.float64), [names.index(w) for w in want])
        if "left" not in names:
            arr = arr[new, ...]
        if "right" not in names:
            arr = arr[..., new]
        return np.ascontiguousarray(arr)

    def _bond_from_canonical(self, arr, like, p):
        """(Dl,2,L,2,Dr) array -> Tensor with the axes (and axis order) of ``like``."""
        names = [str(n) for n in like.axes_names]
        full = self._bond_names(p)
        if "right" not in names:
            arr, full = arr[..., 0], full[:-1]
        if "left" not in names:
            arr, full = arr[0], full[1:]
        return Tensor(elem=np.ascontiguousarray(np.transpose(arr, [full.index(nm) for nm in names])), axes_names=names)

    # ------------------------------------------------------------------ activation / loss derivative (NC:767-835)
    def _elementwise(self, kind, f_elem, label_axis, y=None):
        import torch
        from . import _lib
        eng = self._engine()
        moved = np.moveaxis(np.asarray(f_elem, dtype=np.float64), label_axis, 0)
        L = moved.shape[0]
        flat = np.ascontiguousarray(moved.reshape(L, -1))
        Ns = flat.shape[1]
        dev = torch.from_numpy(flat).to(eng.device)
        out = torch.empty_like(dev)
        st = torch.cuda.current_stream(eng.device).cuda_stream
        if kind == "act":
            _lib.call("tnml_apply_act", dev.data_ptr(), out.data_ptr(), Ns, L, _lib.ACT[self._act_name()], float(self.T),
                      _lib.F64, st)
        else:
            yb = np.broadcast_to(np.asarray(y, dtype=np.float64), np.asarray(f_elem).shape)
            yd = torch.from_numpy(np.ascontiguousarray(np.moveaxis(yb, label_axis, 0).reshape(L, -1))).to(eng.device)
            _lib.call("tnml_loss_derivative", dev.data_ptr(), yd.data_ptr(), out.data_ptr(), Ns, L,
                      _lib.ACT[self._act_name()], _lib.LOSS[self.loss_fn], float(self.T), _lib.F64, st)
        res = out.cpu().numpy().reshape(moved.shape)
        return np.ascontiguousarray(np.moveaxis(res, 0, label_axis))

    def apply_act_func(self, f):
        """Activation of the network output (NC:767-796): identity, sigmoid(f/T) or (un-stabilised) softmax(f/T) over
        the label axis.  Returns a new Tensor with f's axes."""
        if self.act_fn == 'linear':
            return Tensor(elem=np.array(f.elem, copy=True), axes_names=f.axes_names)
        axis = int(f.ax_to_index('l')) if self.act_fn == 'softmax' else 0
        return Tensor(elem=self._elementwise("act", f.elem, axis), axes_names=f.axes_names)

    def compute_loss_derivate(self, f, y):
        """Derivative of the loss w.r.t. the (activated) output (NC:800-835); y is the one-hot target, f's shape."""
        if self.loss_fn == 'cross_entropy' and self.act_fn == 'softmax':
            print("softmax + cross entropy case")                                          # NC:827
        return Tensor(elem=self._elementwise("loss", f.elem, 0, y=y), axes_names=f.axes_names)

    # ------------------------------------------------------------------ tensor_svd (NC:839-962)
    def tensor_svd(self, T, left_dir=False, threshold=0.999):
        """SVD split of a 2-D Tensor with aggregated axes 'i' (rows) and 'j' (columns): returns (U sqrt(S), sqrt(S) Vh)
        truncated to the bond chosen by the truncation rule, disaggregated like NC:918-925 / NC:953-960."""
        if type(T) != Tensor:
            raise TypeError("This function only support object from the class Tensor")
        if len(T.shape) != 2:
            raise ValueError("This function only support a 2D tensors")
        import torch
        from . import _lib
        from .engine import choose_m
        eng = self._engine()
        R, C = T.elem.shape
        Dl = int(T.aggregations.get('i', {}).get('left', 1))
        m = choose_m(self._opts["truncation"], self._opts["max_bond"], bool(left_dir), self.l_pos, self.N, Dl, R, C)
        mx = torch.from_numpy(np.ascontiguousarray(T.elem, dtype=np.float64)).to(eng.device)
        US = torch.empty((R, m), dtype=torch.float64, device=eng.device)
        SVh = torch.empty((m, C), dtype=torch.float64, device=eng.device)
        sv = torch.empty(min(R, C) + 2, dtype=torch.float64, device=eng.device)
        ws = torch.empty(max(1, _lib.lib().tnml_svd_workspace_bytes(R, C) // 8), dtype=torch.float64, device=eng.device)
        _lib.call("tnml_svd", mx.data_ptr(), US.data_ptr(), SVh.data_ptr(), sv.data_ptr(), ws.data_ptr(), R, C, m,
                  2 if self._opts["svd_refine"] else 0, _lib.F64, torch.cuda.current_stream(eng.device).cuda_stream)
        self.last_singular_values = sv.cpu().numpy()[:min(R, C)]
        TU = Tensor(elem=US.cpu().numpy(), axes_names=['i', 'right'])
        TSVh = Tensor(elem=SVh.cpu().numpy(), axes_names=['left', 'j'])
        TU.aggregations['i'] = T.aggregations['i']
        TSVh.aggregations['j'] = T.aggregations['j']
        TU.disaggregate('i')
        TSVh.disaggregate('j')
        return TU, TSVh

    # ------------------------------------------------------------------ compute_L2_reg (NC:966-1179)
    def compute_L2_reg(self, B, weight_dec=0.001, left_dir=False):
        """(weight_dec * ||net||^2 restricted to B's environments, its derivative 2 weight_dec E_L.B.E_R) for the bond
        tensor at the current label position; the norm environments are contracted from the current site tensors."""
        import torch
        from . import _lib
        eng = self._engine()
        eng._label_to("L" if left_dir else "R")
        p = self.l_pos - int(bool(left_dir))
        Bc = self._bond_to_canonical(B, p)
        Dl, _, L, _, Dr = Bc.shape
        eng._build_norm_stack(bool(left_dir))
        # the far side is complete; contract the near side up to the pair
        if not left_dir:
            for s_ in range(0, p):
                eng._norm_step(s_, left_moving=False)
        else:
            for s_ in range(self.N - 1, p + 1, -1):
                eng._norm_step(s_, left_moving=True)
        Bd = torch.from_numpy(Bc.reshape(-1)).to(eng.device)
        G = torch.empty_like(Bd)
        ws = torch.empty_like(Bd)
        _lib.call("tnml_l2_term", Bd.data_ptr(), eng.nrmL[p].data_ptr(), eng.nrmR[p + 2].data_ptr(), G.data_ptr(),
                  ws.data_ptr(), Dl, Dr, L, _lib.F64, torch.cuda.current_stream(eng.device).cuda_stream)
        Gh = G.cpu().numpy().reshape(Bc.shape)
        loss_term = float(weight_dec * np.sum(Bc * Gh))
        return loss_term, self._bond_from_canonical(2 * weight_dec * Gh, B, p)

    # ------------------------------------------------------------------ cached batch (NC:37-47)
    @property
    def TX(self):
        """The last input batch as the reference's list of (batch, d_i) Tensors (NC:222-225), read from the device."""
        eng = self._eng
        if eng is None or eng.phi is None:
            return []
        phi = eng.phi.view(self.N, eng.Ns, 2).cpu().numpy()
        return [Tensor(elem=phi[i], axes_names=['b', 'd' + str(i)]) for i in range(self.N)]

    def _cum(self, right):
        eng = self._eng
        if eng is None or eng.env is None or self._last_f is None:
            return None
        if (right and self.l_pos != 0) or (not right and self.l_pos != self.N - 1):
            return None
        Ns, S = eng.Ns, self.N
        envs = [eng.env[p, :Ns * eng.bonds[p]].view(Ns, eng.bonds[p]).t().cpu().numpy() for p in range(S + 1)]
        if right:   # r_cum_contraction[i] = sites i..S-1 contracted with the input: ('left', 'b'); [0] is the output
            return [self._last_f] + [Tensor(elem=envs[p], axes_names=['left', 'b']) for p in range(1, S)]
        return [Tensor(elem=envs[p + 1], axes_names=['right', 'b']) for p in range(S - 1)] + [self._last_f]

    @property
    def r_cum_contraction(self):
        """Right cumulative contractions of the last forward (NC:231-242); None unless the label sits at site 0."""
        return self._cum(True)

    @property
    def l_cum_contraction(self):
        """Left cumulative contractions of the last forward (NC:244-255); None unless the label sits at site N-1."""
        return self._cum(False)

    # ------------------------------------------------------------------ pickle layout
    def __getstate__(self):
        As = self.As
        state = dict(N=self.N, D=self.D, L=self.L, M=self.M, T=self.T, As=As, l_pos=self.l_pos,
                     act_fn=self.act_fn, loss_fn=self.loss_fn, TX=[], r_cum_contraction=None,
                     l_cum_contraction=None, tnml_options=dict(self._opts))
        return state

    def __setstate__(self, state):
        state = dict(state)
        self._opts = state.pop("tnml_options", dict(truncation='reference', max_bond=None, svd_refine=True))
        self._host_As = state.pop("As")
        for k in ("TX", "r_cum_contraction", "l_cum_contraction"):
            state.pop(k, None)
        self.__dict__.update(state)
        self._device = self._group = self._eng = self._last_f = None
        self._host_fresh, self._host_dirty = True, True


Network.__module__ = "Network_class"     # pickles stay interchangeable with the reference
