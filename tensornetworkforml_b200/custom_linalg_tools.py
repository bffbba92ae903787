"""Named-axis pair contraction: mirror of the reference's ``contract`` / ``partial_trace`` (CLT:10-189).

``contract`` keeps the reference's signature and side effects (both operands end up with their axes permuted to
(unique, common, contracted) order, CLT:74-75) but the arithmetic runs on the GPU through ``tnml_contract``:
no (unique1 x unique2 x common x contracted) outer product is ever materialised.
"""
from __future__ import annotations

import numpy as np

from .Tensor_class import Tensor


def _as_index_list(T, axes):
    """Names / ints / lists of either -> list of axis positions."""
    if isinstance(axes, (list, tuple, np.ndarray)):
        items = list(axes)
    else:
        items = [axes]
    return [int(T.ax_to_index(a)) if isinstance(a, str) else int(a) for a in items]


def _device_contract(a3: np.ndarray, b3: np.ndarray) -> np.ndarray:
    """out[u1,u2,c] = sum_k a3[u1,c,k] * b3[u2,c,k] on the device."""
    import torch
    from . import _lib
    if not torch.cuda.is_available():
        raise RuntimeError("contract() runs on the GPU through libtnml.so; no CUDA device is available")
    dev = torch.device("cuda", torch.cuda.current_device())
    ta = torch.from_numpy(np.ascontiguousarray(a3, dtype=np.float64)).to(dev)
    tb = torch.from_numpy(np.ascontiguousarray(b3, dtype=np.float64)).to(dev)
    U1, Cc, Kc = a3.shape
    U2 = b3.shape[0]
    out = torch.empty((U1, U2, Cc), dtype=torch.float64, device=dev)
    _lib.call("tnml_contract", ta.data_ptr(), tb.data_ptr(), out.data_ptr(), U1, U2, Cc, Kc, _lib.F64,
              torch.cuda.current_stream(dev).cuda_stream)
    return out.cpu().numpy()


def _contract_(T1, T2, contracted_axis1, contracted_axis2, common_axis1=[], common_axis2=[]):
    """Positional-index version (CLT:10-87).  Result axes: T1's unique axes, T2's unique axes, common axes."""
    assert len(common_axis1) == len(common_axis2), "number of common axes is different"
    if type(contracted_axis1) != list:
        assert T1.shape[contracted_axis1] == T2.shape[contracted_axis2], "dimensions of contracted axes do not match"
        contracted_axis1, contracted_axis2 = [contracted_axis1], [contracted_axis2]
    for i1, i2 in zip(common_axis1, common_axis2):
        assert T1.shape[i1] == T2.shape[i2], "dimensions of common axes do not match"

    def to_tail(T, common, contracted):
        tail = [int(i) for i in common] + [int(i) for i in contracted]
        head = [i for i in range(T.rank) if i not in tail]
        T.transpose(T.axes_names[head + tail])       # in-place on the operand, like CLT:74-75
        return len(head)

    n_c, n_k = len(common_axis1), len(contracted_axis1)
    u1 = to_tail(T1, common_axis1, contracted_axis1)
    u2 = to_tail(T2, common_axis2, contracted_axis2)
    names = np.concatenate([T1.axes_names[:u1], T2.axes_names[:T2.rank - n_k]])
    s1, s2 = T1.elem.shape, T2.elem.shape
    U1 = int(np.prod(s1[:u1], dtype=np.int64))
    U2 = int(np.prod(s2[:u2], dtype=np.int64))
    Cc = int(np.prod(s1[u1:u1 + n_c], dtype=np.int64))
    Kc = int(np.prod(s1[u1 + n_c:], dtype=np.int64))
    out = _device_contract(T1.elem.reshape(U1, Cc, Kc), T2.elem.reshape(U2, Cc, Kc))
    out = out.reshape(tuple(s1[:u1]) + tuple(s2[:u2]) + tuple(s1[u1:u1 + n_c]))
    return Tensor(elem=out, axes_names=names)


def contract(T1, T2, contracted_axis1=[], contracted_axis2=[], common_axis1=[], common_axis2=[], contracted=None,
             common=None):
    """Contract two Tensors over named axes (CLT:90-161).  ``contracted`` / ``common`` are the same-name shortcuts."""
    if contracted is not None:
        contracted_axis1 = contracted_axis2 = contracted
    if common is not None:
        common_axis1 = common_axis2 = common
    if type(common_axis1) != list:
        common_axis1 = [common_axis1]
    if type(common_axis2) != list:
        common_axis2 = [common_axis2]
    if type(contracted_axis1) == str:
        contracted_axis1 = T1.ax_to_index(contracted_axis1)
    if type(contracted_axis2) == str:
        contracted_axis2 = T2.ax_to_index(contracted_axis2)
    common_axis1 = _as_index_list(T1, common_axis1)
    common_axis2 = _as_index_list(T2, common_axis2)
    return _contract_(T1, T2, contracted_axis1, contracted_axis2, common_axis1, common_axis2)


def partial_trace(T, ax1, ax2):
    """Trace two named axes against each other (CLT:164-189); pure index bookkeeping, unused by the sweep."""
    traced = np.array([ax1, ax2])
    rest = np.delete(T.axes_names, T.ax_to_index(traced))
    T.transpose(np.concatenate((traced, rest)))
    return Tensor(elem=T.elem.trace(axis1=0, axis2=1), axes_names=rest)
