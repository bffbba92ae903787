"""ctypes binding of libtnml.so (the C ABI declared in include/tnml.h).

There is NO CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtnml.so")

F64, F32 = 0, 1
ACT = {"linear": 0, "sigmoid": 1, "softmax": 2, "softmax_stable": 3}
LOSS = {"MSE": 0, "cross_entropy": 1, "full_cross_ent": 2}

# every symbol include/tnml.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
SIGNATURES = {
    "tnml_version": (C.c_int, []),
    "tnml_kernel_launches": (C.c_uint64, []),
    "tnml_error_string": (C.c_char_p, [C.c_int]),
    "tnml_copy": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tnml_delay": (C.c_int, [_i64, _vp]),
    "tnml_host_register": (C.c_int, [_vp, C.c_uint64]),
    "tnml_host_unregister": (C.c_int, [_vp]),
    "tnml_feature_map": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "tnml_generate_dataset": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _f64, _f64, C.c_uint64, _vp]),
    "tnml_pack_features": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp]),
    "tnml_env_advance": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "tnml_site_transpose": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "tnml_site_predict": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "tnml_act_lossder_workspace_bytes": (_i64, [_i64]),
    "tnml_act_lossder": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _f64, _i32, _vp]),
    "tnml_apply_act": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _f64, _i32, _vp]),
    "tnml_loss_derivative": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _f64, _i32, _vp]),
    "tnml_grad_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "tnml_grad": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "tnml_gemm": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _f64, _vp, _i32, _vp, _i32, _f64, _vp, _i32, _i32, _vp]),
    "tnml_project_workspace_bytes": (_i64, [_i64, _i32, _i32, _i32]),
    "tnml_project": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_bond_update_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "tnml_l2_term": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tnml_bond_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _f64, _f64, _i32, _i32, _vp]),
    "tnml_norm_env_step": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tnml_svd_split_workspace_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "tnml_svd_split": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_svd_split_ev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "tnml_svd_split_tail": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_svd_warm_bytes": (_i64, [_i32, _i32, _i32, _i32]),
    "tnml_svd_split_warm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp,
                                      _vp]),
    "tnml_svd_split_tail_warm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_svd_tail_record_bytes": (_i64, []),
    "tnml_svd_tail_batch": (C.c_int, [_vp, _i32, _vp, _i64, _i32, _vp]),
    "tnml_svd_workspace_bytes": (_i64, [_i32, _i32]),
    "tnml_svd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_label_site_swap": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tnml_contract": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i32, _vp]),
    "tnml_convert_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tnml_site_weights_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
}


class TnmlError(RuntimeError):
    pass


_lib = None
launches = 0   # number of C-ABI compute calls issued (each enqueues >= 1 kernel); read by bench.py


def lib():
    """Load libtnml.so once.  Raises ImportError loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libtnml.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the library does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


class _Registration:
    """One host buffer page-locked in place.  Holds a STRONG reference to the array, so the memory cannot be freed (and
    its address re-used by another allocation) while the driver still has it mapped."""
    __slots__ = ("arr", "ptr", "nbytes")

    def __init__(self, arr):
        self.arr, self.ptr, self.nbytes = arr, arr.ctypes.data, arr.nbytes


_registered = []        # explicit registrations of this process, most recent last


def register_host_array(arr, keep: int = 3) -> bool:
    """Page-lock the NumPy array ``arr`` in place (cudaHostRegister) so that host->device copies DMA straight from it.
    Opt-in: only arrays the caller hands over explicitly are registered, the registry keeps them alive, and the oldest
    registration is released (cudaHostUnregister) when more than ``keep`` are held.  Returns False when the driver
    refuses -- including 'already registered' for a range this registry does not own."""
    for i, r in enumerate(_registered):
        if r.arr is arr:
            _registered.append(_registered.pop(i))
            return True
    while len(_registered) >= keep:
        unregister_host_array(_registered[0].arr)
    reg = _Registration(arr)
    if lib().tnml_host_register(reg.ptr, reg.nbytes) != 0:
        return False
    _registered.append(reg)
    return True


def unregister_host_array(arr) -> None:
    for i, r in enumerate(_registered):
        if r.arr is arr:
            lib().tnml_host_unregister(r.ptr)
            _registered.pop(i)
            return


def is_registered(arr) -> bool:
    """True when exactly this array object (same memory, still alive by construction) is page-locked by this registry."""
    return any(r.arr is arr for r in _registered)


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().tnml_error_string(rc).decode()
        raise TnmlError("%s failed: %s (code %d)" % (what or "tnml call", msg, rc))


def call(name: str, *args):
    """Invoke a compute entry point, raising TnmlError on a non-zero return code."""
    global launches
    launches += 1
    check(getattr(lib(), name)(*args), name)
