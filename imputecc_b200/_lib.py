"""ctypes binding of libcrwr_b200.so (include/crwr_b200.h).  Fails loudly: there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRWR_LIB", os.path.join(_HERE, "libcrwr_b200.so"))

CRWR_F32, CRWR_F64 = 0, 1
CRWR_OK, CRWR_ERR_INVALID, CRWR_ERR_CUDA, CRWR_ERR_CAPACITY, CRWR_ERR_STATE, CRWR_ERR_RETRY = 0, -1, -2, -3, -4, -5
STOP_NAMES = {0: "running", 1: "tolerance", 2: "plateau", 3: "max_steps"}


class Config(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("device", C.c_int32), ("n_contigs", C.c_int64), ("n_markers", C.c_int64),
                ("row_begin", C.c_int64), ("row_end", C.c_int64), ("transition_nnz_cap", C.c_int64),
                ("iterate_cap", C.c_int64), ("n_shards", C.c_int32), ("reserved", C.c_int32)]


class StepRecord(C.Structure):
    _fields_ = [("nnz_in", C.c_int64), ("nnz_new", C.c_int64), ("cut", C.c_double), ("delta", C.c_double),
                ("cut_applied", C.c_int32), ("plateau_hits", C.c_int32), ("fallback_rows", C.c_int32),
                ("max_row_len", C.c_int32), ("max_kept", C.c_int32), ("fast_rows", C.c_int32)]


class Status(C.Structure):
    _fields_ = [("steps_done", C.c_int32), ("stop_reason", C.c_int32), ("capacity_overflow", C.c_int32),
                ("fallback_rows", C.c_int32), ("nnz_iterate", C.c_int64), ("max_row_len", C.c_int64), ("max_kept", C.c_int64),
                ("last_cut", C.c_double), ("last_delta", C.c_double)]


class CrwrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcrwr_b200 error %d: %s" % (code, msg))
        self.code = code


class CapacityError(CrwrError):
    pass


class RetryError(CrwrError):
    """The walk must be run again from crwr_walk_init; the handle has adapted itself."""


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "crwr_abi_version": (C.c_int32, []),
    "crwr_last_error": (C.c_char_p, []),
    "crwr_launch_count": (C.c_int64, [C.c_int32]),
    "crwr_workspace_bytes": (C.c_int64, [C.POINTER(Config)]),
    "crwr_create": (C.c_int32, [C.POINTER(_P), C.POINTER(Config), _P, C.c_int64]),
    "crwr_destroy": (C.c_int32, [_P]),
    "crwr_build_transition": (C.c_int32, [_P, _P, _P, _P, C.c_int64, _P, C.POINTER(C.c_int64), _P]),
    "crwr_set_transition": (C.c_int32, [_P, _P, _P, _P, C.c_int64, _P]),
    "crwr_get_transition": (C.c_int32, [_P, _P, _P, _P, _P]),
    "crwr_transition_nnz": (C.c_int64, [_P]),
    "crwr_walk_init": (C.c_int32, [_P, C.c_double, C.c_double, C.c_double, C.c_int32, _P]),
    "crwr_set_row_hints": (C.c_int32, [_P, C.c_int64, C.c_int64]),
    "crwr_hold_window_select": (C.c_int32, [_P, C.c_int32]),
    "crwr_get_shape": (C.c_int32, [_P, C.POINTER(C.c_int64)]),
    "crwr_get_group_shape": (C.c_int32, [_P, C.POINTER(C.c_int64)]),
    "crwr_profile": (C.c_int32, [_P, C.c_int32]),
    "crwr_profile_phases": (C.c_int32, [_P, C.POINTER(C.c_double)]),
    "crwr_profile_steps": (C.c_int32, [_P, C.POINTER(C.c_double), C.c_int32]),
    "crwr_profile_read": (C.c_int32, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64), _P]),
    "crwr_walk_enqueue": (C.c_int32, [_P, C.c_int32, _P]),
    "crwr_step_propagate": (C.c_int32, [_P, _P]),
    "crwr_step_hist": (C.c_int32, [_P, C.c_int32, _P]),
    "crwr_step_scan": (C.c_int32, [_P, C.c_int32, _P]),
    "crwr_step_gather": (C.c_int32, [_P, _P]),
    "crwr_step_finish": (C.c_int32, [_P, _P, _P]),
    "crwr_exchange_ptr": (_P, [_P, C.c_int32]),
    "crwr_exchange_bytes": (C.c_int64, [_P, C.c_int32]),
    "crwr_peer_block_bytes": (C.c_int64, []),
    "crwr_peer_attach": (C.c_int32, [_P, _P, C.c_int32, C.c_uint64]),
    "crwr_step_peer_scan": (C.c_int32, [_P, C.c_int32, _P]),
    "crwr_step_peer_finish": (C.c_int32, [_P, _P]),
    "crwr_step_window": (C.c_int32, [_P]),
    "crwr_step_peer_tail": (C.c_int32, [_P, _P]),
    "crwr_walk_redo": (C.c_int32, [_P, _P]),
    "crwr_walk_enqueue_peer": (C.c_int32, [_P, C.c_int32, _P]),
    "crwr_walk_poll": (C.c_int32, [_P, C.POINTER(Status), _P]),
    "crwr_walk_trace": (C.c_int32, [_P, C.POINTER(StepRecord), C.c_int32, _P]),
    "crwr_result_count": (C.c_int32, [_P, C.POINTER(C.c_int64), _P]),
    "crwr_result_write": (C.c_int32, [_P, _P, _P, _P, _P]),
    "crwr_iterate_count": (C.c_int32, [_P, C.POINTER(C.c_int64), _P]),
    "crwr_iterate_write": (C.c_int32, [_P, _P, _P, _P, _P]),
    "crwr_symmetrise_scratch_bytes": (C.c_int64, [C.c_int64, C.c_int64]),
    "crwr_symmetrise_count": (C.c_int32, [C.c_int32, C.c_int64, _P, _P, _P, C.c_int64, _P, C.c_int64,
                                          C.POINTER(C.c_int64), _P]),
    "crwr_symmetrise_write": (C.c_int32, [C.c_int32, C.c_int64, _P, _P, _P, C.c_int64, _P, C.c_int64, _P, _P, _P, _P]),
    "crwr_bin_contact_scratch_bytes": (C.c_int64, [C.c_int64]),
    "crwr_bin_contact_weights": (C.c_int32, [C.c_int32, C.c_int64, _P, _P, _P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, _P,
                                             C.c_int64, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build / imputecc_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run `python -m imputecc_b200.build` (needs nvcc). "
                          "imputecc_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.crwr_abi_version() != 1:
        raise ImportError("libcrwr_b200.so ABI mismatch")
    _lib = lib
    return lib


def check(code):
    if code == CRWR_OK:
        return
    msg = load().crwr_last_error().decode("utf-8", "replace")
    if code == CRWR_ERR_CAPACITY:
        raise CapacityError(code, msg)
    if code == CRWR_ERR_RETRY:
        raise RetryError(code, msg)
    if code == CRWR_ERR_INVALID:
        raise ValueError(msg)
    raise CrwrError(code, msg)


def exported_symbols():
    return sorted(_SIGNATURES)
