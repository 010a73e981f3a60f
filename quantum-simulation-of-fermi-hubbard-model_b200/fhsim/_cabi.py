"""ctypes binding to ``libfhsim.so`` (the C-ABI declared in ``include/fhsim.h``).

There is no CPU fallback: if the shared library is missing this module raises at import of
:func:`lib`, and every compute call raises when no CUDA device is present (FH_ECUDA).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfhsim.so")

FH_OK, FH_EINVAL, FH_ECUDA, FH_ENOMEM, FH_ESTATE = 0, -1, -2, -3, -4

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p
_vpp = C.POINTER(C.c_void_p)

# name -> (argtypes); every function returns int unless listed in _SPECIAL
SIGNATURES = {
    "fh_ctx_create": [C.c_int, _vp, _vpp],
    "fh_ctx_destroy": [_vp],
    "fh_ctx_sync": [_vp],
    "fh_ctx_info": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
    "fh_ctx_flush_l2": [_vp, C.c_size_t],
    "fh_ctx_timer_start": [_vp],
    "fh_ctx_timer_stop": [_vp, _f64p],
    "fh_state_create": [_vp, C.c_int, _vpp],
    "fh_state_wrap": [_vp, C.c_int, _vp, _vpp],
    "fh_state_destroy": [_vp],
    "fh_state_set_basis": [_vp, C.c_uint64],
    "fh_state_copy": [_vp, _vp],
    "fh_state_to_host": [_vp, _f64p],
    "fh_state_from_host": [_vp, _f64p],
    "fh_state_device_ptr": [_vp, _vpp],
    "fh_state_inner": [_vp, _vp, _f64p, _f64p],
    "fh_state_norm2": [_vp, _f64p],
    "fh_state_swap_bits": [_vp, _vp, C.c_int, _i32p, _i32p],
    "fh_apply_pair": [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, _f64p],
    "fh_apply_diag": [_vp, C.c_int, _u64p, _f64p],
    "fh_apply_pauli_rot_batch": [_vp, C.c_int, _u64p, _u64p, _f64p],
    "fh_table_upload": [_vp, C.c_int, C.c_int, _u64p, _u64p, _f64p, _f64p, _vpp],
    "fh_table_free": [_vp],
    "fh_table_tile_passes": [_vp, C.POINTER(C.c_int)],
    "fh_apply_table_sector": [_vp, _vp, _vp, C.c_int, C.c_int, _f64p, _f64p],
    "fh_table_info": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "fh_apply_table": [_vp, _vp, _vp, _f64p, _f64p],
    "fh_apply_table_accumulate": [_vp, _vp, _vp, _f64p, _f64p],
    "fh_pool_upload": [_vp, C.c_int, C.c_int, _u64p, _u64p, _u64p, _u64p, _f64p, _f64p, _i32p, C.c_int, _vpp],
    "fh_pool_free": [_vp],
    "fh_pool_gradients": [_vp, _vp, _vp, C.c_int, C.c_int, _f64p],
    "fh_pool_gradients_sector": [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _f64p],
    "fh_pool_gradients_sector_masks": [_vp, _vp, _vp, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, _f64p],
    "fh_program_create": [_vp, C.c_int, C.c_int, _vpp],
    "fh_program_destroy": [_vp],
    "fh_program_add_pair": [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_double,
                            C.c_double, C.c_double, _f64p],
    "fh_program_add_diag": [_vp, C.c_int, _u64p, _f64p, C.c_int],
    "fh_program_begin_tile": [_vp, C.c_int, _i32p],
    "fh_program_end_tile": [_vp],
    "fh_program_finalize": [_vp],
    "fh_program_info": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "fh_program_run": [_vp, _vp, _f64p, C.c_int, C.c_int, C.c_int, C.c_int],
    "fh_program_evaluate": [_vp, C.c_uint64, _f64p, C.c_int, C.c_int, _vpp, _f64p, _f64p, _vp, C.c_int, C.c_int,
                            C.c_int, _f64p, C.c_int, _vpp, _f64p, _vp],
    "fh_program_payload_bytes": [_vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)],
    "fh_program_last_stats": [_vp, _f64p, C.POINTER(C.c_int)],
    "fh_program_sector_info": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64), C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "fh_program_time_items": [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _f64p],
    "fh_lanczos": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, _f64p, _vpp, C.POINTER(C.c_int)],
    "fh_ptable_upload": [_vp, C.c_int, C.c_int, _u64p, _u64p, _f64p, _f64p, _vpp],
    "fh_ptable_free": [_vp],
    "fh_ptable_size": [_vp, C.POINTER(C.c_int)],
    "fh_ptable_download": [_vp, _u64p, _u64p, _f64p, _f64p],
    "fh_ptable_dress": [_vp, C.c_uint64, C.c_uint64, C.c_double, C.c_double],
    "fh_comm_unique_id": [C.POINTER(C.c_ubyte)],
    "fh_comm_init": [_vp, C.POINTER(C.c_ubyte), C.c_int, C.c_int, _vpp],
    "fh_comm_destroy": [_vp],
    "fh_comm_info": [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), _f64p],
    "fh_comm_swap_exchange": [_vp, _vp, _vp, _vp, C.c_int, _i32p, _i32p],
    "fh_comm_last_exchange_ms": [_vp, _f64p],
    "fh_comm_all_reduce_sum": [_vp, _f64p, C.c_int],
    "fh_comm_all_gather": [_vp, _f64p, C.c_int, _f64p],
    "fh_comm_barrier": [_vp],
    "fh_lanczos_sector": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint64, _f64p, _vpp, _f64p,
                          C.POINTER(C.c_int), _f64p],
}
_SPECIAL = {"fh_version": ([], C.c_int), "fh_last_error": ([], C.c_char_p)}

_lib = None

# Interpreter shutdown: module globals and the objects they hold are torn down in no particular order, so a table or a
# state can outlive the context its native handle points to (fh_*_free dereferences the context -> use after free).  Once
# the interpreter is exiting, __del__ methods therefore leave the native handles alone; the driver reclaims device memory
# when the process ends.  Explicit close() calls keep working.
import atexit as _atexit

_shutting_down = False


def _mark_shutdown():
    global _shutting_down
    _shutting_down = True


_atexit.register(_mark_shutdown)


def alive() -> bool:
    return not _shutting_down


class FhsimError(RuntimeError):
    pass


def lib():
    """Load libfhsim.so once; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FhsimError(
                f"{LIB_PATH} not found: build it with `make -C quantum-simulation-of-fermi-hubbard-model_b200/csrc` "
                "(or __graft_entry__.build()).  fhsim has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (argtypes, restype) in _SPECIAL.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = handle
    return _lib


def check(rc: int):
    """Map a status code to the exception the reference's Python API would raise."""
    if rc == FH_OK:
        return
    msg = lib().fh_last_error().decode("utf-8", "replace")
    if rc == FH_EINVAL:
        raise ValueError(msg)
    if rc == FH_ENOMEM:
        raise MemoryError(msg)
    raise FhsimError(msg)


def u64_array(values):
    import numpy as np
    a = np.ascontiguousarray(values, dtype=np.uint64)
    return a, a.ctypes.data_as(_u64p)


def f64_array(values):
    import numpy as np
    a = np.ascontiguousarray(values, dtype=np.float64)
    return a, a.ctypes.data_as(_f64p)


def i32_array(values):
    import numpy as np
    a = np.ascontiguousarray(values, dtype=np.int32)
    return a, a.ctypes.data_as(_i32p)
