"""ctypes binding of include/pgcn_b200.h — the C-ABI drop-in boundary.

Nothing here computes: it loads lib/libpgcn_b200.so, declares every exported symbol and turns
negative status codes into RuntimeError. If the library is missing there is no fallback: the
product path fails loudly (the CPU oracle under oracle/ is test infrastructure only).
"""
import ctypes as C
import os

from . import build as _build

P2P_HANDLE_BYTES = 512
NCCL_ID_BYTES = 128

# every symbol declared in include/pgcn_b200.h
SYMBOLS = [
    "pgcn_version", "pgcn_device_count", "pgcn_last_error",
    "pgcn_plan_create", "pgcn_plan_destroy", "pgcn_plan_set_option", "pgcn_plan_get_option",
    "pgcn_plan_autotune", "pgcn_debug_schedule", "pgcn_plan_slab", "pgcn_algorithmic_bytes", "pgcn_launch_count",
    "pgcn_comm_unique_id", "pgcn_comm_init", "pgcn_comm_share", "pgcn_p2p_export", "pgcn_p2p_import",
    "pgcn_spmm", "pgcn_pack", "pgcn_exchange", "pgcn_unpack_add",
    "pgcn_forward", "pgcn_backward", "pgcn_forward_host", "pgcn_forward_host_async", "pgcn_forward_host_wait",
]


class PgcnBytes(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "nnz", "m", "h", "cols_ref", "spmm_fwd", "spmm_bwd", "gather_fwd", "xchg_out", "xchg_in", "pack")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load libpgcn_b200.so (building it first when stale and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.is_stale():
        try:
            _build.build()
        except RuntimeError:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError("libpgcn_b200.so is missing (%s): build it with __graft_entry__.build(); "
                           "there is no CPU fallback for the PGCN hot path" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.pgcn_version.restype = C.c_char_p
    lib.pgcn_version.argtypes = []
    lib.pgcn_device_count.restype = C.c_int
    lib.pgcn_device_count.argtypes = []
    lib.pgcn_last_error.restype = C.c_char_p
    lib.pgcn_last_error.argtypes = [vp]
    lib.pgcn_plan_create.restype = C.c_int
    lib.pgcn_plan_create.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, i32, i32, i32, C.POINTER(vp)]
    lib.pgcn_plan_destroy.restype = C.c_int
    lib.pgcn_plan_destroy.argtypes = [vp]
    lib.pgcn_plan_set_option.restype = C.c_int
    lib.pgcn_plan_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.pgcn_plan_get_option.restype = i64
    lib.pgcn_plan_get_option.argtypes = [vp, C.c_char_p]
    lib.pgcn_plan_autotune.restype = C.c_int
    lib.pgcn_plan_autotune.argtypes = [vp, i32]
    lib.pgcn_debug_schedule.restype = i64
    lib.pgcn_debug_schedule.argtypes = [vp, i32, i64, i64, vp, i64, vp, vp]
    lib.pgcn_plan_slab.restype = vp
    lib.pgcn_plan_slab.argtypes = [vp, C.c_int]
    lib.pgcn_algorithmic_bytes.restype = C.c_int
    lib.pgcn_algorithmic_bytes.argtypes = [vp, i32, C.POINTER(PgcnBytes)]
    lib.pgcn_launch_count.restype = i64
    lib.pgcn_launch_count.argtypes = [vp]
    lib.pgcn_comm_unique_id.restype = C.c_int
    lib.pgcn_comm_unique_id.argtypes = [vp]
    lib.pgcn_comm_init.restype = C.c_int
    lib.pgcn_comm_init.argtypes = [vp, vp]
    lib.pgcn_p2p_export.restype = C.c_int
    lib.pgcn_p2p_export.argtypes = [vp, vp]
    lib.pgcn_p2p_import.restype = C.c_int
    lib.pgcn_p2p_import.argtypes = [vp, vp]
    lib.pgcn_spmm.restype = C.c_int
    lib.pgcn_spmm.argtypes = [vp, C.c_int, vp, vp, vp, vp, i32, vp]
    lib.pgcn_pack.restype = C.c_int
    lib.pgcn_pack.argtypes = [vp, vp, vp, i32, vp]
    lib.pgcn_exchange.restype = C.c_int
    lib.pgcn_exchange.argtypes = [vp, vp, vp, i32, C.c_int, vp]
    lib.pgcn_unpack_add.restype = C.c_int
    lib.pgcn_unpack_add.argtypes = [vp, vp, vp, i32, vp]
    lib.pgcn_forward.restype = C.c_int
    lib.pgcn_forward.argtypes = [vp, vp, vp, i32, vp]
    lib.pgcn_backward.restype = C.c_int
    lib.pgcn_backward.argtypes = [vp, vp, vp, i32, vp]
    lib.pgcn_forward_host.restype = C.c_int
    lib.pgcn_forward_host.argtypes = [vp, vp, vp, i32]
    lib.pgcn_comm_share.restype = C.c_int
    lib.pgcn_comm_share.argtypes = [vp, vp]
    lib.pgcn_forward_host_async.restype = C.c_int
    lib.pgcn_forward_host_async.argtypes = [vp, vp, vp, i32]
    lib.pgcn_forward_host_wait.restype = C.c_int
    lib.pgcn_forward_host_wait.argtypes = [vp]
    _lib = lib
    return lib


def check(rc, plan=None):
    """Raise RuntimeError carrying pgcn_last_error when a C-ABI call returned a negative status."""
    if rc < 0:
        msg = load().pgcn_last_error(plan)
        raise RuntimeError("pgcn_b200 error %d: %s" % (rc, (msg or b"").decode("utf-8", "replace")))
    return rc
