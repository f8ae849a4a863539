"""Build oracle/liboracle_spmm.so (gcc + OpenMP). Test infrastructure; see spmm_oracle.c."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "spmm_oracle.c")
LIB = os.path.join(HERE, "liboracle_spmm.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", tmp]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed: %s\n%s" % (" ".join(cmd), res.stderr))
    os.replace(tmp, LIB)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build())
        vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
        lib.grb_mxm_plus_times_fp32.argtypes = [i64, vp, vp, vp, vp, i64, vp, i32, i32, C.c_int]
        lib.grb_mxm_plus_times_fp32.restype = None
        lib.grb_aggregate.argtypes = [i64, vp, vp, vp, vp, i64, vp, i32, vp]
        lib.grb_aggregate.restype = None
        lib.spmm_csr_fp32.argtypes = [i64, vp, vp, vp, vp, i64, vp]
        lib.spmm_csr_fp32.restype = None
        lib.oracle_num_threads.restype = C.c_int
        lib.oracle_set_threads.argtypes = [C.c_int]
        lib.oracle_set_threads.restype = None
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def grb_aggregate(rowptr, colidx, vals, Hcat, peer_off, m):
    """AH = A*H_own then += A*H_p per peer (Parallel-GCN/main.c:271,295). Hcat = [H_own ; halo]."""
    import numpy as np
    Hcat = np.ascontiguousarray(Hcat, dtype=np.float32)
    f = Hcat.shape[1]
    out = np.empty((m, f), dtype=np.float32)
    po = np.ascontiguousarray(peer_off, dtype=np.int64)
    load().grb_aggregate(m, _p(rowptr), _p(colidx), _p(vals), _p(Hcat), f, _p(out), len(po) - 1, _p(po))
    return out


def spmm_csr(rowptr, colidx, vals, H, m, out=None):
    import numpy as np
    H = np.ascontiguousarray(H, dtype=np.float32)
    f = H.shape[1]
    if out is None:
        out = np.empty((m, f), dtype=np.float32)
    load().spmm_csr_fp32(m, _p(rowptr), _p(colidx), _p(vals), _p(H), f, _p(out))
    return out


def num_threads():
    return int(load().oracle_num_threads())


def set_threads(n):
    load().oracle_set_threads(int(n))


def best_thread_count(fn, candidates):
    """Time `fn` once per candidate thread count and keep the fastest (SMT siblings often hurt a
    bandwidth-bound loop): the CPU baseline is reported at its best configuration."""
    import time
    best, best_t = None, None
    for c in candidates:
        set_threads(c)
        fn()
        t0 = time.perf_counter(); fn(); t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = c, t
    set_threads(best)
    return best


if __name__ == "__main__":
    print(build(force=True))
