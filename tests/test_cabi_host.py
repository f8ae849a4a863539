"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/pgcn_b200.h declares, and FAILS LOUDLY without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
import pgcn_b200
from pgcn_b200 import cabi


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pgcn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgcn_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(cabi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = cabi.load()
    for name in header_symbols():
        assert hasattr(lib, name), "libpgcn_b200.so does not export " + name
    assert b"sm_100a" in lib.pgcn_version()


def test_built_for_sm_100a():
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", cabi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_invalid_arguments_are_reported_not_crashing():
    lib = cabi.load()
    out = C.c_void_p()
    rc = lib.pgcn_plan_create(None, None, None, 0, 0, None, None, None, None, None, None, 1, 0, 16, C.byref(out))
    assert rc == -1 and b"null" in lib.pgcn_last_error(None)
    assert lib.pgcn_plan_destroy(None) == 0
    assert lib.pgcn_forward(None, None, None, 16, None) == -1


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = cabi.load()
    rp = np.zeros(2, dtype=np.int32)
    off = np.zeros(2, dtype=np.int64)
    out = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.pgcn_plan_create(p(rp), None, None, 1, 0, p(rp), None, None, None, p(off), p(off), 1, 0, 16, C.byref(out))
    assert rc == -4 and not out.value                      # PGCN_ERR_NOGPU
    assert b"no CPU fallback" in lib.pgcn_last_error(None)
    with pytest.raises(RuntimeError):
        cabi.check(rc)


def test_op_rejects_cpu_tensors():
    import torch
    from pgcn_b200 import op

    class FakePlan:
        m = 4; f_max = 8
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        op.aggregate_forward(FakePlan(), torch.zeros(4, 8))


def test_product_never_imports_oracle():
    pkg = os.path.dirname(pgcn_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "liboracle" not in src, fn
