"""The C-ABI library loads and exports every symbol the headers under include/ declare.
No compute call can succeed without a GPU: the product has no CPU path."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "**", "*.h"), recursive=True):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"^[A-Za-z_][\w\s\*]*?\b([A-Za-z_]\w*)\s*\([^;{]*\)\s*;", src, flags=re.M):
            if m.group(1) not in ("defined",) and "typedef" not in m.group(0):
                names.add(m.group(1))
    return sorted(names)


def test_library_exports_all_declared_symbols():
    L = cf.lib()
    names = declared_functions()
    assert "cfnmpc_batch_solve" in names and "acados_solve" in names and "crazyflie_acados_solve" in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_version_and_error_strings():
    L = cf.lib()
    assert b"sm_100a" in L.cfnmpc_version()
    assert L.cfnmpc_batch_create(0, 50, 0.015, 0, ctypes.byref(ctypes.c_void_p())) != 0
    assert len(L.cfnmpc_last_error()) > 0


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(cf.CfnmpcError):
        cf.BatchSolver(4, 10)
    L = cf.lib()
    assert L.acados_create() != 0          # the drop-in surface fails loudly too
    cap = L.crazyflie_acados_create_capsule()
    assert L.crazyflie_acados_create(cap) != 0
    L.crazyflie_acados_free_capsule(cap)


def test_python_mirror_rejects_bad_fields_before_touching_the_device():
    s = object.__new__(cf.BatchSolver)
    s.B, s.N = 2, 5
    with pytest.raises(cf.CfnmpcError):
        cf.BatchSolver.set(s, "nonsense", np.zeros(3))
    with pytest.raises(cf.CfnmpcError):
        cf.BatchSolver.set(s, "x0", np.zeros(5))
    with pytest.raises(cf.CfnmpcError):
        cf.BatchSolver.get(s, "nonsense")
