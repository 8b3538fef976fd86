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


def test_headers_compile_as_c99_and_cxx11(tmp_path):
    """The boundary is a C ABI: every shipped header must be usable from plain C (and from C++11, the node's dialect)."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text("""
#include "cfnmpc.h"
#include "acados_solver_crazyflie.h"
#include "acados_sim_solver_crazyflie.h"
#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/sim_interface.h"
#include "acados_c/external_function_interface.h"
#include "acados/utils/print.h"
#include "acados/utils/types.h"
#include "acados/ocp_nlp/ocp_nlp_constraints_bgh.h"
#include "acados/ocp_nlp/ocp_nlp_cost_ls.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "blasfeo/include/blasfeo_d_aux_ext_dep.h"
#include "crazyflie_model/crazyflie_model.h"
int main(void) { return CFNMPC_OK + CRAZYFLIE_N - 50; }
""")
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-c", str(src), "-o", str(tmp_path / "a.o")], check=True)
    subprocess.run(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-I", inc, "-x", "c++", "-c", str(src), "-o", str(tmp_path / "b.o")], check=True)
