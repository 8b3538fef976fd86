"""SURVEY 8f-4, "other nx, nu compile without touching kernels": the pendulum on a cart (nx = 4, nu = 1) through the same
generator (tools/gen_spec.py --model pendulum -> csrc/cf_spec_pendulum.h) and the SAME kernel sources (preparation
program of cf_rti_warp.h + the dense-stage feedback program of cf_pcond_warp.h with block size 1), selected with
-DCF_SPEC_HEADER.  Parity: golden vectors minted by the reference's own acados/HPIPM build with the CasADi-generated
pendulum functions the reference ships (tests/golden/make_golden_pendulum.py), and that library live where it exists."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_dp, _ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
TS = 1.0 / 20


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "pendulum_golden.npz"))


def take(gold, tag):
    return {k: np.ascontiguousarray(gold[f"{tag}_{k}"]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}


@pytest.fixture(scope="module")
def emu():
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu_pendulum.so"))
    L.cfemu_rti_pcond.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                  _ip, _ip, _ip, _ip, _dp, ctypes.c_int, ctypes.c_void_p]
    nx, nu = ctypes.c_int(), ctypes.c_int()
    L.cfemu_dims(ctypes.byref(nx), ctypes.byref(nu))
    assert (nx.value, nu.value) == (4, 1)

    def run(w, N, n_rti=1):
        B = w["x0"].shape[0]
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
        res = np.zeros((B, 4))
        P = lambda a: a.ctypes.data_as(_dp)
        I = lambda a: a.ctypes.data_as(_ip)
        for _ in range(n_rti):
            assert L.cfemu_rti_pcond(B, N, TS, N, None, P(w["x0"]), P(w["yref"]), P(w["yref_e"]), P(x), P(u), I(st), I(it), I(qs), I(fl), P(res), 2, None) == 0
        return dict(x=x, u=u, status=st, qp_iter=it, qp_status=qs, flags=fl)
    return run


def test_generated_header_is_up_to_date():
    if not os.path.isdir("/root/reference/acados/examples/acados_python/pendulum_on_cart"):
        pytest.skip("needs the reference tree")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_spec.py"), "--model", "pendulum", "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = open(os.path.join(ROOT, "crazyflie_nmpc_b200", "csrc", "cf_spec_pendulum.h")).read()
    assert "#define CF_SPEC_NX 4" in src and "#define CF_SPEC_NU 1" in src and "#define CF_SPEC_N 20" in src


@pytest.mark.parametrize("tag,N,n_rti", [("N20_r1", 20, 1), ("N20_r6", 20, 6), ("N7_r1", 7, 1)])
def test_emulated_kernels_match_the_reference_golden(emu, gold, tag, N, n_rti):
    w = take(gold, tag)
    r = emu(w, N, n_rti)
    assert (r["status"] == 0).all() and (r["qp_status"] == 0).all() and (r["flags"] == 0).all()
    assert np.abs(r["qp_iter"] - gold[f"{tag}_qp_iter"][:, -1]).max() <= 1
    assert rel_err(r["x"], gold[f"{tag}_x"]) < 1e-9 and rel_err(r["u"], gold[f"{tag}_u"]) < 1e-9
    assert np.abs(gold[f"{tag}_u"]).max() > 79.0      # the input bound |F| <= 80 is active: the interior-point part is exercised


def test_reference_library_reproduces_the_golden(gold):
    from oracle.oracle import Ref, ref_available
    if not ref_available("pendulum"):
        pytest.skip("oracle/_ref/libcfref_pendulum.so not built (needs /root/reference)")
    w = take(gold, "N20_r1")
    rs = Ref("pendulum").solver(20, TS)
    for i in range(4):
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        rs.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], x, u)
        assert np.array_equal(x, gold["N20_r1_x"][i]) and np.array_equal(u, gold["N20_r1_u"][i])
    rs.close()


# ---------------------------------------------------------------- GPU: the generic-model libraries
def test_generic_library_loads_and_fails_loudly_without_gpu():
    """libcfnmpc_pendulum.so exports the core batch C-ABI and has no CPU path."""
    import crazyflie_nmpc_b200 as cf
    L = ctypes.CDLL(os.path.join(ROOT, "crazyflie_nmpc_b200", "libcfnmpc_pendulum.so"))
    for sym in ("cfnmpc_batch_create", "cfnmpc_batch_destroy", "cfnmpc_batch_set", "cfnmpc_batch_set_option", "cfnmpc_batch_solve",
                "cfnmpc_batch_prepare", "cfnmpc_batch_feedback", "cfnmpc_batch_sync", "cfnmpc_batch_get", "cfnmpc_batch_last_solve_ms",
                "cfnmpc_batch_info", "cfnmpc_last_error", "cfnmpc_version", "cfnmpc_model_dims"):
        assert hasattr(L, sym), sym
    nx, nu, n0, tf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    L.cfnmpc_model_dims(ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(n0), ctypes.byref(tf))
    assert (nx.value, nu.value, n0.value, tf.value) == (4, 1, 20, 1.0)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        with pytest.raises(cf.CfnmpcError):
            cf.ModelSolver("pendulum", 4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag,N,n_rti", [("N20_r1", 20, 1), ("N20_r6", 20, 6), ("N7_r1", 7, 1)])
def test_gpu_pendulum_matches_the_reference_golden(gold, tag, N, n_rti):
    import crazyflie_nmpc_b200 as cf
    w = take(gold, tag)
    B = w["x0"].shape[0]
    with cf.ModelSolver("pendulum", B, N=N, Ts=TS) as s:
        assert (s.nx, s.nu) == (4, 1)
        s.set_problem(w).solve(n_rti)
        x, u, st, it, qs, fl = s.get("x_all"), s.get("u_all"), s.get("status"), s.get("qp_iter"), s.get("qp_status"), s.get("flags")
        assert np.array_equal(s.get("u", 0), u[:, 0]) and np.array_equal(s.get("x", N), x[:, N])
        if n_rti == 1:     # the two phases separately give the same step
            s.set_problem(w).prepare().feedback()
            assert np.array_equal(s.get("x_all"), x) and np.array_equal(s.get("u_all"), u)
        assert s.info("launches") >= 2 * n_rti and s.last_solve_ms() > 0
    assert (st == 0).all() and (qs == 0).all() and (fl == 0).all()
    assert np.abs(it - gold[f"{tag}_qp_iter"][:, -1]).max() <= 1
    assert rel_err(x, gold[f"{tag}_x"]) < 1e-9 and rel_err(u, gold[f"{tag}_u"]) < 1e-9


@pytest.mark.gpu
def test_gpu_pendulum_large_batch_against_live_reference_or_golden(gold):
    """4096 perturbed swing-up starts: every instance solved, a sample against the reference (where it travelled)."""
    import crazyflie_nmpc_b200 as cf
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_pendulum import pendulum_batch
    N, B = 20, 4096
    w = pendulum_batch(B, N, seed=77)
    with cf.ModelSolver("pendulum", B) as s:
        s.set_problem(w).solve(1)
        x, u, st = s.get("x_all"), s.get("u_all"), s.get("status")
    assert (st == 0).all() and np.isfinite(x).all() and np.abs(u).max() <= 80.0 + 1e-9
    from oracle.oracle import Ref, ref_available
    if ref_available("pendulum"):
        rs = Ref("pendulum").solver(N, TS)
        for i in range(0, B, B // 32):
            xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
            rs.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
            assert rel_err(x[i], xr) < 1e-9 and rel_err(u[i], ur) < 1e-9
        rs.close()


@pytest.mark.gpu
def test_gpu_generic_path_on_the_crazyflie_ocp_equals_the_tuned_library(port):
    """The generic path (preparation kernel + dense-stage feedback program, block size 1) built for the Crazyflie
    description solves the same QP as the tuned uncondensed program: agreement to the interior-point arithmetic."""
    import crazyflie_nmpc_b200 as cf
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 50, 32
    w = wl.helix_batch(B, N, seed=12)
    with cf.ModelSolver("crazyflie_generic", B, N=N, Ts=0.015) as g:
        assert (g.nx, g.nu) == (13, 4)
        g.set_problem(w).solve(1)
        xg, ug, stg = g.get("x_all"), g.get("u_all"), g.get("status")
    with cf.BatchSolver(B, N, 0.015) as s:
        s.set_problem(w).solve(1)
        xt, ut = s.get("x_all"), s.get("u_all")
    assert (stg == 0).all() and rel_err(xg, xt) < 1e-9 and rel_err(ug, ut) < 1e-9
    xo, uo = w["x_init"].copy(), w["u_init"].copy()
    port.batch(N, 0.015, w["x0"], w["yref"], w["yref_e"], xo, uo)
    assert rel_err(xg, xo) < 1e-9 and rel_err(ug, uo) < 1e-9


@pytest.mark.gpu
def test_gpu_pendulum_runtime_weights_and_bounds_against_live_reference():
    """Run-time weights and a tighter input box on the second model, against the reference library given the same values."""
    import crazyflie_nmpc_b200 as cf
    from oracle.oracle import Ref, ref_available
    if not ref_available("pendulum"):
        pytest.skip("oracle/_ref/libcfref_pendulum.so did not travel")
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_pendulum import pendulum_batch
    N, B = 20, 8
    w = pendulum_batch(B, N, seed=5)
    W = np.array([500.0, 3000.0, 0.5, 0.05, 0.004])
    We = np.array([800.0, 2500.0, 0.1, 0.2])
    lbu, ubu = np.array([-25.0]), np.array([40.0])
    with cf.ModelSolver("pendulum", B, N=N, Ts=TS) as s:
        s.set("W", W).set("W_e", We).set("lbu", lbu).set("ubu", ubu)
        s.set_problem(w).solve(2)
        x, u, st = s.get("x_all"), s.get("u_all"), s.get("status")
    rs = Ref("pendulum").solver(N, TS)
    rs.set_weights(W, We)
    rs.set_input_bounds(lbu, ubu)
    for i in range(B):
        xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
        for _ in range(2):
            rs.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
        assert st[i] == 0 and rel_err(x[i], xr) < 1e-9 and rel_err(u[i], ur) < 1e-9
    rs.close()
    assert u.min() >= -25.0 - 1e-9 and u.max() <= 40.0 + 1e-9 and (u.max() > 39.9 or u.min() < -24.9)
