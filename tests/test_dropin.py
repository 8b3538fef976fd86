"""Drop-in boundary: the node's own call sequence against the shipped headers and library.

CPU part: tests/dropin/node_sequence.cpp -- written like NMPC::iteration() of the reference
(crazyflie_controller/src/acados_mpc.cpp:61-84,427-718) -- compiles and links against include/ and
libcfnmpc.so, and fails loudly without a GPU.  GPU part: its printed u0/u1/x4 match the oracle.
"""
import os
import subprocess

import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "dropin", "node_sequence.cpp")
EXE = os.path.join(ROOT, "tests", "dropin", "node_sequence")


@pytest.fixture(scope="module")
def exe():
    cf.lib()
    subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                    "-L", os.path.join(ROOT, "crazyflie_nmpc_b200"), "-lcfnmpc",
                    "-Wl,-rpath," + os.path.join(ROOT, "crazyflie_nmpc_b200")], check=True)
    return EXE


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_node_sequence_compiles_links_and_fails_loudly_without_gpu(exe):
    if _has_gpu():
        pytest.skip("GPU present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "acados_create() returned status" in r.stderr


def test_estimator_sequence_compiles_links_and_fails_loudly_without_gpu():
    """acados_estimator.cpp's predictor calls against include/acados_sim_solver_crazyflie.h."""
    cf.lib()
    src = os.path.join(ROOT, "tests", "dropin", "estimator_sequence.cpp")
    out = os.path.join(ROOT, "tests", "dropin", "estimator_sequence")
    subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                    "-L", os.path.join(ROOT, "crazyflie_nmpc_b200"), "-lcfnmpc",
                    "-Wl,-rpath," + os.path.join(ROOT, "crazyflie_nmpc_b200")], check=True)
    if _has_gpu():
        pytest.skip("GPU present")
    r = subprocess.run([out], capture_output=True, text=True)
    assert r.returncode == 3 and "acados_sim_create() returned status" in r.stderr


@pytest.mark.gpu
def test_node_sequence_matches_oracle(exe, port):
    ticks, N, TS = 4, 50, 0.015
    r = subprocess.run([exe, str(ticks)], capture_output=True, text=True, check=True)
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("tick")]
    assert len(rows) == ticks
    x0 = np.array([0.1, -0.05, 0.3, 1, 0, 0, 0, 0.1, 0, -0.1, 0, 0, 0])
    w = wl.single_hover(N, template_iterate=True, node_yref=True, x0=x0)
    x, u = w["x_init"][0].copy(), w["u_init"][0].copy()
    for t, row in enumerate(rows):
        st, info = port.rti(N, TS, w["x0"][0], w["yref"][0], w["yref_e"][0], x, u)
        vals = np.array(row[6:], float)
        assert int(row[3]) == st
        assert np.abs(vals[0:4] - u[0]).max() < 1e-7 and np.abs(vals[4:8] - u[1]).max() < 1e-7
        assert np.abs(vals[8:21] - x[4]).max() < 1e-7
        assert float(row[5]) > 0.0     # nlp_out->total_time is filled (acados_mpc.cpp:616)


@pytest.mark.gpu
def test_capsule_api_through_ctypes(port):
    """Surface B: crazyflie_acados_* with a non-default horizon via create_with_discretization."""
    import ctypes
    L = cf.lib()
    vp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
    L.crazyflie_acados_create_capsule.restype = vp
    for f in ("crazyflie_acados_get_nlp_in", "crazyflie_acados_get_nlp_out", "crazyflie_acados_get_nlp_config",
              "crazyflie_acados_get_nlp_dims", "crazyflie_acados_get_nlp_solver"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.crazyflie_acados_create_with_discretization.argtypes = [vp, ctypes.c_int, dp]
    L.crazyflie_acados_solve.argtypes = [vp]
    L.crazyflie_acados_free.argtypes = [vp]
    L.crazyflie_acados_free_capsule.argtypes = [vp]
    L.ocp_nlp_constraints_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_cost_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_get.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_get.argtypes = [vp, vp, ctypes.c_char_p, vp]
    N, TS = 20, 0.015
    w = wl.helix_batch(1, N, seed=5)
    cap = L.crazyflie_acados_create_capsule()
    steps = (ctypes.c_double * N)(*([TS] * N))
    assert L.crazyflie_acados_create_with_discretization(cap, N, steps) == 0
    cfg, dims, nin, nout, sol = (L.crazyflie_acados_get_nlp_config(cap), L.crazyflie_acados_get_nlp_dims(cap),
                                 L.crazyflie_acados_get_nlp_in(cap), L.crazyflie_acados_get_nlp_out(cap),
                                 L.crazyflie_acados_get_nlp_solver(cap))
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"lbx", P(w["x0"][0])) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"ubx", P(w["x0"][0])) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"bogus", P(w["x0"][0])) != 0
    for k in range(N):
        assert L.ocp_nlp_cost_model_set(cfg, dims, nin, k, b"yref", P(w["yref"][0, k])) == 0
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"u", P(w["u_init"][0, k]))
    assert L.ocp_nlp_cost_model_set(cfg, dims, nin, N, b"yref", P(w["yref_e"][0])) == 0
    for k in range(N + 1):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"x", P(w["x_init"][0, k]))
    assert L.crazyflie_acados_solve(cap) == 0
    x, u = np.zeros((N + 1, 13)), np.zeros((N, 4))
    for k in range(N + 1):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"x", P(x[k]))
    for k in range(N):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"u", P(u[k]))
    qi = ctypes.c_int()
    L.ocp_nlp_get(cfg, sol, b"qp_iter", ctypes.byref(qi))
    xo, uo = w["x_init"][0].copy(), w["u_init"][0].copy()
    st, info = port.rti(N, TS, w["x0"][0], w["yref"][0], w["yref_e"][0], xo, uo)
    assert np.abs(x - xo).max() < 1e-8 and np.abs(u - uo).max() < 1e-8 and abs(qi.value - info.qp_iter) <= 1
    # multipliers of the iterate (acados_c/ocp_nlp_interface.c:576-590), incl. the restored ones of lbx_0 = ubx_0 = x0
    port.record_multipliers(N)
    try:
        xm, um = w["x_init"][0].copy(), w["u_init"][0].copy()
        port.rti(N, TS, w["x0"][0], w["yref"][0], w["yref_e"][0], xm, um)
        ppi, pl0, pl, pt0, pt = port.multipliers()
    finally:
        port.record_multipliers(0)
    pi, lam0, t0, lam, t = np.zeros((N, 13)), np.zeros(34), np.zeros(34), np.zeros((N, 8)), np.zeros((N, 8))
    for k in range(N):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"pi", P(pi[k]))
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"lam", P(lam0) if k == 0 else P(lam[k]))
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"t", P(t0) if k == 0 else P(t[k]))
    assert np.abs(pi - ppi).max() < 1e-8 * (1 + np.abs(ppi).max())
    assert np.abs(lam0.reshape(2, 17) - pl0).max() < 1e-8 * (1 + np.abs(pl0).max()) and np.abs(t0.reshape(2, 17) - pt0).max() < 1e-7
    assert np.abs(lam[1:].reshape(N - 1, 2, 4) - pl).max() < 1e-8 * (1 + np.abs(pl).max()) and np.abs(t[1:].reshape(N - 1, 2, 4) - pt).max() < 1e-7
    # statistics getters of the SQP_RTI module (ocp_nlp_sqp_rti.c:1361-1425)
    t_tot, t_lin, t_qp = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    L.ocp_nlp_get(cfg, sol, b"time_tot", ctypes.byref(t_tot))
    L.ocp_nlp_get(cfg, sol, b"time_lin", ctypes.byref(t_lin))
    L.ocp_nlp_get(cfg, sol, b"time_qp_sol", ctypes.byref(t_qp))
    assert t_tot.value > 0 and t_lin.value > 0 and t_qp.value > 0 and t_lin.value + t_qp.value <= t_tot.value * 1.05
    sm, sn, stat = ctypes.c_int(), ctypes.c_int(), ctypes.POINTER(ctypes.c_double)()
    L.ocp_nlp_get(cfg, sol, b"stat_m", ctypes.byref(sm))
    L.ocp_nlp_get(cfg, sol, b"stat_n", ctypes.byref(sn))
    L.ocp_nlp_get(cfg, sol, b"stat", ctypes.byref(stat))
    assert (sm.value, sn.value) == (2, 2) and (stat[0], stat[1]) == (0.0, float(qi.value))   # [qp_status, qp_iter]
    # qp_cond_N through the reference's own setter: partial condensing, same solution to the IPM tolerances
    L.crazyflie_acados_update_qp_solver_cond_N.argtypes = [vp, ctypes.c_int]
    assert L.crazyflie_acados_update_qp_solver_cond_N(cap, 7) == 0
    for k in range(N):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"u", P(w["u_init"][0, k]))
    for k in range(N + 1):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"x", P(w["x_init"][0, k]))
    assert L.crazyflie_acados_solve(cap) == 0
    uc = np.zeros((N, 4))
    for k in range(N):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"u", P(uc[k]))
    xp, up = w["x_init"][0].copy(), w["u_init"][0].copy()
    port.rti_pcond(N, TS, 7, w["x0"][0], w["yref"][0], w["yref_e"][0], xp, up)
    assert np.abs(uc - up).max() < 1e-8 and np.abs(uc - uo).max() < 1e-5
    L.crazyflie_acados_free(cap)
    L.crazyflie_acados_free_capsule(cap)


@pytest.mark.gpu
def test_capsule_rti_phases_and_time_steps(port, ref):
    """Surface B + acados_c: ocp_nlp_solver_opts_set(.., "rti_phase", ..) splits the iteration
    (ocp_nlp_sqp_rti.c:189-198,1213-1237) and crazyflie_acados_update_time_steps changes the grid of a live solver
    (acados_solver.in.c:133-153) -- against the reference doing the same."""
    import ctypes
    L = cf.lib()
    vp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
    L.crazyflie_acados_create_capsule.restype = vp
    for f in ("crazyflie_acados_get_nlp_in", "crazyflie_acados_get_nlp_out", "crazyflie_acados_get_nlp_config",
              "crazyflie_acados_get_nlp_dims", "crazyflie_acados_get_nlp_solver", "crazyflie_acados_get_nlp_opts"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.crazyflie_acados_create_with_discretization.argtypes = [vp, ctypes.c_int, dp]
    L.crazyflie_acados_update_time_steps.argtypes = [vp, ctypes.c_int, dp]
    L.crazyflie_acados_solve.argtypes = [vp]
    L.crazyflie_acados_free.argtypes = [vp]
    L.crazyflie_acados_free_capsule.argtypes = [vp]
    L.ocp_nlp_constraints_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_cost_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_get.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_solver_opts_set.argtypes = [vp, vp, ctypes.c_char_p, vp]
    N, TS = 20, 0.015
    dt = TS * np.linspace(0.5, 2.0, N)
    w = wl.helix_batch(1, N, seed=6)
    x0_fb = w["x0"][0] + 0.01
    x0_fb[3:7] /= np.linalg.norm(x0_fb[3:7])
    cap = L.crazyflie_acados_create_capsule()
    steps = (ctypes.c_double * N)(*dt)
    assert L.crazyflie_acados_create_with_discretization(cap, N, steps) == 0
    cfg, dims, nin, nout, opts = (L.crazyflie_acados_get_nlp_config(cap), L.crazyflie_acados_get_nlp_dims(cap),
                                  L.crazyflie_acados_get_nlp_in(cap), L.crazyflie_acados_get_nlp_out(cap),
                                  L.crazyflie_acados_get_nlp_opts(cap))
    P = lambda a: ctypes.c_void_p(a.ctypes.data)

    def load(x0):
        x0 = np.ascontiguousarray(x0)
        assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"lbx", P(x0)) == 0
        assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"ubx", P(x0)) == 0

    def phase(v):
        v = ctypes.c_int(v)
        L.ocp_nlp_solver_opts_set(cfg, opts, b"rti_phase", ctypes.byref(v))

    def read():
        x, u = np.zeros((N + 1, 13)), np.zeros((N, 4))
        for k in range(N + 1):
            L.ocp_nlp_out_get(cfg, dims, nout, k, b"x", P(x[k]))
        for k in range(N):
            L.ocp_nlp_out_get(cfg, dims, nout, k, b"u", P(u[k]))
        return x, u

    for k in range(N):
        assert L.ocp_nlp_cost_model_set(cfg, dims, nin, k, b"yref", P(w["yref"][0, k])) == 0
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"u", P(w["u_init"][0, k]))
    assert L.ocp_nlp_cost_model_set(cfg, dims, nin, N, b"yref", P(w["yref_e"][0])) == 0
    for k in range(N + 1):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"x", P(w["x_init"][0, k]))
    # preparation with the old measurement, feedback with the new one, on the non-uniform grid
    load(w["x0"][0])
    phase(1)
    assert L.crazyflie_acados_solve(cap) == 0
    xp, up = read()
    assert np.array_equal(xp, w["x_init"][0]) and np.array_equal(up, w["u_init"][0])   # the preparation moves nothing
    load(x0_fb)
    phase(2)
    assert L.crazyflie_acados_solve(cap) == 0
    x, u = read()
    sr = ref.solver(N, TS, dt=dt)
    xr, ur = w["x_init"][0].copy(), w["u_init"][0].copy()
    assert sr.rti_split(w["x0"][0], x0_fb, w["yref"][0], w["yref_e"][0], xr, ur)[0] == 0
    sr.close()
    assert np.abs(x - xr).max() < 1e-8 and np.abs(u - ur).max() < 1e-8
    # a second feedback without a preparation fails loudly (the reference would silently re-use stale data)
    assert L.crazyflie_acados_solve(cap) != 0
    # back to the uniform grid on the live solver, fused phase: continues from the iterate above
    phase(0)
    steps_u = (ctypes.c_double * N)(*([TS] * N))
    assert L.crazyflie_acados_update_time_steps(cap, N, steps_u) == 0
    assert L.crazyflie_acados_update_time_steps(cap, N + 1, steps_u) != 0
    assert L.crazyflie_acados_solve(cap) == 0
    x2, u2 = read()
    xo, uo = x.copy(), u.copy()
    st, info = port.rti(N, TS, x0_fb, w["yref"][0], w["yref_e"][0], xo, uo)
    assert st == 0 and np.abs(x2 - xo).max() < 1e-8 and np.abs(u2 - uo).max() < 1e-8
    L.crazyflie_acados_free(cap)
    L.crazyflie_acados_free_capsule(cap)


@pytest.mark.gpu
def test_capsule_per_stage_bounds(ref):
    """ocp_nlp_constraints_model_set(.., stage, "lbu"|"ubu", ..) addresses ONE stage (ocp_nlp_constraints_bgh.c:653-674):
    a narrow box on stage 3 only and the FIXED_U0-style pin of stage 0 (acados_mpc.cpp:604-608), against the reference."""
    import ctypes
    L = cf.lib()
    vp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
    L.crazyflie_acados_create_capsule.restype = vp
    for f in ("crazyflie_acados_get_nlp_in", "crazyflie_acados_get_nlp_out", "crazyflie_acados_get_nlp_config", "crazyflie_acados_get_nlp_dims"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.crazyflie_acados_create_with_discretization.argtypes = [vp, ctypes.c_int, dp]
    L.crazyflie_acados_solve.argtypes = [vp]
    L.crazyflie_acados_free.argtypes = [vp]
    L.crazyflie_acados_free_capsule.argtypes = [vp]
    L.ocp_nlp_constraints_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_cost_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_get.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    N, TS = 20, 0.015
    w = wl.hover_batch(1, N, seed=8)
    cap = L.crazyflie_acados_create_capsule()
    steps = (ctypes.c_double * N)(*([TS] * N))
    assert L.crazyflie_acados_create_with_discretization(cap, N, steps) == 0
    cfg, dims, nin, nout = (L.crazyflie_acados_get_nlp_config(cap), L.crazyflie_acados_get_nlp_dims(cap),
                            L.crazyflie_acados_get_nlp_in(cap), L.crazyflie_acados_get_nlp_out(cap))
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    u0 = np.ascontiguousarray(w["u_init"][0, 0] + 0.25)
    lb3, ub3 = np.full(4, 14.0), np.full(4, 16.5)
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"lbx", P(w["x0"][0])) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"ubx", P(w["x0"][0])) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"lbu", P(u0)) == 0        # FIXED_U0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"ubu", P(u0)) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 3, b"lbu", P(lb3)) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, 3, b"ubu", P(ub3)) == 0
    assert L.ocp_nlp_constraints_model_set(cfg, dims, nin, N, b"lbu", P(lb3)) != 0       # no inputs at the terminal stage
    for k in range(N):
        assert L.ocp_nlp_cost_model_set(cfg, dims, nin, k, b"yref", P(w["yref"][0, k])) == 0
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"u", P(w["u_init"][0, k]))
    assert L.ocp_nlp_cost_model_set(cfg, dims, nin, N, b"yref", P(w["yref_e"][0])) == 0
    for k in range(N + 1):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"x", P(w["x_init"][0, k]))
    assert L.crazyflie_acados_solve(cap) == 0
    x, u = np.zeros((N + 1, 13)), np.zeros((N, 4))
    for k in range(N + 1):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"x", P(x[k]))
    for k in range(N):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"u", P(u[k]))
    sr = ref.solver(N, TS)
    sr.set_input_bounds_at(0, u0, u0)
    sr.set_input_bounds_at(3, lb3, ub3)
    xr, ur = w["x_init"][0].copy(), w["u_init"][0].copy()
    assert sr.rti(w["x0"][0], w["yref"][0], w["yref_e"][0], xr, ur)[0] == 0
    sr.close()
    assert np.abs(x - xr).max() < 1e-8 and np.abs(u - ur).max() < 1e-8
    assert np.abs(u[0] - u0).max() < 1e-7 and (u[3] >= lb3 - 1e-6).all() and (u[3] <= ub3 + 1e-6).all()
    L.crazyflie_acados_free(cap)
    L.crazyflie_acados_free_capsule(cap)
