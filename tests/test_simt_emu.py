"""The CUDA warp program (crazyflie_nmpc_b200/csrc/cf_rti_warp.h) compiled for the host and run with its
32 lanes as lock-step fibers (tests/simt_emu/), against the golden vectors and the oracle.

This exercises the kernel SOURCE -- lane mapping, shared-memory staging, barriers, the mbarrier/bulk-copy
protocol (the emulation poisons staged blocks until the warp waits on the right barrier) -- on machines
without a GPU.  It is test infrastructure: the library has no such path.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
TS = 0.015
_dp, _ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def emu():
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                  _ip, _ip, _ip, _ip, _dp, _dp, ctypes.c_int]

    L.cfemu_rti_batch2.argtypes = L.cfemu_rti_batch.argtypes + [ctypes.POINTER(_dp)]

    def run(w, N, n_rti=1, params=None, per_inst=None, max_ipm_iter=0):
        B = w["x0"].shape[0]
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
        res = np.zeros((B, 4))
        P = lambda a: a.ctypes.data_as(_dp)
        I = lambda a: a.ctypes.data_as(_ip)
        par = None if params is None else np.ascontiguousarray(np.concatenate(params), float)
        pi = None
        if per_inst is not None:   # dict with any of W, W_e, lbu, ubu, lbu0, ubu0 -> [B, width] arrays
            keep = [None if per_inst.get(k) is None else np.ascontiguousarray(per_inst[k], float)
                    for k in ("W", "W_e", "lbu", "ubu", "lbu0", "ubu0")]
            pi = (_dp * 6)(*[P(a) if a is not None else None for a in keep])
        for _ in range(n_rti):
            L.cfemu_rti_batch2(B, N, TS, P(par) if par is not None else None, int(max_ipm_iter), P(w["x0"]), P(w["yref"]), P(w["yref_e"]),
                               P(x), P(u), I(st), I(it), I(qs), I(fl), P(res), None, 4, pi)
        return dict(x=x, u=u, status=st, qp_iter=it, qp_status=qs, flags=fl, res=res)
    return run


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "crazyflie_rti_golden.npz"))


def batch(gold, name, n=None):
    return {k: np.ascontiguousarray(gold[f"{name}_{k}"][:n]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}


@pytest.mark.parametrize("name,N,n", [("hover", 50, 6), ("helix", 50, 6), ("hover20", 20, 4), ("hover100", 100, 2)])
def test_emulated_kernel_matches_golden(emu, gold, name, N, n):
    w = batch(gold, name, n)
    r = emu(w, N)
    assert (r["status"] == 0).all() and (r["flags"] == 0).all() and (r["qp_status"] == 0).all()
    assert np.abs(r["qp_iter"] - gold[f"{name}_qp_iter"][:n]).max() <= 1
    assert rel_err(r["x"], gold[f"{name}_x"][:n]) < 1e-9 and rel_err(r["u"], gold[f"{name}_u"][:n]) < 1e-9


def test_emulated_kernel_config1_degenerate_first_step(emu, gold):
    """u = 0 start: dphi/du = 0, the first step is decided by the cost alone (SURVEY 'B = 0 at u = 0')."""
    for c in (0, 1, 2):
        w = {k: np.ascontiguousarray(gold[f"cfg1_{c}_{k}"]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
        r = emu(w, 50)
        assert rel_err(r["x"], gold[f"cfg1_{c}_x1"]) < 1e-9 and rel_err(r["u"], gold[f"cfg1_{c}_u1"]) < 1e-9
    r5 = emu(w, 50, n_rti=5)
    assert rel_err(r5["x"], gold["cfg1_2_x5"]) < 1e-7 and rel_err(r5["u"], gold["cfg1_2_u5"]) < 1e-7


def test_emulated_kernel_runtime_parameters(emu, gold):
    w = batch(gold, "hover", 2)
    r = emu(w, 50, params=(gold["par_W"], gold["par_WN"], gold["par_lbu"], gold["par_ubu"]))
    assert rel_err(r["x"], gold["par_x"][:2]) < 1e-9 and rel_err(r["u"], gold["par_u"][:2]) < 1e-9


def test_emulated_kernel_tiny_horizons(emu, port):
    from crazyflie_nmpc_b200 import workloads as wl
    for N in (1, 2, 3):
        w = wl.hover_batch(3, N, seed=N)
        r = emu(w, N)
        x, u = w["x_init"].copy(), w["u_init"].copy()
        port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
        assert rel_err(r["x"], x) < 1e-9 and rel_err(r["u"], u) < 1e-9


def per_instance_params(B, seed=5):
    """One weight / bound set per vehicle (SET_WEIGHTS / FIXED_U0 of the node, acados_mpc.cpp:596-608)."""
    rng = np.random.default_rng(seed)
    Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
    W = Q * rng.uniform(0.5, 2.0, (B, 17))
    WN = W[:, :13] * rng.uniform(20, 60, (B, 1))
    lbu = rng.uniform(0.0, 3.0, (B, 4))
    ubu = rng.uniform(19.0, 22.0, (B, 4))
    lbu0 = rng.uniform(10.0, 14.0, (B, 4))   # a narrow stage-0 box around hover
    ubu0 = lbu0 + rng.uniform(1.0, 6.0, (B, 4))
    return dict(W=W, W_e=WN, lbu=lbu, ubu=ubu, lbu0=lbu0, ubu0=ubu0)


def test_emulated_kernel_per_instance_parameters(emu, port, ref):
    """Per-instance W, W_e, input box and a separate stage-0 box, against the port and the reference's own code."""
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 20, 6
    w = wl.hover_batch(B, N, seed=31)
    pp = per_instance_params(B)
    r = emu(w, N, per_inst=pp)
    for i in range(B):
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        p = port.params(Wdiag=pp["W"][i], WNdiag=pp["W_e"][i], lbu=pp["lbu"][i], ubu=pp["ubu"][i], lbu0=pp["lbu0"][i], ubu0=pp["ubu0"][i])
        st, info = port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u, params=p)
        assert st == r["status"][i] and abs(info.qp_iter - r["qp_iter"][i]) <= 1
        assert rel_err(r["x"][i], x) < 1e-9 and rel_err(r["u"][i], u) < 1e-9
        assert (r["u"][i, 0] >= pp["lbu0"][i] - 1e-6).all() and (r["u"][i, 0] <= pp["ubu0"][i] + 1e-6).all()
        s = ref.solver(N, TS)
        s.set_weights(pp["W"][i], pp["W_e"][i])
        s.set_input_bounds(pp["lbu"][i], pp["ubu"][i])
        s.set_input_bounds_stage0(pp["lbu0"][i], pp["ubu0"][i])
        xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
        s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
        s.close()
        assert rel_err(r["x"][i], xr) < 1e-9 and rel_err(r["u"][i], ur) < 1e-9
    # only some arrays given: the rest stays solver-wide
    r2 = emu(w, N, per_inst=dict(lbu0=pp["lbu0"], ubu0=pp["ubu0"]))
    for i in range(2):
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u, params=port.params(lbu0=pp["lbu0"][i], ubu0=pp["ubu0"][i]))
        assert rel_err(r2["x"][i], x) < 1e-9 and rel_err(r2["u"][i], u) < 1e-9


def test_emulated_kernel_nan_measurement(emu, port):
    """A NaN in x0: status ACADOS_QP_FAILURE (4), HPIPM status 3 after one iteration, iterate untouched -- as the
    reference does (ocp_nlp_sqp_rti.c:651-664, x_ocp_qp_ipm.c:2723-2750); other instances unaffected."""
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 20, 5
    w = wl.hover_batch(B, N, seed=77)
    w["x0"][2, 2] = np.nan
    r = emu(w, N)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, it = port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
    assert (r["status"] == st).all() and r["status"][2] == 4 and r["qp_status"][2] == 3 and r["qp_iter"][2] == it[2]
    assert np.array_equal(r["x"][2], w["x_init"][2]) and np.array_equal(r["u"][2], w["u_init"][2])
    ok = np.arange(B) != 2
    assert rel_err(r["x"][ok], x[ok]) < 1e-9 and rel_err(r["u"][ok], u[ok]) < 1e-9


@pytest.fixture(scope="module")
def emu_general():
    """The general kernel variants (per-interval time steps, split phases) under the lock-step emulation."""
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_general.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                    _ip, _ip, _ip, _ip, _dp, ctypes.c_int]

    def run(w, N, dts=None, split=False, x0_fb=None):
        B = w["x0"].shape[0]
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it, qs, fl = [np.full(B, -7, np.int32) for _ in range(4)]
        res = np.zeros((B, 4))
        P = lambda a: a.ctypes.data_as(_dp) if a is not None else None
        I = lambda a: a.ctypes.data_as(_ip)
        dts = None if dts is None else np.ascontiguousarray(dts, float)
        x0_fb = None if x0_fb is None else np.ascontiguousarray(x0_fb, float)
        L.cfemu_rti_general(B, N, TS, P(dts), int(split), P(w["x0"]), P(x0_fb), P(w["yref"]), P(w["yref_e"]), P(x), P(u),
                            I(st), I(it), I(qs), I(fl), P(res), 4)
        return dict(x=x, u=u, status=st, qp_iter=it, qp_status=qs, flags=fl, res=res)
    return run


def moved_measurement(x0, seed, scale=0.02):
    """A second measurement a little away from the first (quaternion re-normalised)."""
    x = x0 + scale * np.random.default_rng(seed).standard_normal(x0.shape)
    x[..., 3:7] /= np.linalg.norm(x[..., 3:7], axis=-1, keepdims=True)
    return np.ascontiguousarray(x)


def test_emulated_general_variant_equals_specialised_kernel(emu, emu_general, gold):
    """VDT variant on a uniform grid and prepare + feedback with an unchanged x0: bit-identical to the fused kernel."""
    w = batch(gold, "helix", 3)
    a = emu(w, 50)
    for kw in (dict(), dict(split=True)):
        b = emu_general(w, 50, **kw)
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["u"], b["u"])
        assert (a["status"] == b["status"]).all() and (a["qp_iter"] == b["qp_iter"]).all() and (b["flags"] == 0).all()


def test_emulated_split_phases_match_reference(emu_general, port, ref):
    """rti_phase 1 with one measurement, rti_phase 2 with the next one (ocp_nlp_sqp_rti.c:495-683), against the
    reference's own two phases and the port."""
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 20, 5
    w = wl.helix_batch(B, N, seed=5)
    x0_fb = moved_measurement(w["x0"], 11)
    r = emu_general(w, N, split=True, x0_fb=x0_fb)
    assert (r["status"] == 0).all() and (r["flags"] == 0).all()
    for i in range(B):
        s = ref.solver(N, TS)
        xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
        st, qi, _ = s.rti_split(w["x0"][i], x0_fb[i], w["yref"][i], w["yref_e"][i], xr, ur)
        s.close()
        assert st == r["status"][i] and abs(qi - r["qp_iter"][i]) <= 1
        assert rel_err(r["x"][i], xr) < 1e-9 and rel_err(r["u"][i], ur) < 1e-9
        xp, up = w["x_init"][i].copy(), w["u_init"][i].copy()
        port.rti_split(N, TS, w["x0"][i], x0_fb[i], w["yref"][i], w["yref_e"][i], xp, up)
        assert rel_err(r["x"][i], xp) < 1e-9 and rel_err(r["u"][i], up) < 1e-9
        assert np.allclose(r["x"][i, 0], x0_fb[i], rtol=0, atol=1e-15)   # x_0 lands on the feedback-time measurement


@pytest.mark.parametrize("split", [False, True])
def test_emulated_nonuniform_grid_matches_reference(emu_general, port, ref, split):
    """One time step per shooting interval = its cost scaling (create_with_discretization,
    acados_solver.in.c:133-153), against the reference built with that grid and the port."""
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 20, 4
    dt = TS * np.concatenate([np.full(6, 0.5), np.linspace(0.6, 2.5, N - 6)])
    w = wl.hover_batch(B, N, seed=9)
    r = emu_general(w, N, dts=dt, split=split)
    assert (r["status"] == 0).all() and (r["flags"] == 0).all()
    port.set_time_steps(dt)
    try:
        for i in range(B):
            s = ref.solver(N, TS, dt=dt)
            xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
            st, qi, _, _ = s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
            s.close()
            assert st == r["status"][i] and abs(qi - r["qp_iter"][i]) <= 1
            assert rel_err(r["x"][i], xr) < 1e-9 and rel_err(r["u"][i], ur) < 1e-9
            xp, up = w["x_init"][i].copy(), w["u_init"][i].copy()
            port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], xp, up)
            assert rel_err(r["x"][i], xp) < 1e-9 and rel_err(r["u"][i], up) < 1e-9
    finally:
        port.set_time_steps(None)
    # and the grid matters: the uniform solution is different
    assert rel_err(emu_general(w, N)["u"], r["u"]) > 1e-4


@pytest.mark.parametrize("name", ["split", "dt", "dtsplit"])
def test_emulated_general_variants_match_phase_golden(emu_general, name):
    """The general kernel variants against fixtures minted by the reference's own two phases / per-interval time steps
    (tests/golden/crazyflie_rti_golden_phases.npz) -- no reference build needed."""
    g = np.load(os.path.join(HERE, "golden", "crazyflie_rti_golden_phases.npz"))
    w = {k: np.ascontiguousarray(g[f"{name}_{k}"]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
    r = emu_general(w, 20, dts=None if name == "split" else g["dt"], split=name != "dt",
                    x0_fb=None if name == "dt" else g[f"{name}_x0_fb"])
    assert (r["status"] == g[f"{name}_status"]).all() and np.abs(r["qp_iter"] - g[f"{name}_qp_iter"]).max() <= 1
    assert (r["flags"] == 0).all()
    assert rel_err(r["x"], g[f"{name}_x"]) < 1e-9 and rel_err(r["u"], g[f"{name}_u"]) < 1e-9


def test_emulated_general_variants_edge_cases(emu_general, port, ref):
    """Tiny horizons (fewer stages than one staging trip of the feedback kernel, N not a multiple of 4), each with its own
    grid and a moved measurement; and a NaN measurement arriving for the feedback phase: QP failure, iterate untouched."""
    from crazyflie_nmpc_b200 import workloads as wl
    for N in (1, 2, 3, 5, 9):
        w = wl.hover_batch(3, N, seed=60 + N)
        dt = TS * np.linspace(0.7, 1.6, N)
        x0_fb = moved_measurement(w["x0"], 70 + N)
        r = emu_general(w, N, dts=dt, split=True, x0_fb=x0_fb)
        port.set_time_steps(dt)
        try:
            for i in range(3):
                x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
                st, info = port.rti_split(N, TS, w["x0"][i], x0_fb[i], w["yref"][i], w["yref_e"][i], x, u)
                assert st == r["status"][i] and abs(info.qp_iter - r["qp_iter"][i]) <= 1
                assert rel_err(r["x"][i], x) < 1e-9 and rel_err(r["u"][i], u) < 1e-9
        finally:
            port.set_time_steps(None)
    N, B = 10, 4
    w = wl.hover_batch(B, N, seed=88)
    x0_fb = w["x0"].copy()
    x0_fb[1, 5] = np.nan
    r = emu_general(w, N, split=True, x0_fb=x0_fb)
    assert r["status"][1] == 4 and r["qp_status"][1] == 3
    assert np.array_equal(r["x"][1], w["x_init"][1]) and np.array_equal(r["u"][1], w["u_init"][1])
    s = ref.solver(N, TS)
    x, u = w["x_init"][1].copy(), w["u_init"][1].copy()
    st, _, qs = s.rti_split(w["x0"][1], x0_fb[1], w["yref"][1], w["yref_e"][1], x, u)
    s.close()
    assert st == 4 and qs == 3 and np.array_equal(x, w["x_init"][1])      # the reference does the same
    ok = np.arange(B) != 1
    assert (r["status"][ok] == 0).all()
    xo, uo = w["x_init"].copy(), w["u_init"].copy()
    port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], xo, uo)
    assert rel_err(r["x"][ok], xo[ok]) < 1e-9 and rel_err(r["u"][ok], uo[ok]) < 1e-9


def stage_bound_table(N, seed=3):
    """A different input box for every stage: [N][8] = lbu | ubu."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(np.concatenate([rng.uniform(0.0, 12.0, (N, 4)), rng.uniform(17.0, 22.0, (N, 4))], axis=1))


@pytest.mark.parametrize("split", [False, True])
def test_emulated_per_stage_input_bounds(emu_general, port, ref, split):
    """Input boxes set one stage at a time (ocp_nlp_constraints_model_set(.., k, "lbu"|"ubu", ..),
    ocp_nlp_constraints_bgh.c:653-674), against the reference driven the same way and the port."""
    from crazyflie_nmpc_b200 import workloads as wl
    N, B = 12, 4
    w = wl.hover_batch(B, N, seed=17)
    tab = stage_bound_table(N)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_set_stage_bounds.argtypes = [_dp]
    L.cfemu_set_stage_bounds(tab.ctypes.data_as(_dp))
    port.set_stage_bounds(tab)
    try:
        r = emu_general(w, N, split=split)
        assert (r["status"] == 0).all() and (r["flags"] == 0).all()
        for i in range(B):
            s = ref.solver(N, TS)
            for k in range(N):
                s.set_input_bounds_at(k, tab[k, :4], tab[k, 4:])
            xr, ur = w["x_init"][i].copy(), w["u_init"][i].copy()
            st, qi, _, _ = s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
            s.close()
            assert st == 0 and abs(qi - r["qp_iter"][i]) <= 1
            assert rel_err(r["x"][i], xr) < 1e-9 and rel_err(r["u"][i], ur) < 1e-9
            xp, up = w["x_init"][i].copy(), w["u_init"][i].copy()
            port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], xp, up)
            assert rel_err(r["x"][i], xp) < 1e-9 and rel_err(r["u"][i], up) < 1e-9
            assert (r["u"][i] >= tab[:, :4] - 1e-6).all() and (r["u"][i] <= tab[:, 4:] + 1e-6).all()
    finally:
        L.cfemu_set_stage_bounds(None)
        port.set_stage_bounds(None)
    assert rel_err(emu_general(w, N, split=split)["u"], r["u"]) > 1e-3     # the table mattered


# ------------------------------------------------------------------ reference edge cases (tests/golden/edge_golden.npz)
XSEL = [1, 4, 50]


@pytest.fixture(scope="module")
def edge():
    return np.load(os.path.join(HERE, "golden", "edge_golden.npz"))


@pytest.mark.parametrize("itmax", [3, 5])
def test_emulated_kernel_qp_maxiter_branch(emu, edge, itmax):
    """HPIPM stops at its iteration limit: the reference applies the step and returns SUCCESS
    (ocp_nlp_sqp_rti.c:651-674); status, qp_status, iteration count and iterate must match the reference's."""
    from crazyflie_nmpc_b200 import workloads as wl
    n = 8
    w = {k: v[:n] for k, v in wl.hover_batch(32, 50, seed=11).items()}
    r = emu(w, 50, max_ipm_iter=itmax)
    assert np.array_equal(r["status"], edge[f"maxiter_{itmax}_status"][:n])
    assert np.array_equal(r["qp_status"], edge[f"maxiter_{itmax}_qp_status"][:n])
    assert np.array_equal(r["qp_iter"], edge[f"maxiter_{itmax}_qp_iter"][:n])
    assert rel_err(r["u"], edge[f"maxiter_{itmax}_u"][:n]) < 1e-9
    assert rel_err(r["x"][:, XSEL], edge[f"maxiter_{itmax}_xsel"][:n]) < 1e-9


def test_emulated_kernel_flags_where_the_reference_nets_fire(emu_general, edge):
    """Ill-conditioned instances: where the reference switched to its LQ factorisation or ran iterative refinement the
    kernel's linear-residual flags must be set; converged, unflagged instances agree with the reference."""
    from crazyflie_nmpc_b200 import workloads as wl
    w = wl.adversarial_batch(128, 50, seed=5)
    fired = np.nonzero((edge["adv_lq"] > 0) | (edge["adv_itref"] > 0))[0]
    calm = np.nonzero((edge["adv_qp_status"] == 0) & (edge["adv_lq"] == 0) & (edge["adv_itref"] == 0))[0][:6]
    sel = np.r_[fired, calm]
    ws = {k: np.ascontiguousarray(w[k][sel]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
    # the linear-residual diagnostics are compiled into the general kernel variants (the library routes lin_res_check there)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    keep = [np.ascontiguousarray(w[k][sel]) for k in ("W", "W_e", "lbu", "ubu")]
    ptrs = (_dp * 6)(*[a.ctypes.data_as(_dp) for a in keep], None, None)
    L.cfemu_set_per_inst.argtypes = [ctypes.POINTER(_dp)]
    L.cfemu_set_per_inst(ptrs)
    try:
        r = emu_general(ws, 50)
    finally:
        L.cfemu_set_per_inst(None)
    nf = len(fired)
    assert nf >= 3
    assert ((r["flags"][:nf] & 3) != 0).all(), r["flags"][:nf]
    assert (r["flags"][nf:] == 0).all() and (r["status"][nf:] == 0).all() and (r["qp_status"][nf:] == 0).all()
    assert rel_err(r["u"][nf:], edge["adv_u"][calm]) < 1e-6
    assert rel_err(r["x"][nf:][:, XSEL], edge["adv_xsel"][calm]) < 1e-6


def test_emulated_multiplier_output_matches_port(gold, port):
    """pi / lam / t of the iterate and the restored multipliers of x_0 = x0 (ocp_nlp_out_get "pi"/"lam"/"t",
    x_ocp_qp_red.c:820-840) from the kernel source against the port (itself pinned to the reference: test_oracle_golden)."""
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                  _ip, _ip, _ip, _ip, _dp, _dp, ctypes.c_int]
    L.cfemu_set_mult.argtypes = [_dp]
    L.cfemu_mult_stride.restype = ctypes.c_long
    N, B = 50, 3
    w = batch(gold, "helix", B)
    stride = L.cfemu_mult_stride(N)
    mult = np.full((B, stride), np.nan)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
    res = np.zeros((B, 4))
    P = lambda a: a.ctypes.data_as(_dp)
    I = lambda a: a.ctypes.data_as(_ip)
    L.cfemu_set_mult(P(mult))
    try:
        L.cfemu_rti_batch(B, N, TS, None, 0, P(w["x0"]), P(w["yref"]), P(w["yref_e"]), P(x), P(u), I(st), I(it), I(qs), I(fl), P(res), None, 2)
    finally:
        L.cfemu_set_mult(None)
    port.record_multipliers(N)
    try:
        for i in range(B):
            xo, uo = w["x_init"][i].copy(), w["u_init"][i].copy()
            port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], xo, uo)
            pi, l0, l, t0, t = port.multipliers()
            m = mult[i]
            k_pi, k_lam, k_t = m[:N * 13].reshape(N, 13), m[N * 13:N * 21].reshape(N, 2, 4), m[N * 21:N * 29].reshape(N, 2, 4)
            k_l0 = m[N * 29:N * 29 + 13]
            sc = 1 + np.abs(pi).max()
            assert np.abs(k_pi - pi).max() / sc < 1e-9
            assert np.abs(k_lam[1:] - l).max() < 1e-9 * (1 + np.abs(l).max()) and np.abs(k_t[1:] - t).max() < 1e-9 * 23
            assert np.abs(k_lam[0] - l0[:, :4]).max() < 1e-9 * (1 + np.abs(l0).max()) and np.abs(k_t[0] - t0[:, :4]).max() < 1e-9 * 23
            signed = np.where(l0[0, 4:] > 1e-16, l0[0, 4:], -l0[1, 4:])
            assert np.abs(k_l0 - signed).max() < 1e-9 * (1 + np.abs(signed).max())
    finally:
        port.record_multipliers(0)


def test_emulated_iterative_refinement(gold):
    """lin_res_check = 2: where the corrector step fails the reference's linear-residual test it is refined as HPIPM does
    (x_ocp_qp_ipm.c:2275-2366: residual of the linear system -> right-hand side -> correction, at most two rounds).  The
    test mode 4 makes every corrector solve 10 % inaccurate in du: refinement must run (CF_FLAG_ITREF), bring the residual
    back under the tolerances within two rounds (no CF_FLAG_ITREF_LEFT) and the interior-point method must still converge
    to the same solution.  On a healthy solve mode 2 never fires and changes nothing."""
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_general.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                    _ip, _ip, _ip, _ip, _dp, ctypes.c_int]
    L.cfemu_set_lin_res_check.argtypes = [ctypes.c_int]
    w = batch(gold, "helix", 3)
    B, N = 3, 50
    P = lambda a: a.ctypes.data_as(_dp)
    I = lambda a: a.ctypes.data_as(_ip)
    out = {}
    try:
        for mode in (1, 2, 4):
            L.cfemu_set_lin_res_check(mode)
            x, u = w["x_init"].copy(), w["u_init"].copy()
            st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
            res = np.zeros((B, 4))
            # the general kernel variant, fused and as two phases (refinement is compiled into the general variants only)
            L.cfemu_rti_general(B, N, TS, None, int(mode == 4), P(w["x0"]), None, P(w["yref"]), P(w["yref_e"]), P(x), P(u), I(st), I(it), I(qs), I(fl), P(res), 2)
            out[mode] = dict(x=x, u=u, st=st, it=it, qs=qs, fl=fl)
    finally:
        L.cfemu_set_lin_res_check(1)
    assert np.array_equal(out[2]["x"], out[1]["x"]) and np.array_equal(out[2]["u"], out[1]["u"]) and (out[2]["fl"] == 0).all()
    f4 = out[4]["fl"]
    assert ((f4 & 32) != 0).all() and ((f4 & 64) == 0).all() and (out[4]["st"] == 0).all() and (out[4]["qs"] == 0).all()
    assert rel_err(out[4]["x"], out[1]["x"]) < 1e-3 and rel_err(out[4]["u"], out[1]["u"]) < 1e-3
