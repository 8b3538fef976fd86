"""Partial condensing on the GPU (crazyflie_nmpc_b200/csrc/cf_pcond_warp.h): option "qp_cond_N" = the
reference's qp_cond_N (ocp_qp_partial_condensing.c:457-576 -> hpipm/cond/x_part_cond.c:505-560,658-742).  Parity is
against the reference run at the SAME qp_cond_N (oracle/_ref where it travelled) and the plain-C oracle, which is pinned
to it (tests/test_pcond_oracle.py)."""
import os

import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl
from conftest import rel_err

pytestmark = pytest.mark.gpu
TS, TOL, TIGHT = 0.015, 1e-6, 1e-9


def gpu(w, N, cond_N, n_rti=1, mode="solve", **fields):
    B = w["x0"].shape[0]
    with cf.BatchSolver(B, N, TS) as s:
        s.set_option("qp_cond_N", cond_N)
        for k, v in fields.items():
            s.set(k, v)
        if mode == "host":
            s.set("x", w["x_init"]).set("u", w["u_init"])
            s.solve_from_host(w["x0"], w["yref"], w["yref_e"], n_chunks=3)
        elif mode == "split":
            s.set_problem(w).prepare().feedback()
        else:
            s.set_problem(w).solve(n_rti)
        return dict(x=s.get("x_all"), u=s.get("u_all"), status=s.get("status"), qp_iter=s.get("qp_iter"),
                    qp_status=s.get("qp_status"), flags=s.get("flags"), block=s.info("pcond_block_size"))


@pytest.mark.parametrize("gen", [wl.hover_batch, wl.helix_batch])
@pytest.mark.parametrize("N,cond_N", [(50, 17), (50, 25), (50, 20), (20, 7), (100, 34)])
def test_pcond_matches_oracle_and_reference(port, gen, N, cond_N):
    B = 96
    w = gen(B, N, seed=5)
    g = gpu(w, N, cond_N)
    xo, uo = w["x_init"].copy(), w["u_init"].copy()
    so, io = port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], xo, uo, cond_N=cond_N)
    assert (g["status"] == so).all() and (g["flags"] == 0).all() and (g["qp_status"] == 0).all()
    assert np.abs(g["qp_iter"] - io).max() <= 1
    assert rel_err(g["x"], xo) <= TIGHT and rel_err(g["u"], uo) <= TIGHT
    from oracle.oracle import Ref, ref_available
    if ref_available():
        xr, ur = w["x_init"].copy(), w["u_init"].copy()
        sr, ir, _ = Ref().batch(N, TS, w["x0"], w["yref"], w["yref_e"], xr, ur, nthreads=os.cpu_count() or 1, cond_N=cond_N)
        assert (g["status"] == sr).all() and np.abs(g["qp_iter"] - ir).max() <= 1
        assert rel_err(g["x"], xr) <= TIGHT and rel_err(g["u"], ur) <= TIGHT
    # the condensed and the uncondensed QP have the same solution: the stated tolerance holds against qp_cond_N = N too
    u = gpu(w, N, 0)
    assert u["block"] == 1 and g["block"] in (2, 3)
    assert rel_err(g["x"], u["x"]) <= TOL and rel_err(g["u"], u["u"]) <= TOL


def test_pcond_paths_and_consecutive_steps(port):
    N, B, cond_N = 50, 64, 17
    w = wl.helix_batch(B, N, seed=8)
    a = gpu(w, N, cond_N)
    for mode in ("host", "split"):
        b = gpu(w, N, cond_N, mode=mode)
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["u"], b["u"]) and np.array_equal(a["status"], b["status"])
    g3 = gpu(w, N, cond_N, n_rti=3)
    xo, uo = w["x_init"].copy(), w["u_init"].copy()
    port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], xo, uo, n_rti=3, cond_N=cond_N)
    assert rel_err(g3["x"], xo) <= 1e-7 and rel_err(g3["u"], uo) <= 1e-7


def test_pcond_per_instance_parameters_and_stage_bounds(port):
    N, B, cond_N = 20, 12, 7
    w = wl.hover_batch(B, N, seed=31)
    rng = np.random.default_rng(2)
    Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
    W = Q * rng.uniform(0.5, 2.0, (B, 17))
    WN = W[:, :13] * rng.uniform(20, 60, (B, 1))
    lbu, ubu = rng.uniform(0.0, 3.0, (B, 4)), rng.uniform(19.0, 22.0, (B, 4))
    g = gpu(w, N, cond_N, W_batch=W, W_e_batch=WN, lbu_batch=lbu, ubu_batch=ubu)
    for i in range(B):
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        p = port.params(Wdiag=W[i], WNdiag=WN[i], lbu=lbu[i], ubu=ubu[i])
        st, info = port.rti_pcond(N, TS, cond_N, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u, params=p)
        assert st == g["status"][i] and abs(info.qp_iter - g["qp_iter"][i]) <= 1
        assert rel_err(g["x"][i], x) <= TIGHT and rel_err(g["u"][i], u) <= TIGHT
    # input box per stage
    tab = np.tile(np.r_[np.zeros(4), np.full(4, 22.0)], (N, 1))
    tab[0, :4], tab[0, 4:] = 12.0, 17.0
    tab[3, 4:] = 16.5
    port.set_stage_bounds(tab)
    try:
        xo, uo = w["x_init"].copy(), w["u_init"].copy()
        port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], xo, uo, cond_N=cond_N)
    finally:
        port.set_stage_bounds(None)
    g = gpu(w, N, cond_N, bounds_stage=tab)
    assert rel_err(g["x"], xo) <= TIGHT and rel_err(g["u"], uo) <= TIGHT
    assert g["u"][:, 3].max() <= 16.5 + 1e-6 and g["u"][:, 0].min() >= 12.0 - 1e-6


def test_pcond_option_errors_and_switching():
    with cf.BatchSolver(4, 50, TS) as s:
        with pytest.raises(cf.CfnmpcError):
            s.set_option("qp_cond_N", 5)        # blocks of 10 stages: not implemented
        s.set_option("qp_cond_N", 17)
        assert s.info("qp_cond_N") == 17 and s.info("pcond_block_size") == 3
        s.set_option("qp_cond_N", 50)
        assert s.info("qp_cond_N") == 50 and s.info("pcond_block_size") == 1


def test_pcond_full_size_sample(port):
    N, B, cond_N = 50, 65536, 17
    w = wl.hover_batch(B, N)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_option("qp_cond_N", cond_N)
        s.set_problem(w).solve(1)
        st, fl = s.get("status"), s.get("flags")
        idx = np.arange(0, B, B // 256)
        x, u = s.get("x_all")[idx], s.get("u_all")[idx]
    assert (st == 0).all() and (fl == 0).all()
    ws = {k: np.ascontiguousarray(v[idx]) for k, v in w.items()}
    xo, uo = ws["x_init"].copy(), ws["u_init"].copy()
    port.batch(N, TS, ws["x0"], ws["yref"], ws["yref_e"], xo, uo, cond_N=cond_N)
    assert rel_err(x, xo) <= TIGHT and rel_err(u, uo) <= TIGHT


def test_full_weight_matrices(port):
    """"W_dense_table": any SPD weight matrix per stage (ocp_nlp_cost_ls.c:301-331) on the dense-Hessian kernel, against the
    port (pinned to the reference with per-stage ocp_nlp_cost_model_set calls: tests/test_pcond_oracle.py); a diagonal table
    reproduces the diagonal-weight kernels; clearing the table returns to them."""
    N, B = 50, 24
    w = wl.helix_batch(B, N, seed=31)
    tab = wl.dense_weight_table(N, seed=8)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).solve(1)
        x_diag, u_diag = s.get("x_all"), s.get("u_all")
        s.set("W_dense_table", tab)
        s.set_problem(w).solve(1)
        x, u, st, it = s.get("x_all"), s.get("u_all"), s.get("status"), s.get("qp_iter")
        with pytest.raises(cf.CfnmpcError):
            s.set_option("qp_cond_N", 25)
        with pytest.raises(cf.CfnmpcError):
            s.set_option("multipliers", 1)
        # the default weights written as full matrices: same problem as the diagonal kernels solve
        Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
        dt = np.zeros((N + 1, 17, 17))
        dt[:N] = np.diag(Q)
        dt[N, :13, :13] = np.diag(50 * Q[:13])
        s.set("W_dense_table", dt)
        s.set_problem(w).solve(1)
        assert rel_err(s.get("x_all"), x_diag) < 1e-9 and rel_err(s.get("u_all"), u_diag) < 1e-9
        s.clear("W_dense_table")
        s.set_problem(w).solve(1)
        assert np.array_equal(s.get("x_all"), x_diag) and np.array_equal(s.get("u_all"), u_diag)
        bad = tab.copy()
        bad[3, 0, 0] = -1.0
        with pytest.raises(cf.CfnmpcError):
            s.set("W_dense_table", bad)
    port.set_dense_weights(tab)
    try:
        for i in range(B):
            xo, uo = w["x_init"][i].copy(), w["u_init"][i].copy()
            so, info = port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], xo, uo)
            assert so == st[i] and abs(info.qp_iter - it[i]) <= 1
            assert rel_err(x[i], xo) < 1e-9 and rel_err(u[i], uo) < 1e-9
    finally:
        port.set_dense_weights(None)
    assert rel_err(u, u_diag) > 1e-4   # the off-diagonal weights matter


def test_capsule_full_weight_matrix_of_one_stage(ref):
    """ocp_nlp_cost_model_set(.., k, "W", ..) with a non-diagonal matrix on the capsule surface, against the reference."""
    import ctypes
    L = cf.lib()
    vp = ctypes.c_void_p
    L.crazyflie_acados_create_capsule.restype = vp
    for f in ("crazyflie_acados_get_nlp_in", "crazyflie_acados_get_nlp_out", "crazyflie_acados_get_nlp_config", "crazyflie_acados_get_nlp_dims"):
        getattr(L, f).restype = vp
        getattr(L, f).argtypes = [vp]
    L.crazyflie_acados_create.argtypes = [vp]
    L.crazyflie_acados_solve.argtypes = [vp]
    L.crazyflie_acados_free.argtypes = [vp]
    L.crazyflie_acados_free_capsule.argtypes = [vp]
    L.ocp_nlp_constraints_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_cost_model_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_set.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    L.ocp_nlp_out_get.argtypes = [vp, vp, vp, ctypes.c_int, ctypes.c_char_p, vp]
    N = 50
    w = wl.helix_batch(1, N, seed=6)
    tab = wl.dense_weight_table(N, seed=2)
    stages = [0, 7, 8, N]     # only these stages get a full matrix; the others keep the default diagonal
    cap = L.crazyflie_acados_create_capsule()
    assert L.crazyflie_acados_create(cap) == 0
    cfg, dims, nin, nout = (L.crazyflie_acados_get_nlp_config(cap), L.crazyflie_acados_get_nlp_dims(cap),
                            L.crazyflie_acados_get_nlp_in(cap), L.crazyflie_acados_get_nlp_out(cap))
    P = lambda a: ctypes.c_void_p(a.ctypes.data)
    rs = ref.solver(N, TS)
    for k in stages:
        Wk = np.ascontiguousarray(tab[k] if k < N else tab[N, :13, :13])
        assert L.ocp_nlp_cost_model_set(cfg, dims, nin, k, b"W", P(Wk)) == 0
        rs.set_W_at(k, Wk)
    L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"lbx", P(w["x0"][0]))
    L.ocp_nlp_constraints_model_set(cfg, dims, nin, 0, b"ubx", P(w["x0"][0]))
    for k in range(N):
        L.ocp_nlp_cost_model_set(cfg, dims, nin, k, b"yref", P(w["yref"][0, k]))
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"u", P(w["u_init"][0, k]))
    L.ocp_nlp_cost_model_set(cfg, dims, nin, N, b"yref", P(w["yref_e"][0]))
    for k in range(N + 1):
        L.ocp_nlp_out_set(cfg, dims, nout, k, b"x", P(w["x_init"][0, k]))
    assert L.crazyflie_acados_solve(cap) == 0
    x, u = np.zeros((N + 1, 13)), np.zeros((N, 4))
    for k in range(N + 1):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"x", P(x[k]))
    for k in range(N):
        L.ocp_nlp_out_get(cfg, dims, nout, k, b"u", P(u[k]))
    xr, ur = w["x_init"][0].copy(), w["u_init"][0].copy()
    sr = rs.rti(w["x0"][0], w["yref"][0], w["yref_e"][0], xr, ur)
    rs.close()
    assert sr[0] == 0 and rel_err(x, xr) < 1e-9 and rel_err(u, ur) < 1e-9
    L.crazyflie_acados_free(cap)
    L.crazyflie_acados_free_capsule(cap)
