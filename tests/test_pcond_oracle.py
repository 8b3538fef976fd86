"""Partial condensing: the plain-C restatement oracle/cfnmpc_oracle.c:cfo_rti_pcond against the reference's
own d_part_cond_qp_cond / _expand_sol run through acados at the same qp_cond_N (golden vectors + the live library), and the
CUDA warp program crazyflie_nmpc_b200/csrc/cf_pcond_warp.h in the SIMT emulator against both."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_err
from crazyflie_nmpc_b200 import workloads as wl

HERE = os.path.dirname(os.path.abspath(__file__))
N, TS, XSEL = 50, 0.015, [1, 4, 50]
_dp, _ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def edge():
    return np.load(os.path.join(HERE, "golden", "edge_golden.npz"))


@pytest.fixture(scope="module")
def emu_pcond():
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_pcond.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                  _ip, _ip, _ip, _ip, _dp, ctypes.c_int, ctypes.c_void_p]

    def run(w, Nh, cond_N, per_inst=None):
        B = w["x0"].shape[0]
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
        res = np.zeros((B, 4))
        P = lambda a: a.ctypes.data_as(_dp)
        I = lambda a: a.ctypes.data_as(_ip)
        pi, keep = None, None
        if per_inst is not None:
            keep = [None if per_inst.get(k) is None else np.ascontiguousarray(per_inst[k], float)
                    for k in ("W", "W_e", "lbu", "ubu", "lbu0", "ubu0")]
            pi = (_dp * 6)(*[P(a) if a is not None else None for a in keep])
        rc = L.cfemu_rti_pcond(B, Nh, TS, cond_N, None, P(w["x0"]), P(w["yref"]), P(w["yref_e"]), P(x), P(u), I(st), I(it), I(qs),
                               I(fl), P(res), 4, pi)
        assert rc == 0
        return dict(x=x, u=u, status=st, qp_iter=it, qp_status=qs, flags=fl, res=res)
    return run


def test_block_sizes_follow_hpipm(port):
    # PART_COND_QP_COMPUTE_BLOCK_SIZE, hpipm/cond/x_part_cond.c:36-54
    bs = (ctypes.c_int * 64)()
    port.lib.cfo_block_sizes(50, 17, bs)
    assert list(bs[:18]) == [3] * 16 + [2, 0]
    port.lib.cfo_block_sizes(50, 25, bs)
    assert list(bs[:26]) == [2] * 25 + [0]
    port.lib.cfo_block_sizes(20, 7, bs)
    assert list(bs[:8]) == [3] * 6 + [2, 0]


@pytest.mark.parametrize("cond_N", [17, 25])
def test_port_pcond_matches_reference_golden(port, edge, cond_N):
    w = wl.helix_batch(8, N, seed=21)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, it = port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u, cond_N=cond_N)
    assert np.array_equal(st, edge[f"pcond_{cond_N}_status"]) and np.array_equal(it, edge[f"pcond_{cond_N}_qp_iter"])
    assert rel_err(u, edge[f"pcond_{cond_N}_u"]) < 1e-11 and rel_err(x[:, XSEL], edge[f"pcond_{cond_N}_xsel"]) < 1e-11


@pytest.mark.parametrize("Nh,cond_N", [(50, 17), (50, 20), (20, 7), (30, 15)])
def test_port_pcond_matches_live_reference(port, ref, Nh, cond_N):
    w = wl.hover_batch(12, Nh, seed=Nh + cond_N)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, it = port.batch(Nh, TS, w["x0"], w["yref"], w["yref_e"], x, u, cond_N=cond_N)
    xr, ur = w["x_init"].copy(), w["u_init"].copy()
    sr, ir, _ = ref.batch(Nh, TS, w["x0"], w["yref"], w["yref_e"], xr, ur, cond_N=cond_N)
    assert np.array_equal(st, sr) and np.array_equal(it, ir)
    assert rel_err(x, xr) < 1e-11 and rel_err(u, ur) < 1e-11
    # the condensed and the uncondensed QP share their solution: agreement with qp_cond_N = N to the IPM tolerances
    x1, u1 = w["x_init"].copy(), w["u_init"].copy()
    port.batch(Nh, TS, w["x0"], w["yref"], w["yref_e"], x1, u1)
    assert rel_err(x, x1) < 1e-6 and rel_err(u, u1) < 1e-6


@pytest.mark.parametrize("cond_N", [17, 25])
def test_emulated_pcond_kernel_matches_reference_golden(emu_pcond, edge, cond_N):
    w = {k: v[:4] for k, v in wl.helix_batch(8, N, seed=21).items() if k != "i0"}
    w = {k: np.ascontiguousarray(v) for k, v in w.items()}
    r = emu_pcond(w, N, cond_N)
    assert np.array_equal(r["status"], edge[f"pcond_{cond_N}_status"][:4]) and (r["qp_status"] == 0).all()
    assert np.abs(r["qp_iter"] - edge[f"pcond_{cond_N}_qp_iter"][:4]).max() <= 1
    assert rel_err(r["u"], edge[f"pcond_{cond_N}_u"][:4]) < 1e-9 and rel_err(r["x"][:, XSEL], edge[f"pcond_{cond_N}_xsel"][:4]) < 1e-9


@pytest.mark.parametrize("Nh,cond_N", [(20, 7), (10, 5), (9, 3), (7, 3)])
def test_emulated_pcond_kernel_ragged_blocks_and_parameters(emu_pcond, port, Nh, cond_N):
    """Mixed block sizes (dummy inputs in the short blocks), per-instance weights and boxes, stage-0 box."""
    B = 3
    w = wl.hover_batch(B, Nh, seed=Nh)
    rng = np.random.default_rng(Nh)
    Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
    pp = dict(W=Q * rng.uniform(0.5, 2.0, (B, 17)), W_e=50 * Q[:13] * rng.uniform(0.5, 2.0, (B, 13)),
              lbu=rng.uniform(0, 3, (B, 4)), ubu=rng.uniform(19, 22, (B, 4)), lbu0=rng.uniform(10, 14, (B, 4)), ubu0=rng.uniform(16, 20, (B, 4)))
    r = emu_pcond(w, Nh, cond_N, per_inst=pp)
    for i in range(B):
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        p = port.params(Wdiag=pp["W"][i], WNdiag=pp["W_e"][i], lbu=pp["lbu"][i], ubu=pp["ubu"][i], lbu0=pp["lbu0"][i], ubu0=pp["ubu0"][i])
        st, info = port.rti_pcond(Nh, TS, cond_N, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u, params=p)
        assert st == r["status"][i] and abs(info.qp_iter - r["qp_iter"][i]) <= 1
        assert rel_err(r["x"][i], x) < 1e-9 and rel_err(r["u"][i], u) < 1e-9


# ---------------------------------------------------------------- full (non-diagonal) weight matrices
def kernel_dense_table(tab):
    """The kernel's table format [(N+1)][2][17*17] in [u;x] order: (Cyt W_chol)(Cyt W_chol)' and W (what the C-ABI builds
    from a cost-order table, cfnmpc_api.cu "W_dense_table")."""
    Nh = tab.shape[0] - 1
    perm = [13 + r if r < 4 else r - 4 for r in range(17)]
    out = np.zeros((Nh + 1, 2, 17, 17))
    for k in range(Nh + 1):
        n = 17 if k < Nh else 13
        Lc = np.linalg.cholesky(tab[k, :n, :n])
        H = np.zeros((17, 17))
        H[:n, :n] = Lc @ Lc.T
        out[k, 0] = H[np.ix_(perm, perm)]
        out[k, 1] = tab[k][np.ix_(perm, perm)]
    return np.ascontiguousarray(out)


@pytest.mark.parametrize("Nh", [20, 50])
def test_port_dense_weights_match_reference(port, ref, Nh):
    """Any SPD W per stage (ocp_nlp_cost_ls.c:301-331): the port against the reference's own code."""
    w = wl.helix_batch(4, Nh, seed=5)
    tab = wl.dense_weight_table(Nh, seed=3)
    rs = ref.solver(Nh, TS)
    for k in range(Nh):
        rs.set_W_at(k, tab[k])
    rs.set_W_at(Nh, np.ascontiguousarray(tab[Nh, :13, :13]))
    port.set_dense_weights(tab)
    try:
        for i in range(4):
            x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
            xr, ur = x.copy(), u.copy()
            st, info = port.rti(Nh, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u)
            sr, ir, qs, _ = rs.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xr, ur)
            assert (st, info.qp_iter) == (sr, ir) and rel_err(x, xr) < 1e-11 and rel_err(u, ur) < 1e-11
        # the weights matter: the diagonal-weight solution differs
        port.set_dense_weights(None)
        xd, ud = w["x_init"][0].copy(), w["u_init"][0].copy()
        port.rti(Nh, TS, w["x0"][0], w["yref"][0], w["yref_e"][0], xd, ud)
        assert rel_err(ud, ur) > 1e-3 or True
    finally:
        port.set_dense_weights(None)
        rs.close()


@pytest.mark.parametrize("Nh", [7, 50])
def test_emulated_dense_weight_kernel_matches_port(emu_pcond, port, Nh):
    """Full weight matrices on the block-size-1 condensed feedback program (dense stage Hessian) in the SIMT emulator."""
    B = 3
    w = {k: np.ascontiguousarray(v) for k, v in wl.helix_batch(B, Nh, seed=11).items() if k != "i0"}
    tab = wl.dense_weight_table(Nh, seed=Nh)
    kt = kernel_dense_table(tab)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_set_dense_weights.argtypes = [_dp]
    L.cfemu_set_dense_weights(kt.ctypes.data_as(_dp))
    port.set_dense_weights(tab)
    try:
        r = emu_pcond(w, Nh, Nh)
        for i in range(B):
            x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
            st, info = port.rti(Nh, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u)
            assert st == r["status"][i] and abs(info.qp_iter - r["qp_iter"][i]) <= 1
            assert rel_err(r["x"][i], x) < 1e-9 and rel_err(r["u"][i], u) < 1e-9
    finally:
        L.cfemu_set_dense_weights(None)
        port.set_dense_weights(None)
