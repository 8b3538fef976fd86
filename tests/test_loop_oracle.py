"""CPU checks of the closed-loop restatements (oracle/loop_oracle.py) and of the predictor oracle against the
reference's own integrator (oracle/_ref: cfref_sim_*)."""
import ctypes
import os

import numpy as np
import pytest

from crazyflie_nmpc_b200 import workloads as wl
from oracle import loop_oracle as lo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = ctypes.POINTER(ctypes.c_double)


def test_node_uss_and_regulation_window():
    assert lo.node_uss() == wl.node_hover_speed()
    N = 50
    y, ye, pol, it = lo.reference_window(lo.REGULATION, 0, [0.1, -0.2, 0.4], None, N, lo.node_uss(), np.zeros((N, 17)), np.zeros(13))
    w = wl.single_hover(N, template_iterate=True, node_yref=True)
    assert pol == lo.REGULATION and it == 0
    assert np.array_equal(y[:, 3:], w["yref"][0][:, 3:]) and np.array_equal(y[0, :3], [0.1, -0.2, 0.4])
    assert np.array_equal(ye[3:], w["yref_e"][0][3:])


def test_tracking_window_end_of_table_and_hold():
    N, T = 50, wl.helix_table()
    rows = T.shape[0]
    y0, ye0 = np.full((N, 17), 7.0), np.full(13, 7.0)
    y, ye, pol, it = lo.reference_window(lo.TRACKING, 10, None, T, N, 1.0, y0, ye0)
    assert it == 11 and pol == lo.TRACKING
    assert np.array_equal(y, T[10:60]) and np.array_equal(ye, T[60, :13])
    # same indexing as the synthetic tracking workload (acados_mpc.cpp:463-482)
    it_last = rows - N - 1
    y, ye, pol, it = lo.reference_window(lo.TRACKING, it_last, None, T, N, 1.0, y0, ye0)
    assert it == rows - N and pol == lo.TRACKING and np.array_equal(ye, T[rows - 1, :13])
    # past the end: nothing is written in the switching tick, then the hold reference
    y2, ye2, pol, it = lo.reference_window(lo.TRACKING, it, None, T, N, 1.0, y, ye)
    assert pol == lo.HOLD and it == rows - N and np.array_equal(y2, y) and np.array_equal(ye2, ye)
    y3, ye3, pol, it = lo.reference_window(pol, it, None, T, N, 15.5, y2, ye2)
    assert pol == lo.HOLD and np.array_equal(y3[7, :3], T[-1, :3]) and y3[7, 3] == 1.0 and (y3[:, 13:] == 15.5).all()
    assert (y3[:, 4:13] == 0).all() and np.array_equal(ye3[:3], T[-1, :3])


def test_commands_known_answers():
    u0 = np.array([15.99, 16.2, 0.4, 21.999999])
    u1 = np.array([16.0, 16.0, 16.0, 16.0])
    x4 = np.zeros(13)
    th = np.deg2rad(10.0)                      # pure pitch of 10 deg about y, quaternion scaled by 2 (node normalises)
    x4[3], x4[5] = 2 * np.cos(th / 2), 2 * np.sin(th / 2)
    x4[12] = 0.5
    m, eu, tw = lo.commands(u0, u1, x4)
    assert m.tolist() == [15, 16, 0, 21] and m.dtype == np.int32
    assert abs(eu[0]) < 1e-15 and abs(eu[1] + th) < 1e-15 and abs(eu[2]) < 1e-15   # theta = -asin(R31)
    assert abs(tw[0] + 10.0) < 1e-12 and tw[1] == 0.0
    assert tw[2] == float(int((16000 - 4070.3) / 0.2685)) and abs(tw[3] - 0.5 * 180 / np.pi) < 1e-12
    m1, _, _ = lo.commands(u0, u1, x4, fixed_u0=True)
    assert m1.tolist() == [16, 16, 16, 16]


def test_predictor_oracle_matches_reference_integrator(port, ref):
    """cfo_sim / cfo_erk4 (the port) against the reference's sim_erk driven like acados_sim_solver_crazyflie.c."""
    L = ref.lib
    L.cfref_sim_create.restype = ctypes.c_void_p
    L.cfref_sim_create.argtypes = [ctypes.c_int, ctypes.c_int]
    L.cfref_sim_solve.argtypes = [ctypes.c_void_p, _dp, _dp, ctypes.c_double, _dp, _dp]
    L.cfref_sim_destroy.argtypes = [ctypes.c_void_p]
    rng = np.random.default_rng(3)
    w = wl.hover_batch(16, 2, seed=17)
    for steps in (1, 3):
        h = L.cfref_sim_create(steps, 1)
        assert h
        for i in range(16):
            x, u, T = w["x0"][i], rng.uniform(0, 22, 4), float(rng.uniform(0.005, 0.06))
            xn, S = np.zeros(13), np.zeros(13 * 17)
            st = L.cfref_sim_solve(h, x.ctypes.data_as(_dp), u.ctypes.data_as(_dp), T, xn.ctypes.data_as(_dp), S.ctypes.data_as(_dp))
            assert st == 0
            assert np.abs(port.sim(x, u, T, steps) - xn).max() <= 1e-15 * (1 + np.abs(xn).max())
            if steps == 1:
                xn2, A, B = port.erk4(x, u, T)
                Sm = S.reshape(17, 13).T
                assert np.abs(Sm[:, :13] - A).max() < 1e-14 and np.abs(Sm[:, 13:] - B).max() < 1e-14
        L.cfref_sim_destroy(h)
