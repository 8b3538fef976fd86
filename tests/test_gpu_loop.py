"""GPU parity of the kernels either side of the RTI step with their oracles:
state predictor vs the port of sim_erk (itself pinned to the reference's integrator in test_loop_oracle.py),
reference-window / command kernels vs the numpy restatement of the node's per-tick logic, and a short
closed-loop run entirely on the device vs the same loop assembled from the CPU oracles."""
import os
import subprocess

import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl
from conftest import rel_err
from oracle import loop_oracle as lo

pytestmark = pytest.mark.gpu
TS = 0.015
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_states(B, seed):
    rng = np.random.default_rng(seed)
    w = wl.hover_batch(B, 1, seed=seed)
    return w["x0"], rng.uniform(0.0, 22.0, (B, 4)), rng.uniform(0.004, 0.06, B)


@pytest.mark.parametrize("B,steps", [(1, 1), (130, 1), (1000, 3)])
def test_predictor_matches_oracle(port, B, steps):
    x, u, T = random_states(B, 100 + B)
    with cf.SimBatch(B, num_steps=steps) as s:
        xn = s.set("x", x).set("u", u).set("T_batch", T).solve().get("xn")
        ref = np.array([port.sim(x[i], u[i], T[i], steps) for i in range(B)])
        assert rel_err(xn, ref) <= 1e-13        # fp64, fused multiply-adds are the only difference
        # one horizon for every instance (the estimator's `delay`)
        xn2 = s.set("T", [0.03]).solve().get("xn")
        ref2 = np.array([port.sim(x[i], u[i], 0.03, steps) for i in range(B)])
        assert rel_err(xn2, ref2) <= 1e-13
        assert s.launches() == 2


def test_predictor_forward_sensitivities(port):
    B = 77
    x, u, T = random_states(B, 5)
    with cf.SimBatch(B, sens_forw=True) as s:
        s.set("x", x).set("u", u).set("T_batch", T).solve()
        xn, S = s.get("xn"), s.get("S_forw")       # S[b][column][row], columns [x | u]
    for i in range(B):
        xr, A, Bm = port.erk4(x[i], u[i], T[i])
        assert rel_err(xn[i], xr) <= 1e-13
        assert np.abs(S[i, :13].T - A).max() <= 1e-12 and np.abs(S[i, 13:].T - Bm).max() <= 1e-12


def test_predictor_errors():
    with cf.SimBatch(4) as s:
        with pytest.raises(cf.CfnmpcError):
            s.get("S_forw")                         # sens_forw is off
        with pytest.raises(cf.CfnmpcError):
            s.opts_set("num_stages", 3)
        with pytest.raises(cf.CfnmpcError):
            s.set("x", np.zeros(5))


def mixed_policies(B, N, rows, seed=2):
    rng = np.random.default_rng(seed)
    pol = rng.integers(0, 3, B).astype(np.int32)
    it = rng.integers(0, rows - N - 1, B).astype(np.int32)
    it[:6] = [rows - N - 2, rows - N - 1, rows - N, rows - N + 5, 0, 1]     # around the end of the table
    pol[:6] = lo.TRACKING
    sp = np.c_[rng.uniform(-1, 1, (B, 2)), rng.uniform(0, 1, B)]
    return pol, it, sp


def test_reference_window_policies():
    N, B = 50, 96
    T = wl.helix_table()
    pol, it, sp = mixed_policies(B, N, T.shape[0])
    w = wl.hover_batch(B, N, seed=8)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).set_trajectory(T).set("policy", pol).set("traj_iter", it).set("setpoint", sp)
        y, ye = w["yref"].copy(), w["yref_e"].copy()
        p, t = pol.copy(), it.copy()
        for tick in range(3):
            s.update_reference()
            gy, gye, gp, gt = s.get("yref"), s.get("yref_e"), s.get("policy"), s.get("traj_iter")
            for i in range(B):
                y[i], ye[i], p[i], t[i] = lo.reference_window(p[i], t[i], sp[i], T, N, lo.node_uss(), y[i], ye[i])
            assert np.array_equal(gy, y) and np.array_equal(gye, ye)      # copies and constants: bit-exact
            assert np.array_equal(gp, p) and np.array_equal(gt, t)
        assert (p[:4] == lo.HOLD).all() or (p[2:4] == lo.HOLD).all()
        # a caller-chosen hover speed
        s.set("uss", [15.5]).set("policy", np.zeros(B, np.int32)).update_reference()
        assert (s.get("yref")[:, :, 13:] == 15.5).all()


def test_commands_match_node_logic():
    N, B = 50, 200
    w = wl.hover_batch(B, N, seed=12)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).solve(2).commands()
        x, u = s.get("x_all"), s.get("u_all")
        m, eu, tw = s.get("motors"), s.get("euler"), s.get("twist")
        s.commands(motors_from_u1=True)
        m1 = s.get("motors")
    for i in range(B):
        mo, euo, two = lo.commands(u[i, 0], u[i, 1], x[i, 4])
        assert np.array_equal(m[i], mo)                                  # same doubles -> same truncation
        assert np.abs(eu[i] - euo).max() <= 1e-14 and np.abs(tw[i] - two).max() <= 1e-11
        assert np.array_equal(m1[i], lo.commands(u[i, 0], u[i, 1], x[i, 4], fixed_u0=True)[0])
    assert m.min() >= 0 and m.max() <= 22


def test_closed_loop_on_device_matches_cpu_loop(port):
    """5 control ticks of tracking vehicles, plant simulated on the device, vs the same loop from the CPU oracles."""
    N, B, ticks = 50, 24, 5
    T = wl.helix_table()
    w = wl.helix_batch(B, N, seed=77)
    rng = np.random.default_rng(1)
    it0 = rng.integers(0, 900, B).astype(np.int32)
    pol0 = np.full(B, lo.TRACKING, np.int32)
    pol0[-4:] = lo.REGULATION
    sp = np.tile([0.0, 0.0, 0.4], (B, 1))
    x0 = w["x0"].copy()
    for i in range(B):   # start near the reference the window will hold
        x0[i, :3] = (T[it0[i], :3] if pol0[i] == lo.TRACKING else sp[i]) + rng.uniform(-0.05, 0.05, 3)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).set("x0", x0).set_trajectory(T).set("policy", pol0).set("traj_iter", it0).set("setpoint", sp)
        hist = []
        for t in range(ticks):
            s.tick()
            hist.append((s.get("u", 0), s.get("x", 4), s.get("twist"), s.get("status")))
            s.plant_step(TS)
            hist[-1] += (s.get("x0"),)
        launches = s.info("launches")
    assert launches >= ticks * 6
    # CPU loop
    y, ye = w["yref"].copy(), w["yref_e"].copy()
    x, u = w["x_init"].copy(), w["u_init"].copy()
    p, itr, xc = pol0.copy(), it0.copy(), x0.copy()
    for t in range(ticks):
        for i in range(B):
            y[i], ye[i], p[i], itr[i] = lo.reference_window(p[i], itr[i], sp[i], T, N, lo.node_uss(), y[i], ye[i])
            st, _ = port.rti(N, TS, xc[i], y[i], ye[i], x[i], u[i])
            assert st == hist[t][3][i]
            xc[i] = port.sim(xc[i], u[i, 0], TS, 1)
        assert rel_err(hist[t][0], u[:, 0]) <= 1e-7 and rel_err(hist[t][1], x[:, 4]) <= 1e-7
        assert rel_err(hist[t][4], xc) <= 1e-7
        tw = np.array([lo.commands(u[i, 0], u[i, 1], x[i, 4])[2] for i in range(B)])
        assert np.abs(hist[t][2][:, [0, 1, 3]] - tw[:, [0, 1, 3]]).max() <= 1e-5
        assert np.abs(hist[t][2][:, 2] - tw[:, 2]).max() <= 1.0           # PWM is an integer: allow one count


def test_plant_step_with_truncated_motors(port):
    N, B = 20, 40
    w = wl.hover_batch(B, N, seed=4)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).solve(1).commands()
        m = s.get("motors")
        s.plant_step(0.03, n_steps=2, truncated_motors=True)
        xn = s.get("x0")
    ref = np.array([port.sim(w["x0"][i], m[i].astype(float), 0.03, 2) for i in range(B)])
    assert rel_err(xn, ref) <= 1e-13


def test_estimator_sequence_dropin(port):
    """The estimator node's call sequence compiled against the drop-in headers (tests/dropin/estimator_sequence.cpp)."""
    exe = os.path.join(ROOT, "tests", "dropin", "estimator_sequence")
    cf.lib()
    subprocess.run(["g++", "-O1", "-I", os.path.join(ROOT, "include"), exe + ".cpp", "-o", exe,
                    "-L", os.path.join(ROOT, "crazyflie_nmpc_b200"), "-lcfnmpc",
                    "-Wl,-rpath," + os.path.join(ROOT, "crazyflie_nmpc_b200")], check=True)
    r = subprocess.run([exe, "4"], capture_output=True, text=True, check=True)
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("pred")]
    assert len(rows) == 4
    x = np.array([0.1, -0.05, 0.3, 0.995, 0.05, -0.03, 0.02, 0.1, 0.0, -0.1, 0.2, -0.1, 0.3])
    u = np.array([15.0, 16, 14, 17])
    for t, row in enumerate(rows):
        assert int(row[3]) == 0
        x = port.sim(x, u, 0.015 + 0.005 * t, 1)
        assert np.abs(np.array(row[6:], float) - x).max() <= 1e-13
