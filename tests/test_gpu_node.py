"""SURVEY 8f-2 / INTEGRATION.md section A on the GPU: (1) the reference's ROS node class, compiled UNMODIFIED against
include/ and linked to libcfnmpc.so (tests/dropin/build_node.py -> tests/dropin/_build/node_ours, a prebuilt binary that
travels to the GPU box because the node source does not), publishes what the same class published on top of the
reference's own acados build (tests/golden/node_loop_golden.npz); (2) the device-side tick cfnmpc_batch_tick builds the
same reference windows bit-for-bit and the same commands."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf
from oracle import loop_oracle as lo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
N, TS = 50, 0.015


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "node_loop_golden.npz"))


def check_published(out, gold):
    pub = gold["published"]
    # int32 motor speeds: equal unless the reference's u0 sits within the parity tolerance of an integer
    u0 = gold["solver_u0"]
    near = np.abs(u0 - np.round(u0)) < 1e-5
    assert (out[:, :4] == pub[:, :4])[~near].all(), "motor commands differ from the reference node's"
    assert np.abs(out[:, :4] - pub[:, :4]).max() <= 1
    # pitch / roll set-points and yaw rate in degrees: functions of x4; stated solver tolerance 1e-6 (1 + |ref|)
    for c in (4, 5, 7):
        assert np.abs(out[:, c] - pub[:, c]).max() <= 1e-6 * (1 + np.abs(pub[:, c]).max()) * 57.3, c
    assert np.abs(out[:, 6] - pub[:, 6]).max() <= 1.0    # thrust PWM is truncated to an integer: one count at most
    assert (out[:, 6] == pub[:, 6]).mean() > 0.9


def test_unmodified_node_on_our_library(gold):
    exe = os.path.join(ROOT, "tests", "dropin", "_build", "node_ours")
    import build_node_path  # noqa: F401  (adds tests/dropin to sys.path)
    import build_node
    build_node.build()      # no-op on the GPU box (no /root/reference): the prebuilt binary is used
    if not os.path.exists(exe):
        pytest.skip("tests/dropin/_build/node_ours was not built (needs /root/reference at build time)")
    import make_node_golden as mg
    out, _ = mg.run_node(exe, gold["scenario"], gold["table"], False)
    check_published(out, gold)


def test_batch_tick_builds_the_nodes_windows_and_commands(gold):
    sc, T = gold["scenario"], gold["table"]
    B = 3   # identical vehicles: every instance must reproduce the node
    with cf.BatchSolver(B, N, TS) as s:
        s.set_trajectory(T)
        s.set("traj_iter", np.zeros(B, np.int32))
        out = np.zeros((sc.shape[0], 8))
        for t in range(sc.shape[0]):
            if sc[t, 0] == 1:
                s.set("policy", np.full(B, cf.POLICY_TRACKING, np.int32))
            elif sc[t, 0] == 0:
                s.set("policy", np.full(B, cf.POLICY_REGULATION, np.int32)).set("setpoint", np.tile(sc[t, 1:4], (B, 1)))
            s.set("x0", np.tile(sc[t, 4:], (B, 1)))
            s.tick()
            y, ye = s.get("yref"), s.get("yref_e")
            for i in range(B):
                assert np.array_equal(y[i].reshape(-1), gold["solver_yref"][t]), (t, i)
                assert np.array_equal(ye[i], gold["solver_yref_e"][t]), (t, i)
            m, tw, st = s.get("motors"), s.get("twist"), s.get("status")
            assert (st == int(gold["solver_status"][t, 0])).all()
            assert (m == m[0]).all() and (tw == tw[0]).all()
            out[t, :4], out[t, 4:] = m[0], tw[0]
            for a, b in ((s.get("u", 0)[0], gold["solver_u0"][t]), (s.get("u", 1)[0], gold["solver_u1"][t]),
                         (s.get("x", 4)[0], gold["solver_x4"][t])):
                assert np.abs(a - b).max() <= 1e-6 * (1 + np.abs(b).max()), (t, a, b)
        assert (s.get("policy") == lo.HOLD).all()
    check_published(out, gold)
