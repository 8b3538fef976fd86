"""The reference's ROS node class, compiled UNMODIFIED (tests/dropin/build_node.py) and run on the reference's own acados
build, is the ground truth for everything the node does around acados_solve() (SURVEY 8f-2,
crazyflie_controller/src/acados_mpc.cpp:430-516,619-670).  tests/golden/node_loop_golden.npz holds one run of it through
Regulation -> set-point change -> Tracking -> end of the trajectory table -> Position_Hold: what was handed to the solver
(x0, yref window) and what was published (int32 motor speeds, body twist).  These CPU tests pin the numpy restatement
oracle/loop_oracle.py and the plain-C solver oracle against it; tests/test_gpu_node.py does the same for the CUDA path."""
import os

import numpy as np
import pytest

from crazyflie_nmpc_b200 import workloads as wl
from oracle import loop_oracle as lo

HERE = os.path.dirname(os.path.abspath(__file__))
N = 50


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "node_loop_golden.npz"))


def replay_windows(gold):
    """The node's policy state machine replayed by the restatement; yields (tick, yref, yref_e, policy)."""
    sc, T = gold["scenario"], gold["table"]
    pol, it, sp = None, 0, np.zeros(3)
    y, ye = np.zeros((N, 17)), np.zeros(13)
    for t in range(sc.shape[0]):
        if sc[t, 0] == 1:
            pol = lo.TRACKING          # callback_dynamic_reconfigure, acados_mpc.cpp:305-329
        elif sc[t, 0] == 0:
            pol, sp = lo.REGULATION, sc[t, 1:4].copy()
        y, ye, pol, it = lo.reference_window(pol, it, sp, T, N, lo.node_uss(), y, ye)
        yield t, y, ye, pol


def test_node_uss_literal():
    # float(sqrt(double(0.033f) * 9.80665 / double(4.0f * 3.25e-4f)))
    assert lo.node_uss() == 15.777770042419434 == wl.node_hover_speed()


def test_reference_windows_bit_for_bit(gold):
    seen = set()
    for t, y, ye, pol in replay_windows(gold):
        assert np.array_equal(y.reshape(-1), gold["solver_yref"][t]), f"tick {t}: yref window differs from the node's"
        assert np.array_equal(ye, gold["solver_yref_e"][t]), f"tick {t}: terminal reference differs"
        seen.add(pol)
    assert seen == {lo.REGULATION, lo.TRACKING, lo.HOLD}
    # the measured state goes to the solver unchanged (lbx = ubx = x0, :581-582)
    assert np.array_equal(gold["solver_x0"], gold["scenario"][:, 4:])


def test_published_commands_bit_for_bit(gold):
    """motors = int32(u0), twist from x4 / u1 (acados_mpc.cpp:619-670) -- from the solver outputs the node itself got."""
    pub = gold["published"]
    for t in range(pub.shape[0]):
        m, eul, tw = lo.commands(gold["solver_u0"][t], gold["solver_u1"][t], gold["solver_x4"][t])
        assert np.array_equal(m, pub[t, :4].astype(np.int32)), t
        assert np.array_equal(tw, pub[t, 4:]), (t, tw, pub[t, 4:])


def test_port_oracle_reproduces_the_nodes_solves(gold, port):
    """The plain-C solver oracle fed with the node's inputs, iterate carried from tick to tick like the node's solver."""
    T = gold["scenario"].shape[0]
    x = np.tile(np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.0]), (N + 1, 1))   # acados_create's initial iterate
    u = np.zeros((N, 4))
    for t in range(T):
        st, info = port.rti(N, 0.015, gold["solver_x0"][t], gold["solver_yref"][t].reshape(N, 17), gold["solver_yref_e"][t], x, u)
        assert st == int(gold["solver_status"][t, 0])
        assert abs(info.qp_iter - int(gold["solver_qp_iter"][t, 0])) <= 1
        for a, b in ((u[0], gold["solver_u0"][t]), (u[1], gold["solver_u1"][t]), (x[4], gold["solver_x4"][t])):
            assert np.abs(a - b).max() <= 1e-7 * (1 + np.abs(b).max()), (t, a, b)
