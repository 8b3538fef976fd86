"""The plain-C oracle on the reference's edge cases (tests/golden/edge_golden.npz, minted with the reference's own
acados/HPIPM build by tests/golden/make_golden_edge.py): the QP-MAXITER branch and ill-conditioned instances where
HPIPM's safety nets fire."""
import os

import numpy as np
import pytest

from conftest import rel_err
from crazyflie_nmpc_b200 import workloads as wl

HERE = os.path.dirname(os.path.abspath(__file__))
N, TS, XSEL = 50, 0.015, [1, 4, 50]


@pytest.fixture(scope="module")
def edge():
    return np.load(os.path.join(HERE, "golden", "edge_golden.npz"))


@pytest.mark.parametrize("itmax", [3, 5])
def test_port_qp_maxiter_branch(port, edge, itmax):
    w = wl.hover_batch(32, N, seed=11)
    port.set_iter_max(itmax)
    try:
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it = port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
    finally:
        port.set_iter_max(50)
    assert np.array_equal(st, edge[f"maxiter_{itmax}_status"]) and np.array_equal(it, edge[f"maxiter_{itmax}_qp_iter"])
    assert (st == 0).all()                      # max-iter is not an error for the RTI step (ocp_nlp_sqp_rti.c:651-674)
    assert rel_err(u, edge[f"maxiter_{itmax}_u"]) < 1e-10 and rel_err(x[:, XSEL], edge[f"maxiter_{itmax}_xsel"]) < 1e-10


def test_port_flags_where_the_reference_nets_fire(port, edge):
    w = wl.adversarial_batch(128, N, seed=5)
    fired = (edge["adv_lq"] > 0) | (edge["adv_itref"] > 0)
    assert fired.sum() >= 3
    flagged = np.zeros(128, bool)
    for i in range(128):
        p = port.params(Wdiag=w["W"][i], WNdiag=w["W_e"][i], lbu=w["lbu"][i], ubu=w["ubu"][i])
        x, u = w["x_init"][i].copy(), w["u_init"][i].copy()
        st, info = port.rti(N, TS, w["x0"][i], w["yref"][i], w["yref_e"][i], x, u, params=p)
        flagged[i] = info.n_lq_flag > 0 or info.n_itref > 0
        assert st == edge["adv_status"][i] and info.qp_status == edge["adv_qp_status"][i], i
        if edge["adv_qp_status"][i] == 0 and not fired[i] and not flagged[i]:
            assert abs(info.qp_iter - edge["adv_qp_iter"][i]) <= 1
            assert rel_err(u, edge["adv_u"][i]) < 1e-6 and rel_err(x[XSEL], edge["adv_xsel"][i]) < 1e-6, i
    assert flagged[fired].all()                 # the restatement detects every instance where the reference's nets fired
    # false alarms only among instances that never converged (the residual threshold is met from the other side)
    assert (edge["adv_qp_status"][flagged & ~fired] != 0).all()
