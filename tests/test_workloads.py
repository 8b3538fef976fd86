import os

import numpy as np
import pytest

from crazyflie_nmpc_b200 import workloads as wl

REF_TRAJ = "/root/reference/crazyflie_controller/traj/helix_traj.txt"


def test_helix_table_shape_and_hold_tail():
    T = wl.helix_table()
    assert T.shape == (1050, 17)
    assert np.allclose(np.hypot(T[:, 0], T[:, 1]), 0.3, atol=1e-4)
    assert T[0, 2] == 0.04 and T[999, 2] == 2.038 and (T[999:] == T[999]).all()
    assert (T[:, 3] == 1).all() and (T[:, 4:13] == 0).all() and (T[:, 13:] == 15.7777).all()


@pytest.mark.skipif(not os.path.exists(REF_TRAJ), reason="reference tree not present")
def test_helix_table_equals_reference_file():
    assert np.array_equal(wl.helix_table(), wl.load_trajectory(REF_TRAJ))


def test_batches_are_seeded_feasible_and_shaped():
    for gen in (wl.hover_batch, wl.helix_batch):
        a, b = gen(16, 50), gen(16, 50)
        for k in ("x0", "yref", "yref_e", "x_init", "u_init"):
            assert np.array_equal(a[k], b[k]) and a[k].flags["C_CONTIGUOUS"] and a[k].dtype == np.float64
        assert a["x0"].shape == (16, 13) and a["yref"].shape == (16, 50, 17) and a["x_init"].shape == (16, 51, 13)
        assert np.allclose(np.linalg.norm(a["x0"][:, 3:7], axis=1), 1.0)
        assert (a["u_init"] > 0).all() and (a["u_init"] < 22).all()
    h = wl.helix_batch(8, 50)
    T = wl.helix_table()
    assert np.array_equal(h["yref"][3], T[h["i0"][3]: h["i0"][3] + 50])
    assert np.array_equal(h["yref_e"][3], T[h["i0"][3] + 50, :13])


def test_node_hover_speed_is_float_precision():
    # acados_mpc.cpp:189,253 computes uss in float with g0 = 9.80665
    assert abs(wl.node_hover_speed() - np.sqrt(0.033 * 9.80665 / (4 * 3.25e-4))) < 1e-5
    assert abs(wl.hover_speed() - 15.777730167257) < 1e-9
