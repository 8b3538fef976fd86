"""The multi-device C handle (cfnmpc_multi_*, SURVEY 8b surface C / 8e): shards on every visible GPU -- and, so that the
one-GPU test box exercises the sharding too, several shards on the same GPU -- must reproduce the single-handle solve."""
import numpy as np
import pytest

import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

pytestmark = pytest.mark.gpu
TS = 0.015


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("devices", [(0, 0, 0), "all"])
def test_sharded_handle_equals_single_handle(devices):
    N, B = 20, 1001                      # ragged: shards of 334 / 334 / 333
    if devices == "all":
        devices = tuple(range(n_gpus())) * (2 if n_gpus() == 1 else 1)
    w = wl.helix_batch(B, N, seed=6)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_problem(w).solve(2)
        ref = dict(x=s.get("x_all"), u=s.get("u_all"), st=s.get("status"), it=s.get("qp_iter"), u0=s.get("u", 0))
    with cf.MultiBatchSolver(B, N, TS, devices=devices) as m:
        sh = m.shards()
        assert len(sh) == len(devices) and sh[0][1] == 0 and sum(c for _, _, c in sh) == B
        assert all(sh[i][1] + sh[i][2] == sh[i + 1][1] for i in range(len(sh) - 1))
        m.set_problem(w).solve(2)
        assert np.array_equal(m.get("x_all"), ref["x"]) and np.array_equal(m.get("u_all"), ref["u"])
        assert np.array_equal(m.get("status"), ref["st"]) and np.array_equal(m.get("qp_iter"), ref["it"])
        assert np.array_equal(m.get("u", 0), ref["u0"]) and m.last_solve_ms() > 0
        # solver-wide parameters reach every shard; the host-fed tick and the condensed path work through the handle
        m.set("ubu", [18.0] * 4).set("x", w["x_init"]).set("u", w["u_init"])
        m.solve_from_host(w["x0"], w["yref"], w["yref_e"], n_chunks=3)
        u = m.get("u_all")
        assert u.max() <= 18.0 + 1e-6 and (m.get("status") == 0).all()
        m.set_option("qp_cond_N", 7).set_problem(w).solve(1)
        assert (m.get("status") == 0).all()
    with pytest.raises(cf.CfnmpcError):
        cf.MultiBatchSolver(8, N, TS, devices=(0, 99))
