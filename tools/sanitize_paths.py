"""Every launch path of the solve, small enough to run under compute-sanitizer (B = 64, N in {3, 50}):
fused kernel, two-kernel step, split phases, per-interval time steps, host-fed tick, closed-loop tick, predictor.
Each path is also compared with the two-kernel result so a sanitizer run doubles as a consistency check.

    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tools/sanitize_paths.py
(driver: tools/run_sanitizer.sh, logs under profiles/sanitizer/)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B = int(os.environ.get("SAN_B", "64"))
HORIZONS = [int(v) for v in os.environ.get("SAN_N", "3,50").split(",")]
TS = 0.015


def run(N):
    w = wl.helix_batch(B, N, seed=3) if N <= 50 else wl.hover_batch(B, N, seed=3)
    res = {}
    with cf.BatchSolver(B, N, TS) as s:
        s.set_option("two_kernels", 1)
        s.set_problem(w).solve(1)
        res["two_kernels"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set_option("two_kernels", 0)
        s.set_problem(w).solve(1)
        res["fused"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set_option("lin_res_check", 1)
        s.set_problem(w).solve(1)
        res["fused+lin_res_check"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set_option("lin_res_check", 0)
        s.set_problem(w).prepare().feedback()
        res["split"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set("time_steps", np.full(N, TS) * (1.0 + 1e-13 * np.arange(N)))   # general (per-interval) kernels
        s.set_problem(w).solve(1)
        res["vdt"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set("time_steps", np.full(N, TS))
        s.set_option("two_kernels", 1)
        s.set("x", w["x_init"]).set("u", w["u_init"])
        s.solve_from_host(w["x0"], w["yref"], w["yref_e"], n_chunks=4)
        res["host_fed"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        # closed-loop tick + plant step
        s.set_trajectory(wl.helix_table())
        s.set("policy", np.full(B, cf.POLICY_TRACKING, np.int32)).set("traj_iter", np.arange(B, dtype=np.int32))
        s.set("x", w["x_init"]).set("u", w["u_init"])
        s.tick().plant_step(TS)
        s.get("motors"), s.get("twist"), s.get("euler")
    # round 2: multiplier output, partial condensing (block sizes 2 and 3; solve, split and host-fed), full weight matrices
    with cf.BatchSolver(B, N, TS) as s:
        s.set_option("multipliers", 1)
        s.set_problem(w).solve(1)
        res["multipliers"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.get("pi_all"), s.get("lam_all"), s.get("t_all"), s.get("lam_x0", 0)
    loose = {}
    for cond_N in sorted({(N + 1) // 2, (N + 2) // 3} - {N, 0}):
        with cf.BatchSolver(B, N, TS) as s:
            s.set_option("qp_cond_N", cond_N)
            s.set_problem(w).solve(1)
            loose[f"qp_cond_N={cond_N}"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
            s.set_problem(w).prepare().feedback()
            loose[f"qp_cond_N={cond_N} split"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
            s.set("x", w["x_init"]).set("u", w["u_init"])
            s.solve_from_host(w["x0"], w["yref"], w["yref_e"], n_chunks=4)
            loose[f"qp_cond_N={cond_N} host_fed"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
    Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
    tab = np.zeros((N + 1, 17, 17))
    tab[:N] = np.diag(Q)
    tab[N, :13, :13] = np.diag(50 * Q[:13])
    with cf.BatchSolver(B, N, TS) as s:
        s.set("W_dense_table", tab)          # the default weights as full matrices: same problem, dense-Hessian kernel
        s.set_problem(w).solve(1)
        loose["W_dense_table"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
        s.set("W_dense_table", wl.dense_weight_table(N, seed=1))
        s.set_problem(w).solve(1)
        s.get("x_all")
    # iterative refinement (general kernels; test mode 4 makes every corrector inaccurate so that the refinement runs)
    with cf.BatchSolver(B, N, TS) as s:
        s.set_option("lin_res_check", 4)
        s.set_problem(w).solve(1)
        fl = s.get("flags")
        assert ((fl & 32) != 0).any() and ((fl & 64) == 0).all(), fl
        s.set_option("lin_res_check", 2)
        s.set_problem(w).solve(1)
        res["lin_res_check=2"] = (s.get("x_all"), s.get("u_all"), s.get("status"))
    ref = res["two_kernels"]
    for k, v in loose.items():   # same QP solution, different arithmetic: interior-point tolerances
        ex = float(np.abs(v[0] - ref[0]).max())
        eu = float(np.abs(v[1] - ref[1]).max())
        ok = (v[2] == ref[2]).all() and ex < 1e-5 and eu < 1e-5
        print(f"N={N} {k:22s} status_ok={int((v[2] == 0).sum())}/{B} max|dx|={ex:.2e} max|du|={eu:.2e} {'OK' if ok else 'MISMATCH'}")
        assert ok, k
    for k, v in res.items():
        ex = float(np.abs(v[0] - ref[0]).max())
        eu = float(np.abs(v[1] - ref[1]).max())
        ok = (v[2] == ref[2]).all() and ex < 1e-9 and eu < 1e-9
        print(f"N={N} {k:22s} status_ok={int((v[2] == 0).sum())}/{B} max|dx|={ex:.2e} max|du|={eu:.2e} {'OK' if ok else 'MISMATCH'}")
        assert ok, k
    with cf.SimBatch(B, sens_forw=True) as sim:
        sim.set("x", w["x0"]).set("u", w["u_init"][:, 0]).set("T", [TS]).solve()
        sim.get("xn"), sim.get("S_forw")
    with cf.SimBatch(B) as sim:
        sim.set("x", w["x0"]).set("u", w["u_init"][:, 0]).set("T", [TS]).solve()
        sim.get("xn")


def run_second_model():
    """The generic-model library (pendulum, nx = 4, nu = 1): preparation kernel + dense-stage feedback program."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    from make_golden_pendulum import pendulum_batch
    wp = pendulum_batch(B, 20, seed=2)
    with cf.ModelSolver("pendulum", B) as s:
        s.set_problem(wp).solve(2)
        st = s.get("status")
        s.set_problem(wp).prepare().feedback()
        print(f"pendulum status_ok={int((st == 0).sum())}/{B} OK")
        assert (st == 0).all()


for N in HORIZONS:
    run(N)
run_second_model()
print("sanitize_paths: done")
