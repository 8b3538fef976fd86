"""Device time of one RTI step of the generic-model libraries (preparation kernel + dense-stage feedback program):
the pendulum on a cart (nx = 4, nu = 1, N = 20) and the Crazyflie OCP on the generic path, against the tuned library.
Usage (GPU box): python profiles/tools/second_model_probe.py [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden"))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl
from make_golden_pendulum import pendulum_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536


def run(name, solver, w):
    ms = []
    for _ in range(4):
        solver.set_problem(w).solve(1)
        ms.append(solver.last_solve_ms())
    st, it = solver.get("status"), solver.get("qp_iter")
    best = min(ms[1:])
    print(f"{name:34s}: step {best:8.2f} ms -> {B / best:8.1f} k solves/s, ok {int((st == 0).sum())}/{B}, mean IPM iterations {it.mean():.2f}", flush=True)


w = pendulum_batch(B, 20, seed=1)
with cf.ModelSolver("pendulum", B) as s:
    run(f"pendulum nx=4 nu=1 N=20 (regs {s.info('regs_preparation')}/{s.info('regs_feedback')})", s, w)
w = wl.hover_batch(B, 50)
with cf.ModelSolver("crazyflie_generic", B, N=50, Ts=0.015) as s:
    run(f"crazyflie, generic path (regs {s.info('regs_preparation')}/{s.info('regs_feedback')})", s, w)
with cf.BatchSolver(B, 50, 0.015) as s:
    run("crazyflie, tuned library", s, w)
