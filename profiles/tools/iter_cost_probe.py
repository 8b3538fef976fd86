"""Device time of the RTI kernel as a function of the (truncated) IPM iteration count: intercept = linearisation +
first residual/factorisation sweep, slope = one interior-point iteration.  Usage (GPU box): python profiles/tools/iter_cost_probe.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B, N = 65536, 50
w = wl.hover_batch(B, N)
with cf.BatchSolver(B, N, 0.015) as s:
    pts = []
    for m in (1, 2, 3, 4):
        s.debug_max_ipm_iter(m)
        ts = []
        for _ in range(3):
            s.set_problem(w).solve(1)
            ts.append(s.last_solve_ms())
        it = s.get("qp_iter").mean()
        pts.append((it, min(ts)))
        print(f"max_iter {m}: mean iters {it:.3f}  {min(ts):8.2f} ms")
    a = np.polyfit([p[0] for p in pts], [p[1] for p in pts], 1)
    print(f"slope {a[0]:.2f} ms per IPM iteration, intercept {a[1]:.2f} ms (B = {B})")
