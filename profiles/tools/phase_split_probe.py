"""Device time of the two halves of a split real-time iteration (cfnmpc_batch_prepare / _feedback) next to the fused
step, and of the general (per-interval time step) kernel on a uniform grid.  Usage (GPU box):
python profiles/tools/phase_split_probe.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = wl.hover_batch(B, N)
with cf.BatchSolver(B, N, 0.015) as s:
    fused, prep, fb, gen = [], [], [], []
    for r in range(4):
        s.set_problem(w).solve(1); fused.append(s.last_solve_ms())
    ref = s.get("u_all")
    for r in range(4):
        s.set_problem(w).prepare(); prep.append(s.last_solve_ms())
        s.feedback(); fb.append(s.last_solve_ms())
    same = np.array_equal(ref, s.get("u_all"))
    dt = np.full(N, 0.015); dt[0] = np.nextafter(0.015, 1)   # not uniform -> the general kernel, practically the same problem
    s.set("time_steps", dt)
    for r in range(4):
        s.set_problem(w).solve(1); gen.append(s.last_solve_ms())
    print(f"B={B} N={N}: fused {min(fused):.2f} ms | prepare {min(prep):.2f} ms + feedback {min(fb):.2f} ms "
          f"(bit-identical to fused: {same}) | general kernel {min(gen):.2f} ms")
