"""Device time of one batched RTI step for the three register/occupancy variants of the kernel.
Usage (GPU box): python profiles/occupancy_sweep.py [B] [N]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np, crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl
w = wl.hover_batch({B}, {N})
with cf.BatchSolver({B}, {N}, 0.015) as s:
    ts = []
    for r in range(4):
        s.set_problem(w).solve(1); ts.append(s.last_solve_ms())
    print("warps/block", s.info("warps_per_block"), "blocks/SM", s.info("blocks_per_sm"), "regs", s.info("regs_per_thread"), "grid", s.info("grid"),
          "ms", [round(t, 2) for t in ts], "solves/s %.0f" % ({B} / (min(ts) * 1e-3)), "iters", s.get("qp_iter").mean())
"""
for wpb, mb in ((4, 3), (4, 4), (4, 5), (2, 9)):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CFNMPC_MIN_BLOCKS=str(mb), CFNMPC_WARPS_PER_BLOCK=str(wpb)))

