"""Device time of one RTI step with the QP partially condensed (option qp_cond_N) against the uncondensed default.
Usage (GPU box): python profiles/tools/pcond_probe.py [B] [N]   (launch shape: CFNMPC_PC_WARPS_PER_BLOCK / CFNMPC_PC_MIN_BLOCKS)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = wl.hover_batch(B, N)
for cond_N in [0] + [int(v) for v in os.environ.get("PCOND_LIST", "17,25").split(",")]:
    with cf.BatchSolver(B, N, 0.015) as s:
        s.set_option("qp_cond_N", cond_N)
        ms, ph = [], []
        for _ in range(4):
            s.set_problem(w).solve(1)
            ms.append(s.last_solve_ms())
            ph.append(s.last_phase_ms())
        st, it = s.get("status"), s.get("qp_iter")
        best = min(ms[1:])
        print(f"qp_cond_N={cond_N or N:3d} block={s.info('pcond_block_size')} shape={s.info('pcond_warps_per_block')}x{s.info('pcond_blocks_per_sm')} "
              f"regs={s.info('pcond_regs_per_thread')}: step {best:7.2f} ms (prep {ph[-1][0]:.2f} + feedback {ph[-1][1]:.2f}) "
              f"-> {B / best * 1e3 / 1e3:7.1f} k solves/s, ok {int((st == 0).sum())}/{B}, mean IPM iterations {it.mean():.2f}", flush=True)
