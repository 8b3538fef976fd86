"""Device time of one batched RTI step as a function of the batch size (hover workload, N = 50).
Usage (GPU box): python profiles/tools/batch_size_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

N = 50
for B in (1, 64, 1024, 1776, 4096, 16384, 65536, 262144):
    w = wl.hover_batch(B, N)
    with cf.BatchSolver(B, N, 0.015) as s:
        ts = []
        for _ in range(3):
            s.set_problem(w).solve(1)
            ts.append(s.last_solve_ms())
        print(f"B {B:7d}  {min(ts):9.3f} ms  {B / (min(ts) * 1e-3):10.0f} solves/s  warps {s.info('n_slots')}")
