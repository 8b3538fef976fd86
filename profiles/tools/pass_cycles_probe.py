"""Where does a warp spend its time?  Per-pass elapsed cycles (summed over warps) of one batched RTI step.
Usage (GPU box): python profiles/tools/pass_cycles_probe.py [B] [N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
w = wl.hover_batch(B, N)
with cf.BatchSolver(B, N, 0.015) as s:
    s.set_option("qp_cond_N", int(os.environ.get("PCOND", "0")))
    s.set_problem(w).solve(1)
    s.debug_pass_cycles(read=False)
    s.set_problem(w).solve(1)
    ms = s.last_solve_ms()
    r = s.debug_pass_cycles()
    tot = sum(v[0] for v in r.values())
    print(f"kernel {ms:.2f} ms, B={B}, N={N}")
    for k, (cyc, calls) in r.items():
        print(f"  {k:20s} {100 * cyc / tot:5.1f}% of warp time  {calls:9d} calls  {cyc / max(calls, 1):10.0f} cycles/call  "
              f"{cyc / max(calls, 1) / (N + 1):7.0f} cycles/stage")
