"""Does instruction-cache / phase diversity limit the RTI kernel?  Solve a batch of IDENTICAL instances (every warp
executes the same instruction stream at nearly the same time) and compare the device time per IPM iteration with
the usual random batch.  Usage (GPU box): python profiles/tools/phase_lock_probe.py [B]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import crazyflie_nmpc_b200 as cf
from crazyflie_nmpc_b200 import workloads as wl

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 50
w = wl.hover_batch(B, N)
with cf.BatchSolver(B, N, 0.015) as s:
    s.set_option("qp_cond_N", int(os.environ.get("PCOND", "0")))
    def run(ww):
        ts = []
        for _ in range(3):
            s.set_problem(ww).solve(1)
            ts.append(s.last_solve_ms())
        it = s.get("qp_iter")
        return min(ts), it.mean()
    t, it = run(w)
    print(f"random batch     : {t:8.2f} ms  iters {it:.2f}  -> {t / (it + 1.2) * 1e3 / B:.4f} us per instance-iteration")
    iters = s.get("qp_iter")
    for target in (5, 6, 7, 8):
        i = int(np.nonzero(iters == target)[0][0])
        wi = {k: np.ascontiguousarray(np.broadcast_to(v[i:i + 1], v.shape)) for k, v in w.items()}
        t, it = run(wi)
        print(f"identical (it={target}) : {t:8.2f} ms  iters {it:.2f}  -> {t / (it + 1.2) * 1e3 / B:.4f} us per instance-iteration")
