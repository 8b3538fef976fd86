"""Mint tests/golden/edge_golden.npz with the reference's own acados/HPIPM/BLASFEO build (oracle/_ref/libcfref.so):

* maxiter_{3,5}_*   hover batch (32 instances, seed 11) solved with qp_iter_max = 3 / 5: the reference applies the QP
                    step and returns ACADOS_SUCCESS when HPIPM stops at its iteration limit
                    (ocp_nlp_sqp_rti.c:651-674, ocp_qp_hpipm.c:307-311)
* adv_*             128 deliberately ill-conditioned instances (workloads.adversarial_batch, seed 5) with the HPIPM
                    statistics that show where the reference's safety nets fired: LQ re-factorisation
                    (x_ocp_qp_ipm.c:2029-2059, stat column 11) and corrector iterative refinement (:2311-2318, column 13)

* pcond_{17,25}_*   8 helix instances (seed 21) solved by the reference with qp_cond_N = 17 / 25 (partial condensing,
                    ocp_qp_partial_condensing.c:457-576)

Inputs are regenerated from the seeds by the tests; only the reference's outputs are stored (u of every stage, x of
stages 1, 4, N).  Needs /root/reference:  python tests/golden/make_golden_edge.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from crazyflie_nmpc_b200 import workloads as wl  # noqa: E402
from oracle.oracle import Ref  # noqa: E402

N, TS = 50, 0.015
XSEL = [1, 4, N]


def ref_solve_maxiter(ref, w, itmax):
    B = w["x0"].shape[0]
    s = ref.solver(N, TS)
    s.set_opt_int("qp_iter_max", itmax)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, qi, qs = (np.zeros(B, np.int32) for _ in range(3))
    for i in range(B):
        st[i], qi[i], qs[i], _ = s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], x[i], u[i])
    return dict(status=st, qp_iter=qi, qp_status=qs, x=x, u=u)


def ref_solve_adversarial(ref, w):
    B = w["x0"].shape[0]
    s = ref.solver(N, TS)
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, qi, qs, lq, itref = (np.zeros(B, np.int32) for _ in range(5))
    for i in range(B):
        s.set_weights(w["W"][i], w["W_e"][i])
        s.set_input_bounds(w["lbu"][i], w["ubu"][i])
        st[i], qi[i], qs[i], _ = s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], x[i], u[i])
        stat = s.ipm_stat()
        lq[i], itref[i] = int(stat[1:, 11].sum()), int(stat[1:, 13].sum())
    return dict(status=st, qp_iter=qi, qp_status=qs, lq=lq, itref=itref, x=x, u=u)


def main():
    ref = Ref()
    out = {}
    w = wl.hover_batch(32, N, seed=11)
    for itmax in (3, 5):
        r = ref_solve_maxiter(ref, w, itmax)
        for k in ("status", "qp_iter", "qp_status"):
            out[f"maxiter_{itmax}_{k}"] = r[k]
        out[f"maxiter_{itmax}_u"] = r["u"]
        out[f"maxiter_{itmax}_xsel"] = r["x"][:, XSEL]
        print("maxiter", itmax, np.unique(np.c_[r["status"], r["qp_status"], r["qp_iter"]], axis=0).tolist())
    wa = wl.adversarial_batch(128, N, seed=5)
    r = ref_solve_adversarial(ref, wa)
    for k in ("status", "qp_iter", "qp_status", "lq", "itref"):
        out[f"adv_{k}"] = r[k]
    out["adv_u"] = r["u"]
    out["adv_xsel"] = r["x"][:, XSEL]
    print("adversarial: converged", int((r["qp_status"] == 0).sum()), "maxiter", int((r["qp_status"] == 1).sum()),
          "LQ fired", np.nonzero(r["lq"])[0].tolist(), "itref fired", np.nonzero(r["itref"])[0].tolist())
    # partial condensing: the reference at qp_cond_N = 17 (blocks of 3 and 2 stages) and 25 (blocks of 2)
    wp = wl.helix_batch(8, N, seed=21)
    for cn in (17, 25):
        x, u = wp["x_init"].copy(), wp["u_init"].copy()
        st, it, _ = ref.batch(N, TS, wp["x0"], wp["yref"], wp["yref_e"], x, u, cond_N=cn)
        out[f"pcond_{cn}_status"], out[f"pcond_{cn}_qp_iter"], out[f"pcond_{cn}_u"], out[f"pcond_{cn}_xsel"] = st, it, u, x[:, XSEL]
        print("pcond", cn, "status", np.unique(st).tolist(), "iterations", it.tolist())
    np.savez_compressed(os.path.join(HERE, "edge_golden.npz"), **out)


if __name__ == "__main__":
    main()
