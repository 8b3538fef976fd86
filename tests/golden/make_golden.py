"""Generate tests/golden/crazyflie_rti_golden.npz with the REFERENCE's own implementation.

Runs where /root/reference exists: oracle/Makefile compiles acados + HPIPM + BLASFEO from the
reference tree into oracle/_ref/libcfref.so and oracle/ref_harness.c drives the Crazyflie OCP
through acados_c.  The vectors pin (a) the plain-C oracle, (b) the SIMT-emulated kernel source and
(c) the CUDA path to the reference, also on machines where the reference tree is absent.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from crazyflie_nmpc_b200 import workloads as wl  # noqa: E402
from oracle.oracle import Ref, build  # noqa: E402

TS = 0.015
OUT = os.path.join(ROOT, "tests", "golden", "crazyflie_rti_golden.npz")


def solve_batch(ref, w, N, n_rti=1):
    x, u = w["x_init"].copy(), w["u_init"].copy()
    st, it, _ = ref.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u, n_rti=n_rti, nthreads=4)
    return x, u, st, it


def main():
    build(ref=True)
    ref = Ref()
    g = {}
    # --- known-answer sequence: template iterate, 6 consecutive RTI calls
    N = 50
    w = wl.single_hover(N)
    s = ref.solver(N, TS)
    x, u = w["x_init"][0].copy(), w["u_init"][0].copy()
    ka = []
    for r in range(6):
        st, qi, qs, _ = s.rti(w["x0"][0], w["yref"][0], w["yref_e"][0], x, u)
        ka.append(np.r_[st, qi, u[0, :], x[1], x[4]])
    g["ka_seq"] = np.array(ka)  # [status, qp_iter, u0(4), x1(13), x4(13)] per call
    g["ka_x_final"], g["ka_u_final"] = x.copy(), u.copy()
    # QP data of one more call on the same solver (intermediates)
    st, qi, qs, _ = s.rti(w["x0"][0], w["yref"][0], w["yref_e"][0], x, u)
    s.close()
    # --- batches
    for name, gen, seed, n, Nh in (("hover", wl.hover_batch, 101, 12, 50), ("helix", wl.helix_batch, 102, 12, 50),
                                   ("hover20", wl.hover_batch, 103, 4, 20), ("hover100", wl.hover_batch, 104, 4, 100)):
        w = gen(n, Nh, seed=seed)
        x, u, st, it = solve_batch(ref, w, Nh)
        for k in ("x0", "yref", "yref_e", "x_init", "u_init"):
            g[f"{name}_{k}"] = w[k]
        g[f"{name}_x"], g[f"{name}_u"], g[f"{name}_status"], g[f"{name}_qp_iter"] = x, u, st, it
        if name in ("hover", "helix"):
            x5, u5, st5, it5 = solve_batch(ref, w, Nh, n_rti=5)
            g[f"{name}_x5"], g[f"{name}_u5"] = x5, u5
    # --- config 1 variants
    c = 0
    for template in (True, False):
        for node in (False, True):
            for x0 in (None, [.1, -.05, .3, 1, 0, 0, 0, .1, 0, -.1, 0, 0, 0]):
                w = wl.single_hover(N, template, node, x0)
                for k in ("x0", "yref", "yref_e", "x_init", "u_init"):
                    g[f"cfg1_{c}_{k}"] = w[k]
                for n_rti in (1, 5):
                    x, u, st, it = solve_batch(ref, w, N, n_rti)
                    g[f"cfg1_{c}_x{n_rti}"], g[f"cfg1_{c}_u{n_rti}"] = x, u
                c += 1
    g["cfg1_count"] = np.array(c)
    # --- QP intermediates of the first hover instance
    w = {k: g[f"hover_{k}"][:1] for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
    s = ref.solver(N, TS)
    x, u = w["x_init"][0].copy(), w["u_init"][0].copy()
    s.rti(w["x0"][0], w["yref"][0], w["yref_e"][0], x, u)
    for k, v in s.qp().items():
        g[f"qp_{k}"] = v
    g["qp_ipm_stat"] = s.ipm_stat()
    s.close()
    # --- runtime weights / bounds
    W = np.array([80, 90, 150, 1e-2, 1e-2, 1e-2, 1e-2, 1.0, 1.0, 2.0, 1e-4, 1e-4, 5.0, 0.1, 0.1, 0.2, 0.2])
    WN = 30 * W[:13]
    lbu, ubu = np.array([1.0, 1.0, 2.0, 2.0]), np.array([20.0, 21.0, 20.0, 21.0])
    w = {k: g[f"hover_{k}"][:4] for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
    xs, us = w["x_init"].copy(), w["u_init"].copy()
    for i in range(4):
        s = ref.solver(N, TS)
        s.set_weights(W, WN)
        s.set_input_bounds(lbu, ubu)
        s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], xs[i], us[i])
        s.close()
    g["par_W"], g["par_WN"], g["par_lbu"], g["par_ubu"], g["par_x"], g["par_u"] = W, WN, lbu, ubu, xs, us
    np.savez_compressed(OUT, **g)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB,", len(g), "arrays")


OUT_PHASES = os.path.join(ROOT, "tests", "golden", "crazyflie_rti_golden_phases.npz")


def moved(x0, seed, scale=0.02):
    """A second measurement a little away from the first (quaternion re-normalised)."""
    x = x0 + scale * np.random.default_rng(seed).standard_normal(x0.shape)
    x[..., 3:7] /= np.linalg.norm(x[..., 3:7], axis=-1, keepdims=True)
    return np.ascontiguousarray(x)


def main_phases():
    """Split real-time iteration (rti_phase 1, new measurement, rti_phase 2: ocp_nlp_sqp_rti.c:189-198,1213-1237) and
    non-uniform shooting grids (create_with_discretization, acados_solver.in.c:133-153) run by the reference itself."""
    build(ref=True)
    ref = Ref()
    g = {}
    N, n = 20, 6
    dt = TS * np.concatenate([np.full(6, 0.5), np.linspace(0.6, 2.5, N - 6)])
    g["dt"] = dt
    for name, gen, seed, grid in (("split", wl.helix_batch, 201, None), ("dt", wl.hover_batch, 202, dt),
                                  ("dtsplit", wl.hover_batch, 203, dt)):
        w = gen(n, N, seed=seed)
        x0_fb = moved(w["x0"], seed + 50)
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it = np.zeros(n, np.int32), np.zeros(n, np.int32)
        for i in range(n):
            s = ref.solver(N, TS, dt=grid)
            if name == "dt":
                st[i], it[i], _, _ = s.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], x[i], u[i])
            else:
                st[i], it[i], _ = s.rti_split(w["x0"][i], x0_fb[i], w["yref"][i], w["yref_e"][i], x[i], u[i])
            s.close()
        for k in ("x0", "yref", "yref_e", "x_init", "u_init"):
            g[f"{name}_{k}"] = w[k]
        g[f"{name}_x0_fb"] = x0_fb
        g[f"{name}_x"], g[f"{name}_u"], g[f"{name}_status"], g[f"{name}_qp_iter"] = x, u, st, it
    np.savez_compressed(OUT_PHASES, **g)
    print("wrote", OUT_PHASES, os.path.getsize(OUT_PHASES) // 1024, "KiB,", len(g), "arrays")


if __name__ == "__main__":
    if "--phases-only" not in sys.argv:
        main()
    main_phases()
