"""Golden vectors of the second model (pendulum on a cart, nx = 4, nu = 1) minted by the reference's own code:
oracle/_ref/libcfref_pendulum.so = acados + HPIPM + BLASFEO and the CasADi-generated pendulum functions the reference
ships (acados/examples/c/pendulum_model/), driven by oracle/ref_harness.c with the OCP data of
acados/examples/acados_python/tests/test_ocp_setting.py:150-205 (N = 20, Tf = 1, Q, R, |F| <= 80).
Run where /root/reference exists:  python tests/golden/make_golden_pendulum.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import Ref  # noqa: E402


def pendulum_batch(B, N, seed=0):
    """Swing-up starts around the hanging position (x0 of the test: (0, pi, 0, 0)), initial guess = straight line in the angle
    (test_ocp_setting.py:226-229), zero inputs, zero references."""
    rng = np.random.default_rng(seed)
    x0 = np.array([0.0, np.pi, 0.0, 0.0]) + rng.uniform(-0.4, 0.4, (B, 4)) * np.array([1.0, 1.0, 2.0, 2.0])
    x0[0] = [0.0, np.pi, 0.0, 0.0]
    x_init = np.zeros((B, N + 1, 4))
    x_init[:, :, 0] = x0[:, None, 0]
    x_init[:, :, 1] = np.linspace(x0[:, 1], 0.0, N + 1, axis=1)
    return dict(x0=np.ascontiguousarray(x0), yref=np.zeros((B, N, 5)), yref_e=np.zeros((B, 4)),
                x_init=np.ascontiguousarray(x_init), u_init=np.zeros((B, N, 1)))


def main():
    out = {}
    for N, B, n_rti in ((20, 16, 1), (20, 4, 6), (7, 4, 1)):
        w = pendulum_batch(B, N, seed=N + n_rti)
        rs = Ref("pendulum").solver(N, 1.0 / 20)
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it = np.zeros((B, n_rti), np.int32), np.zeros((B, n_rti), np.int32)
        for i in range(B):
            for r in range(n_rti):
                s, qi, qs, _ = rs.rti(w["x0"][i], w["yref"][i], w["yref_e"][i], x[i], u[i])
                st[i, r], it[i, r] = s, qi
        rs.close()
        tag = f"N{N}_r{n_rti}"
        for k, v in w.items():
            out[f"{tag}_{k}"] = v
        out[f"{tag}_x"], out[f"{tag}_u"], out[f"{tag}_status"], out[f"{tag}_qp_iter"] = x, u, st, it
    path = os.path.join(ROOT, "tests", "golden", "pendulum_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.endswith("_u")})


if __name__ == "__main__":
    main()
