"""Mint tests/golden/node_loop_golden.npz by running the reference's UNMODIFIED ROS node class (NMPC::iteration of
crazyflie_controller/src/acados_mpc.cpp, compiled by tests/dropin/build_node.py with stand-in ROS headers) on top of the
reference's own acados/HPIPM/BLASFEO build, through Regulation -> set-point change -> Tracking -> end of the table ->
Position_Hold.  Needs /root/reference (run in the build container):  python tests/golden/make_node_golden.py"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from crazyflie_nmpc_b200 import workloads as wl  # noqa: E402

N, NX, NY = 50, 13, 17
TABLE_ROWS = 100          # tracking is valid while iter < TABLE_ROWS - N = 50
T_TRACK, N_TICKS = 15, 82


def scenario(seed=5):
    rng = np.random.default_rng(seed)
    T = wl.helix_table()[:TABLE_ROWS]
    sc = np.zeros((N_TICKS, 17))
    sc[:, 0] = -1
    sc[0, :4] = [0, 0.2, -0.1, 0.4]
    sc[5, :4] = [0, -0.1, 0.15, 0.6]
    sc[T_TRACK, 0] = 1
    p = np.array([0.0, 0.0, 0.3])
    for t in range(N_TICKS):
        if t < T_TRACK:
            sp = sc[0, 1:4] if t < 5 else sc[5, 1:4]
            p = sp + (p - sp) * 0.93
            x = np.r_[p, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0]
        else:
            it = min(t - T_TRACK, TABLE_ROWS - 1)
            x = T[it, :NX].copy()
        x[0:3] += rng.uniform(-0.05, 0.05, 3)
        rpy = rng.uniform(-np.deg2rad(8), np.deg2rad(8), 3)
        x[3:7] = wl._quat_from_rpy(rpy[0], rpy[1], rpy[2])
        x[7:10] = rng.uniform(-0.2, 0.2, 3)
        x[10:13] = rng.uniform(-0.5, 0.5, 3)
        sc[t, 4:] = x
    return sc, T


def write_table(path, T):
    """Same text format as crazyflie_controller/traj/*.txt (whitespace separated, 17 columns)."""
    with open(path, "w") as f:
        for r in T:
            f.write(" ".join(f"{v:.4f}" for v in r) + "\n")


def run_node(exe, sc, T, want_log):
    with tempfile.TemporaryDirectory() as d:
        scn, tab, out, log = (os.path.join(d, n) for n in ("scenario.bin", "traj.txt", "out.bin", "log.bin"))
        np.concatenate([[float(sc.shape[0])], sc.reshape(-1)]).astype(np.float64).tofile(scn)
        write_table(tab, T)
        subprocess.run([exe, scn, tab, out] + ([log] if want_log else []), check=True)
        o = np.fromfile(out).reshape(sc.shape[0], 8)
        lg = np.fromfile(log).reshape(sc.shape[0], -1) if want_log else None
    return o, lg


def main():
    sys.path.insert(0, os.path.join(ROOT, "tests", "dropin"))
    import build_node
    if not build_node.build():
        raise SystemExit("needs /root/reference")
    sc, T = scenario()
    out, log = run_node(os.path.join(ROOT, "tests", "dropin", "_build", "node_ref"), sc, T, True)
    o = 0
    f = {}
    for name, w in (("x0", NX), ("yref", N * NY), ("yref_e", NX), ("status", 1), ("u0", 4), ("u1", 4), ("x4", NX), ("qp_iter", 1)):
        f[name] = log[:, o:o + w]
        o += w
    assert o == log.shape[1]
    np.savez_compressed(os.path.join(HERE, "node_loop_golden.npz"), scenario=sc, table=T, published=out,
                        **{"solver_" + k: v for k, v in f.items()})
    print("ticks", sc.shape[0], "status", np.unique(f["status"]), "qp_iter", f["qp_iter"].min(), f["qp_iter"].max())
    print("motors[0:3]", out[:3, :4], "twist[0]", out[0, 4:])


if __name__ == "__main__":
    main()
