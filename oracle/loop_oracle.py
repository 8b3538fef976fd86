"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference node's per-tick logic around acados_solve().

Follows crazyflie_controller/src/acados_mpc.cpp (file:line in each function) one vehicle at a time in plain
Python/numpy; used by tests/ to check the device kernels of crazyflie_nmpc_b200/csrc/cf_loop_kernels.h.
Parity PINNED: tests/test_node_golden.py checks these restatements bit-for-bit against what the reference's own,
unmodified NMPC::iteration published and handed to the solver when it was compiled with stand-in ROS headers and run
on the reference's acados build (tests/dropin/build_node.py, tests/golden/make_node_golden.py ->
tests/golden/node_loop_golden.npz: Regulation -> set-point change -> Tracking -> end of table -> Position_Hold); the
integrator part is pinned against the reference's own sim_erk (oracle/ref_harness.c: cfref_sim_*).
"""
import math

import numpy as np

NX, NU, NY = 13, 4, 17
REGULATION, TRACKING, HOLD = 0, 1, 2   # enum order, acados_mpc.cpp:129-133
PI_NODE = 3.14159265358979323846       # acados_mpc.cpp:105


def node_uss():
    """uss = sqrt((mq*g0)/(4*Ct)) as the node's C++ evaluates it (acados_mpc.cpp:107,189,253): mq, Ct, uss are float, g0
    is a double macro -> mq*g0 and the quotient are double, 4*Ct is float, the result is rounded to float."""
    mq, Ct = np.float32(33e-3), np.float32(3.25e-4)
    return float(np.float32(np.sqrt((float(mq) * 9.80665) / float(np.float32(4) * Ct))))


def reference_window(policy, it, setpoint, traj, N, uss, yref_prev, yref_e_prev):
    """acados_mpc.cpp:430-516 for one vehicle. Returns (yref[N,17], yref_e[13], policy', iter')."""
    ys = np.concatenate([yref_prev.reshape(-1), np.r_[yref_e_prev, np.zeros(NU)]]).reshape(N + 1, NY).copy()
    n_steps = 0 if traj is None else traj.shape[0]
    if policy == REGULATION:                       # :435-454
        for k in range(N + 1):
            ys[k] = [setpoint[0], setpoint[1], setpoint[2], 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, uss, uss, uss, uss]
    elif policy == TRACKING:                       # :457-488
        if it < n_steps - N:
            for k in range(N + 1):
                ys[k] = traj[it + k]
            it += 1
        else:
            policy = HOLD                          # nothing is written in this tick
    elif policy == HOLD:                           # :490-513
        last = traj[n_steps - 1]
        for k in range(N + 1):
            ys[k] = [last[0], last[1], last[2], 1.0, 0, 0, 0, 0, 0, 0, 0, 0, 0, uss, uss, uss, uss]
    return ys[:N].copy(), ys[N, :NX].copy(), policy, it   # :584-594


def quatern2euler(w, x, y, z):
    """acados_mpc.cpp:384-404"""
    R11 = 2 * (w * w + x * x) - 1
    R21 = 2 * (x * y - w * z)
    R31 = 2 * (x * z + w * y)
    R32 = 2 * (y * z - w * x)
    R33 = 2 * (w * w + z * z) - 1
    return math.atan2(R32, R33), -math.asin(R31), math.atan2(R21, R11)


def rad2deg(r):
    return r * 180.0 / PI_NODE                      # :411-414


def krpm2pwm(krpm):
    return int(((krpm * 1000) - 4070.3) / 0.2685)   # :421-425 (int conversion truncates)


def commands(u0, u1, x4, fixed_u0=False):
    """What the node publishes (:619-670): int32 motor speeds, Euler set-point, body twist."""
    um = u1 if fixed_u0 else u0
    motors = np.array([int(v) for v in um], np.int32)            # msg/PropellerSpeedsStamped.msg: int32 fields
    q = np.array([x4[3], x4[4], x4[5], x4[6]])
    q = q / math.sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0])   # Eigen normalize()
    phi, theta, psi = quatern2euler(*q)
    twist = np.array([1.0 * rad2deg(theta), -1.0 * rad2deg(phi),
                      float(krpm2pwm((u1[0] + u1[1] + u1[2] + u1[3]) / 4)), rad2deg(x4[12])])
    return motors, np.array([phi, theta, psi]), twist
