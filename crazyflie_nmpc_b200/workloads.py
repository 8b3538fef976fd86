"""Synthetic batches of Crazyflie NMPC problems (host side, numpy).

Input distributions follow SURVEY.md section 8(d): the dynamic-reconfigure ranges of
`crazyflie_controller/config/crazyflie_params.cfg:8-36` for set-points, and the
reference-window indexing of `NMPC::iteration`
(`crazyflie_controller/src/acados_mpc.cpp:435-454` regulation, `:463-482` tracking).

Every generator returns a dict of C-contiguous float64 arrays:
  x0      [B,13]        measured state (becomes lbx_0 = ubx_0)
  yref    [B,N,17]      stage references, cost order y = [x(13); u(4)]
  yref_e  [B,13]        terminal reference
  x_init  [B,N+1,13]    initial iterate, states
  u_init  [B,N,4]       initial iterate, inputs
"""
import numpy as np

NX, NU, NY = 13, 4, 17
# model constants, crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py:34-42
G0, MQ, CT = 9.8066, 33e-3, 3.25e-4
TS = 0.015  # Tf/N = 0.75/50, generate_c_code.py:41-42
U_MAX = 22.0


def hover_speed():
    """hov_w of generate_c_code.py:55 (model gravity 9.8066)."""
    return float(np.sqrt((MQ * G0) / (4 * CT)))


def node_hover_speed():
    """uss as the ROS node computes it (acados_mpc.cpp:107,189,253): `uss = sqrt((mq*g0)/(4*Ct))` with float mq, Ct,
    uss and the double macro g0 = 9.80665 -- double product and quotient, float 4*Ct, result rounded to float."""
    num = float(np.float32(MQ)) * 9.80665
    den = float(np.float32(4.0) * np.float32(CT))
    return float(np.float32(np.sqrt(num / den)))


def helix_table():
    """The 1050x17 reference trajectory `crazyflie_controller/traj/helix_traj.txt`,
    regenerated analytically: radius 0.3 m, angle linspace(0,15 rad,1000), z
    linspace(0.04,2.038,1000), level attitude, hover thrust 15.7777, values rounded
    to 4 decimals as in the file, last sample held for 50 more rows.
    tests/test_workloads.py checks it against the file where /root/reference exists."""
    k = np.minimum(np.arange(1050), 999)
    th = np.linspace(0.0, 15.0, 1000)[k]
    T = np.zeros((1050, NY))
    T[:, 0] = 0.3 * np.cos(th)
    T[:, 1] = 0.3 * np.sin(th)
    T[:, 2] = np.linspace(0.04, 2.038, 1000)[k]
    T[:, 3] = 1.0
    T[:, 13:] = 15.7777
    return np.round(T, 4) + 0.0  # +0.0 turns -0.0 into 0.0


def load_trajectory(path):
    """Read a whitespace table [x(13) u(4)] per 15 ms row, the reference's
    trajectory file format (acados_mpc.cpp:258-283)."""
    T = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if T.shape[1] != NY:
        raise ValueError(f"trajectory file must have {NY} columns, got {T.shape[1]}")
    return np.ascontiguousarray(T)


def _quat_from_rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r / 2), np.sin(r / 2), np.cos(p / 2), np.sin(p / 2), np.cos(y / 2), np.sin(y / 2)
    return np.stack([cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy,
                     cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy], axis=-1)


def _perturbation(rng, B):
    """Config-2 perturbation: position +-0.3 m, roll/pitch/yaw +-20 deg, v_b +-0.5 m/s, omega +-1 rad/s."""
    dp = rng.uniform(-0.3, 0.3, (B, 3))
    rpy = rng.uniform(-np.deg2rad(20), np.deg2rad(20), (B, 3))
    q = _quat_from_rpy(rpy[:, 0], rpy[:, 1], rpy[:, 2])
    v = rng.uniform(-0.5, 0.5, (B, 3))
    w = rng.uniform(-1.0, 1.0, (B, 3))
    return dp, q, v, w


def hover_batch(B, N=50, seed=20261017):
    """Config 2: hover regulation about random set-points, random feasible x0,
    hover-style initial iterate (x_k = x0, u_k = uss)."""
    rng = np.random.default_rng(seed)
    uss = hover_speed()
    pd = np.concatenate([rng.uniform(-1, 1, (B, 2)), rng.uniform(0, 1, (B, 1))], axis=1)
    dp, q, v, w = _perturbation(rng, B)
    x0 = np.concatenate([pd + dp, q, v, w], axis=1)
    y = np.zeros((B, NY))
    y[:, 0:3] = pd
    y[:, 3] = 1.0
    y[:, 13:] = uss
    yref = np.ascontiguousarray(np.broadcast_to(y[:, None, :], (B, N, NY)))
    return dict(x0=np.ascontiguousarray(x0), yref=yref, yref_e=np.ascontiguousarray(y[:, :NX]),
                x_init=np.ascontiguousarray(np.broadcast_to(x0[:, None, :], (B, N + 1, NX))),
                u_init=np.full((B, N, NU), uss))


def helix_batch(B, N=50, seed=20261018, table=None):
    """Config 3: helix tracking. Window start i0 ~ randint(0, rows-N); yref_k = table[i0+k],
    yref_e = table[i0+N][:13] (acados_mpc.cpp:463-482); x0 = table[i0][:13] + perturbation with
    the quaternion re-normalised; iterate x_k = yref_k[:13], u_k = uss."""
    rng = np.random.default_rng(seed)
    T = helix_table() if table is None else table
    uss = hover_speed()
    i0 = rng.integers(0, T.shape[0] - N, B)
    idx = i0[:, None] + np.arange(N + 1)[None, :]
    win = T[idx]  # [B,N+1,17]
    dp, q, v, w = _perturbation(rng, B)
    x0 = win[:, 0, :NX].copy()
    x0[:, 0:3] += dp
    # compose the small random attitude with the (level) reference attitude, then renormalise
    x0[:, 3:7] = q
    x0[:, 3:7] /= np.linalg.norm(x0[:, 3:7], axis=1, keepdims=True)
    x0[:, 7:10] += v
    x0[:, 10:13] += w
    return dict(x0=np.ascontiguousarray(x0), yref=np.ascontiguousarray(win[:, :N, :]),
                yref_e=np.ascontiguousarray(win[:, N, :NX]), x_init=np.ascontiguousarray(win[:, :, :NX]),
                u_init=np.full((B, N, NU), uss), i0=i0)


def single_hover(N=50, template_iterate=True, node_yref=False, x0=None):
    """Config 1: the single hover-regulation OCP. `template_iterate` starts from the generated
    solver's initial guess x_k = [0,0,0,1,0...], u_k = 0 (acados_solver.in.c:2323-2352);
    `node_yref` uses the ROS node's regulation reference (set-point (0,0,0.40), float uss)."""
    uss = hover_speed()
    if node_yref:
        u_ref = node_hover_speed()
        y = np.array([0, 0, 0.40, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, u_ref, u_ref, u_ref, u_ref], float)
    else:
        y = np.array([0, 0, 0.5, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, uss, uss, uss, uss], float)
    x0 = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], float) if x0 is None else np.asarray(x0, float)
    xi = np.array([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0], float) if template_iterate else x0
    ui = 0.0 if template_iterate else uss
    return dict(x0=x0[None].copy(), yref=np.ascontiguousarray(np.broadcast_to(y, (1, N, NY))),
                yref_e=y[None, :NX].copy(), x_init=np.ascontiguousarray(np.broadcast_to(xi, (1, N + 1, NX))),
                u_init=np.full((1, N, NU), ui))


def adversarial_batch(B, N=50, seed=7, level=1.0):
    """Deliberately ill-conditioned instances for probing the interior-point safety nets (the reference's LQ
    re-factorisation / iterative refinement, external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2029-2059,2311-2318): weights spread
    log-uniformly over 1e-8..1e6 per component, 60-90 degree tilts, |omega| up to 20 rad/s, fast translation, input
    boxes that are nearly coincident or exclude the hover thrust.  Extra keys: W [B,17], W_e [B,13], lbu, ubu [B,4]."""
    rng = np.random.default_rng(seed)
    uss = hover_speed()
    w = hover_batch(B, N, seed=seed + 1)
    sgn = rng.choice([-1.0, 1.0], (B, 3))
    rpy = sgn * rng.uniform(np.deg2rad(60), np.deg2rad(90), (B, 3)) * level
    rpy[:, 2] = rng.uniform(-np.pi, np.pi, B)
    x0 = w["x0"].copy()
    x0[:, 3:7] = _quat_from_rpy(rpy[:, 0], rpy[:, 1], rpy[:, 2])
    x0[:, 7:10] = rng.uniform(-3, 3, (B, 3)) * level
    om = rng.normal(size=(B, 3))
    x0[:, 10:13] = om / np.linalg.norm(om, axis=1, keepdims=True) * rng.uniform(10, 20, (B, 1)) * level
    W = 10.0 ** rng.uniform(-8, 6, (B, NY))
    W_e = 10.0 ** rng.uniform(-8, 6, (B, NX))
    kind = rng.integers(0, 3, B)
    mid = rng.uniform(0.5, 21.5, (B, NU))
    half = 10.0 ** rng.uniform(-7, 0, (B, NU))
    lbu = np.where(kind[:, None] == 0, 0.0, mid - half)
    ubu = np.where(kind[:, None] == 0, U_MAX, mid + half)
    u_init = np.broadcast_to((0.5 * (lbu + ubu))[:, None, :], (B, N, NU)) if False else np.full((B, N, NU), uss)
    w.update(x0=np.ascontiguousarray(x0), x_init=np.ascontiguousarray(np.broadcast_to(x0[:, None, :], (B, N + 1, NX))),
             u_init=np.ascontiguousarray(u_init), W=W, W_e=W_e, lbu=np.ascontiguousarray(lbu), ubu=np.ascontiguousarray(ubu))
    return w


def dense_weight_table(N, seed=0, coupling=0.3):
    """Full (non-diagonal) SPD weight matrices, one per stage, [N+1][17][17] in cost order y = [x;u] (row N: W_e in its
    leading 13 x 13 block): the node's diagonal weights scaled per stage and correlated by a random SPD factor.  What a
    caller builds with per-stage ocp_nlp_cost_model_set(.., k, "W", ..) calls (ocp_nlp_cost_ls.c:301-331)."""
    rng = np.random.default_rng(seed)
    Q = np.array([120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0, 0.06, 0.06, 0.06, 0.06])
    tab = np.zeros((N + 1, NY, NY))
    for k in range(N + 1):
        n = NY if k < N else NX
        d = np.sqrt(Q[:n] * (50.0 if k == N else 1.0) * rng.uniform(0.5, 2.0, n))
        C = np.eye(n) + coupling * rng.uniform(-1, 1, (n, n))
        C = 0.5 * (C + C.T)
        Wk = d[:, None] * (C @ C.T) * d[None, :]
        tab[k, :n, :n] = 0.5 * (Wk + Wk.T)
    return tab
