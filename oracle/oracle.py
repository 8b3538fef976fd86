"""TEST INFRASTRUCTURE ONLY -- ctypes access to the two CPU checkers.

* ``Port``  : oracle/liboracle.so, our plain-C restatement (cfnmpc_oracle.c).
* ``Ref``   : oracle/_ref/libcfref.so, the reference's own acados/HPIPM/BLASFEO
              code compiled from /root/reference (oracle/Makefile, ref_harness.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(crazyflie_nmpc_b200) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NX, NU, NV = 13, 4, 17
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _P(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _I(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def build(ref=True, quiet=True):
    """Compile the checkers (idempotent). `ref` is skipped where /root/reference is absent."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-j8", "-C", HERE] + targets, check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.DEVNULL if quiet else None)


class CfoParams(ctypes.Structure):
    _fields_ = [("Wdiag", ctypes.c_double * 17), ("WNdiag", ctypes.c_double * 13),
                ("lbu", ctypes.c_double * 4), ("ubu", ctypes.c_double * 4), ("has_u0", ctypes.c_int),
                ("lbu0", ctypes.c_double * 4), ("ubu0", ctypes.c_double * 4)]


class CfoInfo(ctypes.Structure):
    _fields_ = [("qp_iter", ctypes.c_int), ("qp_status", ctypes.c_int), ("n_lq_flag", ctypes.c_int),
                ("n_itref", ctypes.c_int), ("res", ctypes.c_double * 4)]


class Port:
    """Plain-C restatement."""

    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = ctypes.CDLL(path)
        L.cfo_rti.restype = ctypes.c_int
        L.cfo_rti.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams), _dp, _dp, _dp, _dp, _dp,
                              ctypes.POINTER(CfoInfo), _dp, _dp]
        L.cfo_linearize.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams)] + [_dp] * 10
        L.cfo_batch.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams), ctypes.c_int, ctypes.c_int,
                                _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.cfo_batch_pcond.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.cfo_rti_pcond.restype = ctypes.c_int
        L.cfo_rti_pcond.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams), ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                    ctypes.POINTER(CfoInfo)]
        L.cfo_erk4.argtypes = [_dp, _dp, ctypes.c_double, _dp, _dp, _dp]
        L.cfo_sim.argtypes = [_dp, _dp, ctypes.c_double, ctypes.c_int, _dp]
        L.cfo_ode.argtypes = [_dp, _dp, _dp]
        L.cfo_default_params.argtypes = [ctypes.POINTER(CfoParams)]
        L.cfo_set_iter_max.argtypes = [ctypes.c_int]
        L.cfo_set_dense_weights.argtypes = [_dp, ctypes.c_int]
        L.cfo_record_multipliers.argtypes = [ctypes.c_int]
        L.cfo_last_multipliers.argtypes = [_dp, _dp, _dp]
        L.cfo_set_stage_weights.argtypes = [_dp, ctypes.c_int]
        L.cfo_set_time_steps.argtypes = [_dp, ctypes.c_int]
        L.cfo_set_stage_bounds.argtypes = [_dp, ctypes.c_int]
        L.cfo_rti_split.restype = ctypes.c_int
        L.cfo_rti_split.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.POINTER(CfoParams), _dp, _dp, _dp, _dp, _dp, _dp,
                                    ctypes.POINTER(CfoInfo)]

    def record_multipliers(self, N=0):
        """Keep the multipliers of the iterate after the next rti() calls with horizon N (0: off); global in the checker."""
        self.lib.cfo_record_multipliers(int(N))
        self._mult_N = int(N)

    def multipliers(self):
        """(pi [N,13], lam0 [2,17], lam [N-1,2,4], t0 [2,17], t [N-1,2,4]) of the last rti(); lam = [lower | upper]."""
        N = self._mult_N
        pi, lam, t = np.zeros((N, NX)), np.zeros(2 * NV + 2 * NU * (N - 1)), np.zeros(2 * NV + 2 * NU * (N - 1))
        self.lib.cfo_last_multipliers(_P(pi), _P(lam), _P(t))
        return (pi, lam[:2 * NV].reshape(2, NV), lam[2 * NV:].reshape(N - 1, 2, NU),
                t[:2 * NV].reshape(2, NV), t[2 * NV:].reshape(N - 1, 2, NU))

    def set_dense_weights(self, tab=None):
        """Full weight matrices per stage, [N+1][17][17] in cost order y = [x;u] (row N: W_e in its leading 13 x 13 block);
        global in the checker, None returns to the diagonal weights."""
        if tab is None:
            self.lib.cfo_set_dense_weights(None, 0)
        else:
            tab = np.ascontiguousarray(tab, float)
            self.lib.cfo_set_dense_weights(_P(tab), tab.shape[0])

    def set_iter_max(self, n=50):
        """qp_iter_max of the interior-point loop (global in the checker; 50 = the reference configuration)."""
        self.lib.cfo_set_iter_max(int(n))

    def set_stage_weights(self, tab=None):
        """Cost weights per stage, [N+1][17] (row N = W_e, 13 used); global in the checker, None returns to the params."""
        if tab is None:
            self.lib.cfo_set_stage_weights(None, 0)
        else:
            tab = np.ascontiguousarray(tab, float)
            self.lib.cfo_set_stage_weights(_P(tab), tab.shape[0])

    def set_time_steps(self, dt=None):
        """Non-uniform shooting grid (global in the checker; None returns to the uniform grid)."""
        if dt is None:
            self.lib.cfo_set_time_steps(None, 0)
        else:
            dt = np.ascontiguousarray(dt, float)
            self.lib.cfo_set_time_steps(_P(dt), dt.size)

    def set_stage_bounds(self, tab=None):
        """Input box per stage, [N][8] = lbu | ubu (global in the checker; None returns to the boxes of the params)."""
        if tab is None:
            self.lib.cfo_set_stage_bounds(None, 0)
        else:
            tab = np.ascontiguousarray(tab, float)
            self.lib.cfo_set_stage_bounds(_P(tab), tab.shape[0])

    def rti_split(self, N, Ts, x0_prep, x0_fb, yref, yref_e, x, u, params=None):
        """Preparation with x0_prep, feedback with x0_fb; x,u updated in place."""
        info = CfoInfo()
        a = [np.ascontiguousarray(v, float) for v in (x0_prep, x0_fb, yref, yref_e)]
        st = self.lib.cfo_rti_split(N, Ts, ctypes.byref(params) if params else None, *[_P(v) for v in a], _P(x), _P(u),
                                    ctypes.byref(info))
        return st, info

    def params(self, Wdiag=None, WNdiag=None, lbu=None, ubu=None, lbu0=None, ubu0=None):
        p = CfoParams()
        self.lib.cfo_default_params(ctypes.byref(p))
        if lbu0 is not None or ubu0 is not None:   # stage-0 bounds default to the path bounds
            p.has_u0 = 1
            lbu0 = lbu0 if lbu0 is not None else (lbu if lbu is not None else list(p.lbu))
            ubu0 = ubu0 if ubu0 is not None else (ubu if ubu is not None else list(p.ubu))
        for name, v in (("Wdiag", Wdiag), ("WNdiag", WNdiag), ("lbu", lbu), ("ubu", ubu), ("lbu0", lbu0), ("ubu0", ubu0)):
            if v is not None:
                arr = getattr(p, name)
                for i, e in enumerate(v):
                    arr[i] = float(e)
        return p

    def ode(self, x, u):
        f = np.zeros(NX)
        self.lib.cfo_ode(_P(np.ascontiguousarray(x, float)), _P(np.ascontiguousarray(u, float)), _P(f))
        return f

    def erk4(self, x, u, h):
        xn, A, B = np.zeros(NX), np.zeros((NX, NX)), np.zeros((NX, NU))
        self.lib.cfo_erk4(_P(np.ascontiguousarray(x, float)), _P(np.ascontiguousarray(u, float)), h, _P(xn), _P(A), _P(B))
        return xn, A, B

    def sim(self, x, u, T, n_steps=1):
        xn = np.zeros(NX)
        self.lib.cfo_sim(_P(np.ascontiguousarray(x, float)), _P(np.ascontiguousarray(u, float)), T, n_steps, _P(xn))
        return xn

    def linearize(self, N, Ts, x0, yref, yref_e, x, u, params=None):
        BAbt, b = np.zeros((N, NV, NX)), np.zeros((N, NX))
        rqz, dl, du = np.zeros(N * NV + NX), np.zeros(NV + NU * (N - 1)), np.zeros(NV + NU * (N - 1))
        a = [np.ascontiguousarray(v, float) for v in (x0, yref, yref_e, x, u)]
        self.lib.cfo_linearize(N, Ts, ctypes.byref(params) if params else None, *[_P(v) for v in a],
                               _P(BAbt), _P(b), _P(rqz), _P(dl), _P(du))
        return dict(BAbt=BAbt, b=b, rqz=rqz, d_lb=dl, d_ub=du)

    def rti(self, N, Ts, x0, yref, yref_e, x, u, params=None, want_step=False):
        """One RTI step; x,u are updated in place. Returns (status, info[, dux, dpi])."""
        info = CfoInfo()
        dux = np.zeros(N * NV + NX) if want_step else None
        dpi = np.zeros(N * NX) if want_step else None
        a = [np.ascontiguousarray(v, float) for v in (x0, yref, yref_e)]
        st = self.lib.cfo_rti(N, Ts, ctypes.byref(params) if params else None, *[_P(v) for v in a], _P(x), _P(u),
                              ctypes.byref(info), _P(dux), _P(dpi))
        return (st, info, dux, dpi) if want_step else (st, info)

    def batch(self, N, Ts, x0, yref, yref_e, x, u, n_rti=1, params=None, cond_N=0):
        """cond_N: partial condensing to cond_N stages before the interior-point solve (the reference's qp_cond_N)."""
        n = x0.shape[0]
        status, qp_iter = np.zeros(n, np.int32), np.zeros(n, np.int32)
        if cond_N and cond_N < N:
            self.lib.cfo_batch_pcond(N, Ts, ctypes.byref(params) if params else None, cond_N, n_rti, n, _P(x0), _P(yref),
                                     _P(yref_e), _P(x), _P(u), _I(status), _I(qp_iter))
        else:
            self.lib.cfo_batch(N, Ts, ctypes.byref(params) if params else None, n_rti, n, _P(x0), _P(yref), _P(yref_e),
                               _P(x), _P(u), _I(status), _I(qp_iter))
        return status, qp_iter

    def rti_pcond(self, N, Ts, cond_N, x0, yref, yref_e, x, u, params=None):
        info = CfoInfo()
        a = [np.ascontiguousarray(v, float) for v in (x0, yref, yref_e)]
        st = self.lib.cfo_rti_pcond(N, Ts, ctypes.byref(params) if params else None, cond_N, *[_P(v) for v in a], _P(x), _P(u),
                                    ctypes.byref(info))
        return st, info


def ref_available(model="crazyflie"):
    return os.path.exists(os.path.join(HERE, "_ref", "libcfref.so" if model == "crazyflie" else f"libcfref_{model}.so"))


class Ref:
    """The reference's own implementation (needs oracle/_ref/libcfref.so; model="pendulum": the same harness built for
    the second model, oracle/_ref/libcfref_pendulum.so -- sizes in .nx / .nu, only solver() / RefSolver.rti are meant
    for it)."""

    def __init__(self, model="crazyflie"):
        path = os.path.join(HERE, "_ref", "libcfref.so" if model == "crazyflie" else f"libcfref_{model}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = ctypes.CDLL(path)
        nx, nu = ctypes.c_int(), ctypes.c_int()
        L.cfref_dims(ctypes.byref(nx), ctypes.byref(nu))
        self.nx, self.nu = nx.value, nu.value
        L.cfref_create.restype = ctypes.c_void_p
        L.cfref_create.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int]
        L.cfref_destroy.argtypes = [ctypes.c_void_p]
        L.cfref_rti.restype = ctypes.c_int
        L.cfref_rti.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp]
        L.cfref_get_qp.argtypes = [ctypes.c_void_p] + [_dp] * 7
        L.cfref_get_ipm_stat.restype = ctypes.c_int
        L.cfref_get_ipm_stat.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, _ip]
        L.cfref_set_weights.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.cfref_set_input_bounds.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.cfref_set_input_bounds_stage0.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.cfref_set_input_bounds_at.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp, _dp]
        L.cfref_create_dt.restype = ctypes.c_void_p
        L.cfref_create_dt.argtypes = [ctypes.c_int, _dp, ctypes.c_int]
        L.cfref_rti_split.restype = ctypes.c_int
        L.cfref_rti_split.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip]
        L.cfref_set_opt_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.cfref_get_multipliers.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.cfref_set_W_at.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
        L.cfref_get_stat.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
        L.cfref_batch.restype = ctypes.c_int
        L.cfref_batch.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  _dp, _dp, _dp, _dp, _dp, _ip, _ip, _dp]

    def solver(self, N=50, Ts=0.015, cond_N=0, dt=None):
        """dt: one time step per shooting interval (create_with_discretization) instead of the uniform Ts."""
        return RefSolver(self, N, Ts, cond_N, dt)

    def batch(self, N, Ts, x0, yref, yref_e, x, u, n_rti=1, nthreads=1, cond_N=0):
        n = x0.shape[0]
        status, qp_iter, t = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n)
        fail = self.lib.cfref_batch(N, Ts, cond_N, nthreads, n_rti, n, _P(x0), _P(yref), _P(yref_e), _P(x), _P(u),
                                    _I(status), _I(qp_iter), _P(t))
        if fail:
            raise RuntimeError("reference solver creation failed")
        return status, qp_iter, t


class RefSolver:
    def __init__(self, ref, N, Ts, cond_N, dt=None):
        self.lib, self.N, self.Ts = ref.lib, N, Ts
        if dt is None:
            self.h = self.lib.cfref_create(N, Ts, cond_N)
        else:
            dt = np.ascontiguousarray(dt, float)
            assert dt.size == N
            self.h = self.lib.cfref_create_dt(N, _P(dt), cond_N)
        if not self.h:
            raise RuntimeError("cfref_create failed")

    def close(self):
        if self.h:
            self.lib.cfref_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_weights(self, Wdiag, WNdiag):
        self.lib.cfref_set_weights(self.h, _P(np.ascontiguousarray(Wdiag, float)), _P(np.ascontiguousarray(WNdiag, float)))

    def set_input_bounds(self, lbu, ubu):
        self.lib.cfref_set_input_bounds(self.h, _P(np.ascontiguousarray(lbu, float)), _P(np.ascontiguousarray(ubu, float)))

    def set_input_bounds_at(self, stage, lbu, ubu):
        self.lib.cfref_set_input_bounds_at(self.h, int(stage), _P(np.ascontiguousarray(lbu, float)), _P(np.ascontiguousarray(ubu, float)))

    def set_input_bounds_stage0(self, lbu0, ubu0):
        self.lib.cfref_set_input_bounds_stage0(self.h, _P(np.ascontiguousarray(lbu0, float)), _P(np.ascontiguousarray(ubu0, float)))

    def set_opt_int(self, name, value):
        """ocp_nlp_solver_opts_set with an int value ("qp_iter_max", ...)."""
        self.lib.cfref_set_opt_int(self.h, name.encode(), int(value))

    def set_W_at(self, stage, W):
        """Full weight matrix of one stage (column-major; symmetric, so the order does not matter)."""
        self.lib.cfref_set_W_at(self.h, int(stage), _P(np.ascontiguousarray(W, float)))

    def multipliers(self):
        """(pi [N,13], lam0 [2,17], lam [N-1,2,4]) of the iterate after the last solve; lam = [lower | upper]."""
        N = self.N
        pi, lam = np.zeros((N, NX)), np.zeros(2 * NV + 2 * NU * (N - 1))
        self.lib.cfref_get_multipliers(self.h, _P(pi), _P(lam))
        return pi, lam[:2 * NV].reshape(2, NV), lam[2 * NV:].reshape(N - 1, 2, NU)

    def stat_double(self, name):
        v = ctypes.c_double()
        self.lib.cfref_get_stat(self.h, name.encode(), ctypes.byref(v))
        return v.value

    def stat_int(self, name):
        v = ctypes.c_int()
        self.lib.cfref_get_stat(self.h, name.encode(), ctypes.byref(v))
        return v.value

    def rti(self, x0, yref, yref_e, x, u):
        """One RTI step; x,u updated in place. Returns (status, qp_iter, qp_status, times[5])."""
        qi, qs, t = ctypes.c_int(), ctypes.c_int(), np.zeros(5)
        a = [np.ascontiguousarray(v, float) for v in (x0, yref, yref_e)]
        st = self.lib.cfref_rti(self.h, *[_P(v) for v in a], _P(x), _P(u), ctypes.byref(qi), ctypes.byref(qs), _P(t))
        return st, qi.value, qs.value, t

    def rti_split(self, x0_prep, x0_fb, yref, yref_e, x, u):
        """rti_phase 1 with x0_prep, then rti_phase 2 with x0_fb; x,u updated in place. Returns (status, qp_iter, qp_status)."""
        qi, qs = ctypes.c_int(), ctypes.c_int()
        a = [np.ascontiguousarray(v, float) for v in (x0_prep, x0_fb, yref, yref_e)]
        st = self.lib.cfref_rti_split(self.h, *[_P(v) for v in a], _P(x), _P(u), ctypes.byref(qi), ctypes.byref(qs))
        return st, qi.value, qs.value

    def qp(self):
        N = self.N
        BAbt, b = np.zeros((N, NV, NX)), np.zeros((N, NX))
        rqz, dl, du = np.zeros(N * NV + NX), np.zeros(NV + NU * (N - 1)), np.zeros(NV + NU * (N - 1))
        dux, dpi = np.zeros(N * NV + NX), np.zeros(N * NX)
        self.lib.cfref_get_qp(self.h, _P(BAbt), _P(b), _P(rqz), _P(dl), _P(du), _P(dux), _P(dpi))
        return dict(BAbt=BAbt, b=b, rqz=rqz, d_lb=dl, d_ub=du, dux=dux, dpi=dpi)

    def ipm_stat(self):
        m = ctypes.c_int()
        buf = np.zeros(64 * 32)
        rows = self.lib.cfref_get_ipm_stat(self.h, _P(buf), 64, ctypes.byref(m))
        return buf[: rows * m.value].reshape(rows, m.value)
