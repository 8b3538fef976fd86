"""crazyflie_nmpc_b200 -- batched real-time-iteration NMPC for the Crazyflie OCP on B200.

Host-side Python mirror of the batch C-ABI in include/cfnmpc.h (which itself mirrors the
reference's per-tick call sequence, crazyflie_controller/src/acados_mpc.cpp:581-625).
All compute happens in the CUDA library crazyflie_nmpc_b200/libcfnmpc.so; there is no CPU
path -- if the library or a GPU is missing, construction raises.
"""
import ctypes
import os

import numpy as np

from . import workloads  # noqa: F401

NX, NU, NY = 13, 4, 17
_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("CFNMPC_LIB") or os.path.join(_PKG, "libcfnmpc.so")   # CFNMPC_LIB: A/B experiments only
_lib = None

# acados status codes, acados/acados/utils/types.h:75-83
ACADOS_SUCCESS, ACADOS_NAN_DETECTED, ACADOS_MAXITER, ACADOS_MINSTEP, ACADOS_QP_FAILURE = 0, 1, 2, 3, 4


# enum order of the node's `policy` (crazyflie_controller/src/acados_mpc.cpp:129-133)
POLICY_REGULATION, POLICY_TRACKING, POLICY_HOLD = 0, 1, 2


class CfnmpcError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once). Raises if it has not been built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise CfnmpcError(f"{_LIB_PATH} is missing: build it with `python -m crazyflie_nmpc_b200.build` "
                              "(nvcc, sm_100a). There is no CPU implementation of the solve path.")
        L = ctypes.CDLL(_LIB_PATH)
        vp, cp, ci, cd = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_double
        L.cfnmpc_batch_create.argtypes = [ci, ci, cd, ci, ctypes.POINTER(vp)]
        L.cfnmpc_batch_destroy.argtypes = [vp]
        L.cfnmpc_batch_set_stream.argtypes = [vp, vp]
        L.cfnmpc_batch_set.argtypes = [vp, cp, vp, ci]
        L.cfnmpc_batch_clear.argtypes = [vp, cp]
        L.cfnmpc_batch_set_option.argtypes = [vp, cp, ci]
        L.cfnmpc_batch_solve.argtypes = [vp, ci]
        L.cfnmpc_batch_sync.argtypes = [vp]
        L.cfnmpc_batch_prepare.argtypes = [vp]
        L.cfnmpc_batch_feedback.argtypes = [vp]
        L.cfnmpc_batch_solve_from_host.argtypes = [vp, vp, vp, vp, ci]
        L.cfnmpc_batch_get.argtypes = [vp, cp, ci, vp, ci]
        L.cfnmpc_batch_device_ptr.argtypes = [vp, cp, ctypes.POINTER(vp)]
        L.cfnmpc_batch_info.argtypes = [vp, cp, ctypes.POINTER(ctypes.c_longlong)]
        L.cfnmpc_batch_last_solve_ms.argtypes = [vp, ctypes.POINTER(cd)]
        L.cfnmpc_batch_last_phase_ms.argtypes = [vp, ctypes.POINTER(cd)]
        L.cfnmpc_debug_scratch.argtypes = [vp, vp, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_longlong)]
        L.cfnmpc_debug_max_ipm_iter.argtypes = [vp, ci]
        L.cfnmpc_debug_pass_cycles.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
        L.cfnmpc_batch_set_trajectory.argtypes = [vp, vp, ci, ci]
        L.cfnmpc_batch_update_reference.argtypes = [vp]
        L.cfnmpc_batch_commands.argtypes = [vp, ci]
        L.cfnmpc_batch_tick.argtypes = [vp, ci]
        L.cfnmpc_batch_plant_step.argtypes = [vp, cd, ci, ci]
        L.cfnmpc_sim_create.argtypes = [ci, ci, ctypes.POINTER(vp)]
        L.cfnmpc_sim_destroy.argtypes = [vp]
        L.cfnmpc_sim_set_stream.argtypes = [vp, vp]
        L.cfnmpc_sim_opts_set.argtypes = [vp, cp, ci]
        L.cfnmpc_sim_set.argtypes = [vp, cp, vp, ci]
        L.cfnmpc_sim_solve.argtypes = [vp]
        L.cfnmpc_sim_get.argtypes = [vp, cp, vp, ci]
        L.cfnmpc_sim_launches.argtypes = [vp, ctypes.POINTER(ctypes.c_longlong)]
        L.cfnmpc_multi_create.argtypes = [ci, ci, cd, ci, ctypes.POINTER(ci), ctypes.POINTER(vp)]
        L.cfnmpc_multi_destroy.argtypes = [vp]
        L.cfnmpc_multi_num_shards.argtypes = [vp]
        L.cfnmpc_multi_shard.argtypes = [vp, ci, ctypes.POINTER(vp), ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
        L.cfnmpc_multi_set.argtypes = [vp, cp, vp]
        L.cfnmpc_multi_set_option.argtypes = [vp, cp, ci]
        L.cfnmpc_multi_set_trajectory.argtypes = [vp, vp, ci]
        L.cfnmpc_multi_solve.argtypes = [vp, ci]
        L.cfnmpc_multi_solve_from_host.argtypes = [vp, vp, vp, vp, ci]
        L.cfnmpc_multi_tick.argtypes = [vp, ci]
        L.cfnmpc_multi_sync.argtypes = [vp]
        L.cfnmpc_multi_get.argtypes = [vp, cp, ci, vp]
        L.cfnmpc_multi_last_solve_ms.argtypes = [vp, ctypes.POINTER(cd)]
        L.cfnmpc_multi_last_error.restype = cp
        L.cfnmpc_measure_fp64_peak.argtypes = [ci, ctypes.POINTER(cd)]
        L.cfnmpc_last_error.restype = cp
        L.cfnmpc_version.restype = cp
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CfnmpcError(f"cfnmpc error {rc}: {lib().cfnmpc_last_error().decode()}")


def _ptr(a):
    """(address, on_device, keepalive) for a numpy array or a torch tensor."""
    if hasattr(a, "data_ptr"):  # torch tensor (duck-typed; torch is not a dependency of this module)
        if not a.is_contiguous():
            a = a.contiguous()
        return a.data_ptr(), 1 if a.is_cuda else 0, a
    a = np.ascontiguousarray(a)
    return a.ctypes.data, 0, a


def _torch_dtype_is(t, is_int):
    """float64 (or int32) torch tensor?  Compared by name: torch is not a dependency of this module."""
    return str(t.dtype) == ("torch.int32" if is_int else "torch.float64")


def measure_fp64_peak(device=0):
    """Measured fp64 FMA throughput of the device in TFLOP/s (roofline denominator of bench.py)."""
    v = ctypes.c_double()
    _check(lib().cfnmpc_measure_fp64_peak(int(device), ctypes.byref(v)))
    return v.value


class BatchSolver:
    """B independent Crazyflie NMPC instances advanced together, one warp per instance.

    set(field, array)   "x0" [B,13] | "yref" [B,N,17] | "yref_e" [B,13] | "x" [B,N+1,13] | "u" [B,N,4]
                        | "W" [17] | "W_e" [13] | "lbu" | "ubu" | "lbu0" | "ubu0" [4]   (solver-wide)
                        | "W_batch" [B,17] | "W_e_batch" [B,13] | "lbu_batch" | "ubu_batch" | "lbu0_batch"
                        | "ubu0_batch" [B,4]   (per instance; clear(field) returns to the solver-wide value)
                        Device-resident sources (CUDA tensors) are copied on the solver's stream: call
                        set_stream(torch.cuda.current_stream().cuda_stream) first, or synchronise the producing stream.
    solve(n_rti=1)      enqueue RTI steps (asynchronous)
    get(field, stage)   "u"/"x" at a stage, "u_all", "x_all", "status", "qp_iter", "qp_status", "flags", "res"
    """
    _SHAPES = {"status": np.int32, "qp_iter": np.int32, "qp_status": np.int32, "flags": np.int32}

    def __init__(self, batch, N=50, Ts=0.015, device=0):
        self._h = ctypes.c_void_p()
        self.B, self.N, self.Ts, self.device = int(batch), int(N), float(Ts), int(device)
        _check(lib().cfnmpc_batch_create(self.B, self.N, self.Ts, self.device, ctypes.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().cfnmpc_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_handle):
        """Run on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0/None = own stream."""
        _check(lib().cfnmpc_batch_set_stream(self._h, ctypes.c_void_p(cuda_stream_handle or 0)))

    def _expected(self, field):
        B, N = self.B, self.N
        return {"x0": (B, NX), "yref": (B, N, NY), "yref_e": (B, NX), "x": (B, N + 1, NX), "u": (B, N, NU),
                "W": (NY,), "W_e": (NX,), "lbu": (NU,), "ubu": (NU,), "lbu0": (NU,), "ubu0": (NU,),
                "W_batch": (B, NY), "W_e_batch": (B, NX), "lbu_batch": (B, NU), "ubu_batch": (B, NU),
                "lbu0_batch": (B, NU), "ubu0_batch": (B, NU), "setpoint": (B, 3), "uss": (1,),
                "policy": (B,), "traj_iter": (B,), "time_steps": (N,), "bounds_stage": (N, 8), "W_stage": (N + 1, NY),
                "W_dense_table": (N + 1, NY, NY)}.get(field)

    def set_option(self, option, value):
        """"lin_res_check" (0/1: the reference's linear-system residual diagnostics -> "flags"), "max_ipm_iter"."""
        _check(lib().cfnmpc_batch_set_option(self._h, option.encode(), int(value)))
        return self

    def clear(self, field):
        _check(lib().cfnmpc_batch_clear(self._h, field.encode()))
        return self

    def set(self, field, a):
        if self._expected(field) is None:
            raise CfnmpcError(f"unknown field '{field}'")
        n = int(np.prod(self._expected(field)))
        is_int = field in ("policy", "traj_iter")
        if hasattr(a, "data_ptr"):
            if a.numel() != n or not _torch_dtype_is(a, is_int):
                raise CfnmpcError(f"'{field}' needs {n} {'int32' if is_int else 'float64'} values, got {a.numel()} of {a.dtype}")
        else:
            a = np.ascontiguousarray(a, dtype=np.int32 if is_int else np.float64)
            if a.size != n:
                raise CfnmpcError(f"'{field}' needs {n} values, got {a.size}")
        p, dev, keep = _ptr(a)
        _check(lib().cfnmpc_batch_set(self._h, field.encode(), ctypes.c_void_p(p), dev))
        if not dev and not (hasattr(a, "is_pinned") and a.is_pinned()):
            self.sync()  # pageable host memory: make the copy complete before the caller may reuse it
        return self

    def set_problem(self, w):
        """Load a workload dict as produced by crazyflie_nmpc_b200.workloads."""
        self.set("x0", w["x0"]).set("yref", w["yref"]).set("yref_e", w["yref_e"])
        self.set("x", w["x_init"]).set("u", w["u_init"])
        return self

    def solve(self, n_rti=1):
        _check(lib().cfnmpc_batch_solve(self._h, int(n_rti)))
        return self

    def prepare(self):
        """Preparation phase of a split real-time iteration (the reference's rti_phase 1): linearise around the iterate."""
        _check(lib().cfnmpc_batch_prepare(self._h))
        return self

    def feedback(self):
        """Feedback phase (rti_phase 2): take the current "x0", solve the QP, update the iterate."""
        _check(lib().cfnmpc_batch_feedback(self._h))
        return self

    def solve_from_host(self, x0, yref, yref_e, n_chunks=4):
        """One tick from host buffers (numpy arrays or pinned torch CPU tensors): chunked upload overlapped with the solve."""
        ptrs = []
        for a, n in ((x0, self.B * NX), (yref, self.B * self.N * NY), (yref_e, self.B * NX)):
            if hasattr(a, "data_ptr"):
                if a.is_cuda or a.numel() != n or a.element_size() != 8:
                    raise CfnmpcError("solve_from_host needs float64 host tensors of the batch shapes")
            else:
                a = np.ascontiguousarray(a, dtype=np.float64)
                if a.size != n:
                    raise CfnmpcError("solve_from_host: wrong array size")
            ptrs.append(_ptr(a))
        _check(lib().cfnmpc_batch_solve_from_host(self._h, ctypes.c_void_p(ptrs[0][0]), ctypes.c_void_p(ptrs[1][0]),
                                                  ctypes.c_void_p(ptrs[2][0]), int(n_chunks)))
        if not all(hasattr(p[2], "is_pinned") and p[2].is_pinned() for p in ptrs):
            self.sync()  # pageable host memory: the copies must be complete before the caller may reuse the buffers
        return self

    def sync(self):
        _check(lib().cfnmpc_batch_sync(self._h))

    def get(self, field, stage=0, out=None):
        B, N = self.B, self.N
        shapes = {"u": ((B, NU), np.float64), "x": ((B, NX), np.float64), "u_all": ((B, N, NU), np.float64),
                  "x_all": ((B, N + 1, NX), np.float64), "status": ((B,), np.int32), "qp_iter": ((B,), np.int32),
                  "qp_status": ((B,), np.int32), "flags": ((B,), np.int32), "res": ((B, 4), np.float64),
                  "policy": ((B,), np.int32), "traj_iter": ((B,), np.int32), "motors": ((B, NU), np.int32),
                  "euler": ((B, 3), np.float64), "twist": ((B, 4), np.float64), "x0": ((B, NX), np.float64),
                  "yref": ((B, N, NY), np.float64), "yref_e": ((B, NX), np.float64), "setpoint": ((B, 3), np.float64),
                  # option "multipliers": per stage, see include/cfnmpc.h
                  "pi": ((B, NX), np.float64), "lam": ((B, 8), np.float64), "t": ((B, 8), np.float64),
                  "lam_x0": ((B, NX), np.float64), "pi_all": ((B, N, NX), np.float64), "lam_all": ((B, N, 8), np.float64),
                  "t_all": ((B, N, 8), np.float64)}
        if field not in shapes:
            raise CfnmpcError(f"unknown field '{field}'")
        shape, dt = shapes[field]
        if out is None:
            out = np.empty(shape, dt)
        else:   # a caller-provided buffer is written by a raw device copy: size, type and layout must be right
            n = int(np.prod(shape))
            if hasattr(out, "data_ptr"):
                if out.numel() != n or not _torch_dtype_is(out, dt is np.int32) or not out.is_contiguous():
                    raise CfnmpcError(f"get('{field}', out=...): need a contiguous {np.dtype(dt).name} tensor of {n} elements")
            elif not (isinstance(out, np.ndarray) and out.size == n and out.dtype == dt and out.flags.c_contiguous):
                raise CfnmpcError(f"get('{field}', out=...): need a C-contiguous {np.dtype(dt).name} array of {n} elements")
        p, dev, keep = _ptr(out)
        _check(lib().cfnmpc_batch_get(self._h, field.encode(), int(stage), ctypes.c_void_p(p), dev))
        return out

    # ---- closed-loop driver (the rest of NMPC::iteration, acados_mpc.cpp:430-516,619-670)
    def set_trajectory(self, table):
        """Trajectory table [rows,17] (crazyflie_controller/traj/*.txt format) for the Tracking / Hold policies."""
        if hasattr(table, "data_ptr"):
            if table.dim() != 2 or table.shape[1] != NY or not _torch_dtype_is(table, False):
                raise CfnmpcError("trajectory table must be a float64 tensor [rows, 17]")
            rows = table.shape[0]
        else:
            table = np.ascontiguousarray(table, dtype=np.float64)
            rows = table.shape[0]
            if table.ndim != 2 or table.shape[1] != NY:
                raise CfnmpcError("trajectory table must be [rows, 17]")
        p, dev, keep = _ptr(table)
        _check(lib().cfnmpc_batch_set_trajectory(self._h, ctypes.c_void_p(p), int(rows), dev))
        if not dev:
            self.sync()
        return self

    def update_reference(self):
        _check(lib().cfnmpc_batch_update_reference(self._h))
        return self

    def commands(self, motors_from_u1=False):
        _check(lib().cfnmpc_batch_commands(self._h, 1 if motors_from_u1 else 0))
        return self

    def tick(self, motors_from_u1=False):
        """update_reference + one RTI step + commands."""
        _check(lib().cfnmpc_batch_tick(self._h, 1 if motors_from_u1 else 0))
        return self

    def plant_step(self, dt, n_steps=1, truncated_motors=False):
        """x0 <- ERK4(x0, u_0 or int32 motors, dt): simulated vehicle for closed-loop runs on the device."""
        _check(lib().cfnmpc_batch_plant_step(self._h, float(dt), int(n_steps), 1 if truncated_motors else 0))
        return self

    def device_ptr(self, field):
        p = ctypes.c_void_p()
        _check(lib().cfnmpc_batch_device_ptr(self._h, field.encode(), ctypes.byref(p)))
        return p.value

    def info(self, what):
        v = ctypes.c_longlong()
        _check(lib().cfnmpc_batch_info(self._h, what.encode(), ctypes.byref(v)))
        return v.value

    def last_solve_ms(self):
        v = ctypes.c_double()
        _check(lib().cfnmpc_batch_last_solve_ms(self._h, ctypes.byref(v)))
        return v.value

    def last_phase_ms(self):
        """(preparation kernel ms, feedback kernel ms) of the last step; (0, total) when it ran as one fused kernel."""
        v = (ctypes.c_double * 2)()
        _check(lib().cfnmpc_batch_last_phase_ms(self._h, v))
        return v[0], v[1]

    # ---- test hooks
    def debug_scratch(self):
        n, offs = ctypes.c_size_t(), (ctypes.c_longlong * 12)()
        _check(lib().cfnmpc_debug_scratch(self._h, None, 0, ctypes.byref(n), offs))
        buf = np.empty(n.value)
        _check(lib().cfnmpc_debug_scratch(self._h, ctypes.c_void_p(buf.ctypes.data), n.value, None, None))
        names = ["total", "blk_stride", "b_m", "b_lu", "b_px", "r_ux", "r_pi", "r_rq", "r_b", "r_resg", "r_dux", "r_d"]
        return buf, dict(zip(names, list(offs)))

    def debug_pass_cycles(self, read=True):
        """Per-pass warp cycles / call counts since the last read (first call enables the counters)."""
        buf = (ctypes.c_ulonglong * 12)()
        _check(lib().cfnmpc_debug_pass_cycles(self._h, buf if read else None))
        names = ["linearize", "residual_factorize", "forward", "backward_rhs", "mu_aff", "update"]
        return {n: (buf[2 * i], buf[2 * i + 1]) for i, n in enumerate(names)}

    def debug_max_ipm_iter(self, n):
        _check(lib().cfnmpc_debug_max_ipm_iter(self._h, int(n)))


class SimBatch:
    """B independent state predictions x+ = ERK4(x, u, T): the estimator node's delay compensation
    (crazyflie_controller/src/acados_estimator.cpp:573-593), mirror of cfnmpc_sim_* in include/cfnmpc.h."""

    def __init__(self, batch, device=0, num_steps=1, sens_forw=False):
        self._h = ctypes.c_void_p()
        self.B = int(batch)
        _check(lib().cfnmpc_sim_create(self.B, int(device), ctypes.byref(self._h)))
        self.opts_set("num_steps", num_steps)
        self.opts_set("sens_forw", 1 if sens_forw else 0)
        self.sens_forw = bool(sens_forw)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().cfnmpc_sim_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream_handle):
        _check(lib().cfnmpc_sim_set_stream(self._h, ctypes.c_void_p(cuda_stream_handle or 0)))

    def opts_set(self, field, value):
        _check(lib().cfnmpc_sim_opts_set(self._h, field.encode(), int(value)))
        if field == "sens_forw":
            self.sens_forw = bool(value)
        return self

    def set(self, field, a):
        n = {"x": self.B * NX, "u": self.B * NU, "T": 1, "T_batch": self.B}.get(field)
        if n is None:
            raise CfnmpcError(f"unknown field '{field}'")
        if hasattr(a, "data_ptr"):
            if a.numel() != n or a.element_size() != 8:
                raise CfnmpcError(f"'{field}' needs {n} float64 values")
        else:
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.size != n:
                raise CfnmpcError(f"'{field}' needs {n} float64 values, got {a.size}")
        p, dev, keep = _ptr(a)
        _check(lib().cfnmpc_sim_set(self._h, field.encode(), ctypes.c_void_p(p), dev))
        if not dev:
            self.get("xn")  # pageable host source: drain the stream before the caller may reuse the buffer
        return self

    def solve(self):
        _check(lib().cfnmpc_sim_solve(self._h))
        return self

    def get(self, field, out=None):
        shape = {"xn": (self.B, NX), "S_forw": (self.B, NX + NU, NX)}.get(field)
        if shape is None:
            raise CfnmpcError(f"unknown field '{field}'")
        if out is None:
            out = np.empty(shape)
        p, dev, keep = _ptr(out)
        _check(lib().cfnmpc_sim_get(self._h, field.encode(), ctypes.c_void_p(p), dev))
        return out

    def launches(self):
        v = ctypes.c_longlong()
        _check(lib().cfnmpc_sim_launches(self._h, ctypes.byref(v)))
        return v.value


class MultiBatchSolver:
    """B instances sharded over several GPUs of one node behind one handle (mirror of cfnmpc_multi_* in include/cfnmpc.h):
    contiguous shards, one host thread and stream per device, no exchange between shards.  Host (numpy) arrays only."""
    _INT = ("status", "qp_iter", "qp_status", "flags", "policy", "traj_iter", "motors")

    def __init__(self, batch, N=50, Ts=0.015, devices=(0,)):
        self._h = ctypes.c_void_p()
        self.B, self.N = int(batch), int(N)
        devs = (ctypes.c_int * len(devices))(*devices)
        self._check(lib().cfnmpc_multi_create(self.B, self.N, float(Ts), len(devices), devs, ctypes.byref(self._h)))

    @staticmethod
    def _check(rc):
        if rc != 0:
            raise CfnmpcError(f"cfnmpc_multi error {rc}: {lib().cfnmpc_multi_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().cfnmpc_multi_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self):
        out = []
        for i in range(lib().cfnmpc_multi_num_shards(self._h)):
            d, f, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            self._check(lib().cfnmpc_multi_shard(self._h, i, None, ctypes.byref(d), ctypes.byref(f), ctypes.byref(c)))
            out.append((d.value, f.value, c.value))
        return out

    def set(self, field, a):
        a = np.ascontiguousarray(a, dtype=np.int32 if field in self._INT else np.float64)
        self._check(lib().cfnmpc_multi_set(self._h, field.encode(), ctypes.c_void_p(a.ctypes.data)))
        return self

    def set_option(self, option, value):
        self._check(lib().cfnmpc_multi_set_option(self._h, option.encode(), int(value)))
        return self

    def set_problem(self, w):
        for k, f in (("x0", "x0"), ("yref", "yref"), ("yref_e", "yref_e"), ("x_init", "x"), ("u_init", "u")):
            self.set(f, w[k])
        return self

    def solve(self, n_rti=1):
        self._check(lib().cfnmpc_multi_solve(self._h, int(n_rti)))
        return self

    def solve_from_host(self, x0, yref, yref_e, n_chunks=4):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x0, yref, yref_e)]
        self._check(lib().cfnmpc_multi_solve_from_host(self._h, *[ctypes.c_void_p(v.ctypes.data) for v in a], int(n_chunks)))
        return self

    def sync(self):
        self._check(lib().cfnmpc_multi_sync(self._h))

    def get(self, field, stage=0):
        B, N = self.B, self.N
        shapes = {"u": (B, NU), "x": (B, NX), "u_all": (B, N, NU), "x_all": (B, N + 1, NX), "status": (B,), "qp_iter": (B,),
                  "qp_status": (B,), "flags": (B,), "res": (B, 4), "motors": (B, NU), "euler": (B, 3), "twist": (B, 4)}
        out = np.empty(shapes[field], np.int32 if field in self._INT else np.float64)
        self._check(lib().cfnmpc_multi_get(self._h, field.encode(), int(stage), ctypes.c_void_p(out.ctypes.data)))
        return out

    def last_solve_ms(self):
        v = ctypes.c_double()
        self._check(lib().cfnmpc_multi_last_solve_ms(self._h, ctypes.byref(v)))
        return v.value


class ModelSolver:
    """Batch solver of a generic-model library, crazyflie_nmpc_b200/libcfnmpc_<model>.so: the same kernel sources compiled
    for another generated OCP description (tools/gen_spec.py --model <model>).  Core subset of the batch
    C-ABI (include/cfnmpc.h): set / solve / prepare / feedback / get.  Sizes come from the library (cfnmpc_model_dims)."""

    def __init__(self, model, batch, N=None, Ts=None, device=0):
        path = os.path.join(_PKG, f"libcfnmpc_{model}.so")
        if not os.path.exists(path):
            raise CfnmpcError(f"{path} is missing: build it with `python -m crazyflie_nmpc_b200.build`. There is no CPU implementation.")
        L = self._L = ctypes.CDLL(path)
        L.cfnmpc_last_error.restype = ctypes.c_char_p
        nx, nu, n0, tf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        L.cfnmpc_model_dims(ctypes.byref(nx), ctypes.byref(nu), ctypes.byref(n0), ctypes.byref(tf))
        self.nx, self.nu, self.ny = nx.value, nu.value, nx.value + nu.value
        self.B, self.N = int(batch), int(N if N is not None else n0.value)
        self.Ts = float(Ts if Ts is not None else tf.value / n0.value)
        L.cfnmpc_batch_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
        L.cfnmpc_batch_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int]
        L.cfnmpc_batch_get.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.cfnmpc_batch_set_option.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.cfnmpc_batch_info.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]
        L.cfnmpc_batch_last_solve_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
        for f in ("cfnmpc_batch_solve",):
            getattr(L, f).argtypes = [ctypes.c_void_p, ctypes.c_int]
        for f in ("cfnmpc_batch_prepare", "cfnmpc_batch_feedback", "cfnmpc_batch_sync", "cfnmpc_batch_destroy"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        self._h = ctypes.c_void_p()
        self._check(L.cfnmpc_batch_create(self.B, self.N, self.Ts, int(device), ctypes.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            raise CfnmpcError(f"cfnmpc error {rc}: {self._L.cfnmpc_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.cfnmpc_batch_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _shape(self, field):
        B, N, nx, nu, ny = self.B, self.N, self.nx, self.nu, self.ny
        return {"x0": (B, nx), "yref": (B, N, ny), "yref_e": (B, nx), "x": (B, N + 1, nx), "u": (B, N, nu), "W": (ny,), "W_e": (nx,),
                "lbu": (nu,), "ubu": (nu,), "lbu0": (nu,), "ubu0": (nu,), "time_steps": (N,), "bounds_stage": (N, 2 * nu),
                "W_stage": (N + 1, ny)}.get(field)

    def set(self, field, a):
        shape = self._shape(field)
        if shape is None:
            raise CfnmpcError(f"unknown field '{field}'")
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.size != int(np.prod(shape)):
            raise CfnmpcError(f"'{field}' needs {int(np.prod(shape))} values, got {a.size}")
        self._check(self._L.cfnmpc_batch_set(self._h, field.encode(), ctypes.c_void_p(a.ctypes.data), 0))
        self._check(self._L.cfnmpc_batch_sync(self._h))
        return self

    def set_problem(self, w):
        for k, f in (("x0", "x0"), ("yref", "yref"), ("yref_e", "yref_e"), ("x_init", "x"), ("u_init", "u")):
            self.set(f, w[k])
        return self

    def set_option(self, option, value):
        self._check(self._L.cfnmpc_batch_set_option(self._h, option.encode(), int(value)))
        return self

    def solve(self, n_rti=1):
        self._check(self._L.cfnmpc_batch_solve(self._h, int(n_rti)))
        return self

    def prepare(self):
        self._check(self._L.cfnmpc_batch_prepare(self._h))
        return self

    def feedback(self):
        self._check(self._L.cfnmpc_batch_feedback(self._h))
        return self

    def get(self, field, stage=0):
        B, N = self.B, self.N
        shapes = {"x": ((B, self.nx), np.float64), "u": ((B, self.nu), np.float64), "x_all": ((B, N + 1, self.nx), np.float64),
                  "u_all": ((B, N, self.nu), np.float64), "status": ((B,), np.int32), "qp_iter": ((B,), np.int32),
                  "qp_status": ((B,), np.int32), "flags": ((B,), np.int32), "res": ((B, 4), np.float64)}
        if field not in shapes:
            raise CfnmpcError(f"unknown field '{field}'")
        out = np.empty(*shapes[field])
        self._check(self._L.cfnmpc_batch_get(self._h, field.encode(), int(stage), ctypes.c_void_p(out.ctypes.data), 0))
        return out

    def last_solve_ms(self):
        v = ctypes.c_double()
        self._check(self._L.cfnmpc_batch_last_solve_ms(self._h, ctypes.byref(v)))
        return v.value

    def info(self, what):
        v = ctypes.c_longlong()
        self._check(self._L.cfnmpc_batch_info(self._h, what.encode(), ctypes.byref(v)))
        return v.value
