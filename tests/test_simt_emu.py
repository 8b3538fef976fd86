"""The CUDA warp program (crazyflie_nmpc_b200/csrc/cf_rti_warp.h) compiled for the host and run with its
32 lanes as lock-step fibers (tests/simt_emu/), against the golden vectors and the oracle.

This exercises the kernel SOURCE -- lane mapping, shared-memory staging, barriers, the mbarrier/bulk-copy
protocol (the emulation poisons staged blocks until the warp waits on the right barrier) -- on machines
without a GPU.  It is test infrastructure: the library has no such path.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
TS = 0.015
_dp, _ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def emu():
    subprocess.run([sys.executable, os.path.join(HERE, "simt_emu", "build.py")], check=True)
    L = ctypes.CDLL(os.path.join(HERE, "simt_emu", "libcfemu.so"))
    L.cfemu_rti_batch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int, _dp, _dp, _dp, _dp, _dp,
                                  _ip, _ip, _ip, _ip, _dp, _dp, ctypes.c_int]

    def run(w, N, n_rti=1, params=None):
        B = w["x0"].shape[0]
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it, qs, fl = [np.zeros(B, np.int32) for _ in range(4)]
        res = np.zeros((B, 4))
        P = lambda a: a.ctypes.data_as(_dp)
        I = lambda a: a.ctypes.data_as(_ip)
        par = None if params is None else np.ascontiguousarray(np.concatenate(params), float)
        for _ in range(n_rti):
            L.cfemu_rti_batch(B, N, TS, P(par) if par is not None else None, 0, P(w["x0"]), P(w["yref"]), P(w["yref_e"]),
                              P(x), P(u), I(st), I(it), I(qs), I(fl), P(res), None, 4)
        return dict(x=x, u=u, status=st, qp_iter=it, qp_status=qs, flags=fl, res=res)
    return run


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "crazyflie_rti_golden.npz"))


def batch(gold, name, n=None):
    return {k: np.ascontiguousarray(gold[f"{name}_{k}"][:n]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}


@pytest.mark.parametrize("name,N,n", [("hover", 50, 6), ("helix", 50, 6), ("hover20", 20, 4), ("hover100", 100, 2)])
def test_emulated_kernel_matches_golden(emu, gold, name, N, n):
    w = batch(gold, name, n)
    r = emu(w, N)
    assert (r["status"] == 0).all() and (r["flags"] == 0).all() and (r["qp_status"] == 0).all()
    assert np.abs(r["qp_iter"] - gold[f"{name}_qp_iter"][:n]).max() <= 1
    assert rel_err(r["x"], gold[f"{name}_x"][:n]) < 1e-9 and rel_err(r["u"], gold[f"{name}_u"][:n]) < 1e-9


def test_emulated_kernel_config1_degenerate_first_step(emu, gold):
    """u = 0 start: dphi/du = 0, the first step is decided by the cost alone (SURVEY 'B = 0 at u = 0')."""
    for c in (0, 1, 2):
        w = {k: np.ascontiguousarray(gold[f"cfg1_{c}_{k}"]) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
        r = emu(w, 50)
        assert rel_err(r["x"], gold[f"cfg1_{c}_x1"]) < 1e-9 and rel_err(r["u"], gold[f"cfg1_{c}_u1"]) < 1e-9
    r5 = emu(w, 50, n_rti=5)
    assert rel_err(r5["x"], gold["cfg1_2_x5"]) < 1e-7 and rel_err(r5["u"], gold["cfg1_2_u5"]) < 1e-7


def test_emulated_kernel_runtime_parameters(emu, gold):
    w = batch(gold, "hover", 2)
    r = emu(w, 50, params=(gold["par_W"], gold["par_WN"], gold["par_lbu"], gold["par_ubu"]))
    assert rel_err(r["x"], gold["par_x"][:2]) < 1e-9 and rel_err(r["u"], gold["par_u"][:2]) < 1e-9


def test_emulated_kernel_tiny_horizons(emu, port):
    from crazyflie_nmpc_b200 import workloads as wl
    for N in (1, 2, 3):
        w = wl.hover_batch(3, N, seed=N)
        r = emu(w, N)
        x, u = w["x_init"].copy(), w["u_init"].copy()
        port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
        assert rel_err(r["x"], x) < 1e-9 and rel_err(r["u"], u) < 1e-9
