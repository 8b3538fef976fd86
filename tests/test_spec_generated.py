"""Spec-driven specialisation: the model / weight header the kernels compile against is generated from
the reference's own OCP description by tools/gen_spec.py.  Checks: the committed header is what the generator produces
from /root/reference (where that tree exists), and the generated model agrees with the hand-derived one and with the
oracle's (independent) restatement."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "crazyflie_nmpc_b200", "csrc", "cf_spec_generated.h")


@pytest.mark.skipif(not os.path.isdir("/root/reference/crazyflie_controller"), reason="needs the reference tree")
def test_committed_header_is_up_to_date():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_spec.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_spec_constants_match_the_reference_description():
    src = open(HDR).read()
    assert "#define CF_SPEC_N 50" in src and "#define CF_SPEC_TF 0.75" in src
    assert "W[17] = {120.0, 100.0, 100.0, 0.001, 0.001, 0.001, 0.001, 0.7, 1.0, 4.0, 1e-05, 1e-05, 10.0, 0.06, 0.06, 0.06, 0.06}" in src
    assert "ubu[4] = {22.0, 22.0, 22.0, 22.0}" in src and "nlp_solver_type = SQP_RTI" in src


@pytest.fixture(scope="module")
def model_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("specmodel")
    src = d / "m.cpp"
    src.write_text('''
#define CF_SIMT_EMU 1
#include "cf_model.h"
namespace cfemu { int lane() { return 0; } void barrier(int) {} double exch(double v, int, int) { return v; }
int atomic_add(int *p, int v) { int o = *p; *p += v; return o; } void bulk_expect(uint64_t *, int, int) {}
void bulk_g2s(void *, const void *, int, uint64_t *, int) {} void bulk_wait(uint64_t *, unsigned, int) {}
void bulk_s2g(void *, const void *, int, int) {} void bulk_s2g_wait(int, int) {} void dmma(double &, double &, double, double, int) {} }
extern "C" void gen(const double *x, const double *u, const double *d, int j, double *f, double *o, double *c)
{ cf_ode(x, u, f); cf_jvp_x(x, u, d, o); for (int i = 0; i < 13; i++) c[i] = 0; cf_add_ju_col(x, u, j, c); }
extern "C" void hand(const double *x, const double *u, const double *d, int j, double *f, double *o, double *c)
{ cf_ode_hand(x, u, f); cf_jvp_x_hand(x, d, o); for (int i = 0; i < 13; i++) c[i] = 0; cf_add_ju_col_hand(u, j, c); }
''')
    so = d / "m.so"
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-I", os.path.join(ROOT, "crazyflie_nmpc_b200", "csrc"), str(src), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def test_generated_model_matches_hand_derived_and_oracle(model_lib, port):
    dp = ctypes.POINTER(ctypes.c_double)
    P = lambda a: a.ctypes.data_as(dp)
    rng = np.random.default_rng(0)
    for _ in range(50):
        x, u, d = rng.normal(size=13), rng.uniform(0, 22, 4), rng.normal(size=13)
        j = int(rng.integers(0, 4))
        out = {}
        for name in ("gen", "hand"):
            f, o, c = np.zeros(13), np.zeros(13), np.zeros(13)
            getattr(model_lib, name)(P(x), P(u), P(d), j, P(f), P(o), P(c))
            out[name] = (f, o, c)
        for a, b in zip(out["gen"], out["hand"]):
            assert np.abs(a - b).max() <= 1e-12 * (1 + np.abs(b).max())
        assert np.abs(out["gen"][0] - port.ode(x, u)).max() <= 1e-12 * (1 + np.abs(out["gen"][0]).max())


def test_free_states_have_unit_rows_in_the_linearisation(port):
    """CF_SPEC_NFREE (tools/gen_spec.py: leading states with a zero column of df/dx) is what lets the feedback program of
    cf_rti_warp.h drop their rows of [B';A']: in the oracle's linearisation (the reference's sim_erk forward
    sensitivities) those rows are EXACTLY the unit vectors, whatever the state, on benign and on adversarial instances."""
    from crazyflie_nmpc_b200 import workloads as wl
    src = open(HDR).read()
    assert "#define CF_SPEC_NFREE 3 " in src
    assert "#define CF_SPEC_NFREE 1 " in open(HDR.replace("cf_spec_generated.h", "cf_spec_pendulum.h")).read()
    N = 50
    for w in (wl.hover_batch(4, N, seed=77), wl.helix_batch(4, N, seed=78), wl.adversarial_batch(4, N, seed=79)):
        for i in range(4):
            lin = port.linearize(N, 0.015, w["x0"][i], w["yref"][i], w["yref_e"][i], w["x_init"][i], w["u_init"][i])
            assert (lin["BAbt"][:, 4:7, :] == np.eye(13)[:3][None]).all()
