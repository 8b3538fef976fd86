// Crazyflie quadrotor model: the OCP specification the kernels are specialised for.
//
// Replaces (for this one OCP) the CasADi-generated external functions
// crazyflie_expl_ode_fun / crazyflie_expl_vde_forw that the reference produces from
//   crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py:34-42,85-101
// via acados_template/casadi_function_generation.py:137-151.
// State x = [p(3), q(4: w,x,y,z), v_body(3), omega(3)], input u = 4 motor speeds [krpm].
//
// The functions the kernels call (cf_ode, cf_jvp_x, cf_add_ju_col) and the default weights / bounds / horizon come from
// cf_spec_generated.h, which tools/gen_spec.py derives from those two reference files (sympy, common-subexpression
// elimination).  The hand-derived versions below (*_hand) are kept as an independent cross-check
// (tests/test_spec_generated.py).
#pragma once
#include "cf_simt.h"
// The OCP description the library is built for.  Default: the Crazyflie OCP; another generated header (tools/gen_spec.py
// --model ...) is selected with -DCF_SPEC_HEADER='"cf_spec_pendulum.h"' -- the same kernel sources then compile for its
// sizes: preparation kernel + the dense-stage feedback program of cf_pcond_warp.h.  The hand-tuned
// uncondensed feedback program (cf_rti_warp.h) and the node-specific kernels exist for nx = 13, nu = 4 only.
#ifndef CF_SPEC_HEADER
#define CF_SPEC_HEADER "cf_spec_generated.h"
#endif
#include CF_SPEC_HEADER

#define CF_NX CF_SPEC_NX
#define CF_NU CF_SPEC_NU
#define CF_NV (CF_NX + CF_NU)  // stage variables [u; x]  (acados/ocp_nlp/ocp_nlp_common.h:229)
#define CF_NY CF_NV            // cost output y = [x; u]
#if CF_SPEC_NX == 13 && CF_SPEC_NU == 4
#define CF_CRAZYFLIE 1
#else
#define CF_CRAZYFLIE 0
#endif
static_assert(CF_NV + 1 <= 32, "one row of the (nv + 1) x nx stage matrices per lane");
static_assert(CF_NV % 2 == 1, "nv + 1 rows must be an even number (16-byte aligned columns): pad the model by one variable");
static_assert(CF_NX <= 16, "two 8-column tiles hold the cost-to-go Hessian");

#if CF_CRAZYFLIE
struct CfModel
{
    // export_ode_model.py:34-42
    static constexpr double g0 = 9.8066, mq = 33e-3, Ixx = 1.395e-5, Iyy = 1.395e-5, Izz = 2.173e-5;
    static constexpr double Cd = 7.9379e-06, Ct = 3.25e-4, arm = 65e-3 / 2;
};

// xdot = f(x,u).  Only x[3..12] enter (the dynamics do not depend on position).
CF_DEV void cf_ode_hand(const double *x, const double *u, double *f)
{
    const double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vx = x[7], vy = x[8], vz = x[9];
    const double wx = x[10], wy = x[11], wz = x[12];
    const double s1 = u[0] * u[0], s2 = u[1] * u[1], s3 = u[2] * u[2], s4 = u[3] * u[3];
    f[0] = vx * (2 * q1 * q1 + 2 * q2 * q2 - 1) - vy * (2 * q1 * q4 - 2 * q2 * q3) + vz * (2 * q1 * q3 + 2 * q2 * q4);
    f[1] = vy * (2 * q1 * q1 + 2 * q3 * q3 - 1) + vx * (2 * q1 * q4 + 2 * q2 * q3) - vz * (2 * q1 * q2 - 2 * q3 * q4);
    f[2] = vz * (2 * q1 * q1 + 2 * q4 * q4 - 1) - vx * (2 * q1 * q3 - 2 * q2 * q4) + vy * (2 * q1 * q2 + 2 * q3 * q4);
    f[3] = -(q2 * wx) / 2 - (q3 * wy) / 2 - (q4 * wz) / 2;
    f[4] = (q1 * wx) / 2 - (q4 * wy) / 2 + (q3 * wz) / 2;
    f[5] = (q4 * wx) / 2 + (q1 * wy) / 2 - (q2 * wz) / 2;
    f[6] = (q2 * wy) / 2 - (q3 * wx) / 2 + (q1 * wz) / 2;
    f[7] = vy * wz - vz * wy + CfModel::g0 * (2 * q1 * q3 - 2 * q2 * q4);
    f[8] = vz * wx - vx * wz - CfModel::g0 * (2 * q1 * q2 + 2 * q3 * q4);
    f[9] = vx * wy - vy * wx - CfModel::g0 * (2 * q1 * q1 + 2 * q4 * q4 - 1) + (CfModel::Ct * (s1 + s2 + s3 + s4)) / CfModel::mq;
    f[10] = -(CfModel::Ct * CfModel::arm * (s1 + s2 - s3 - s4) - CfModel::Iyy * wy * wz + CfModel::Izz * wy * wz) / CfModel::Ixx;
    f[11] = -(CfModel::Ct * CfModel::arm * (s1 - s2 - s3 + s4) + CfModel::Ixx * wx * wz - CfModel::Izz * wx * wz) / CfModel::Iyy;
    f[12] = -(CfModel::Cd * (s1 - s2 + s3 - s4) - CfModel::Ixx * wx * wy + CfModel::Iyy * wx * wy) / CfModel::Izz;
}

// Directional derivative o = (df/dx)(x) * d.  The Jacobian is never formed: each lane
// of the warp pushes one sensitivity column through this product (73 structural
// non-zeros).  d[0..2] (position part) does not enter.
CF_DEV void cf_jvp_x_hand(const double *x, const double *d, double *o)
{
    const double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vx = x[7], vy = x[8], vz = x[9];
    const double wx = x[10], wy = x[11], wz = x[12];
    const double e1 = d[3], e2 = d[4], e3 = d[5], e4 = d[6], dvx = d[7], dvy = d[8], dvz = d[9];
    const double dwx = d[10], dwy = d[11], dwz = d[12];
    const double a00 = 2 * q1 * q1 + 2 * q2 * q2 - 1, a01 = 2 * q2 * q3 - 2 * q1 * q4, a02 = 2 * q1 * q3 + 2 * q2 * q4;
    const double a10 = 2 * q1 * q4 + 2 * q2 * q3, a11 = 2 * q1 * q1 + 2 * q3 * q3 - 1, a12 = 2 * q3 * q4 - 2 * q1 * q2;
    const double a20 = 2 * q2 * q4 - 2 * q1 * q3, a21 = 2 * q1 * q2 + 2 * q3 * q4, a22 = 2 * q1 * q1 + 2 * q4 * q4 - 1;
    const double b00 = 4 * (q1 * e1 + q2 * e2);
    const double b01 = 2 * (q2 * e3 + q3 * e2 - q1 * e4 - q4 * e1);
    const double b02 = 2 * (q1 * e3 + q3 * e1 + q2 * e4 + q4 * e2);
    const double b10 = 2 * (q1 * e4 + q4 * e1 + q2 * e3 + q3 * e2);
    const double b11 = 4 * (q1 * e1 + q3 * e3);
    const double b12 = 2 * (q3 * e4 + q4 * e3 - q1 * e2 - q2 * e1);
    const double b20 = 2 * (q2 * e4 + q4 * e2 - q1 * e3 - q3 * e1);
    const double b21 = 2 * (q1 * e2 + q2 * e1 + q3 * e4 + q4 * e3);
    const double b22 = 4 * (q1 * e1 + q4 * e4);
    o[0] = dvx * a00 + dvy * a01 + dvz * a02 + vx * b00 + vy * b01 + vz * b02;
    o[1] = dvx * a10 + dvy * a11 + dvz * a12 + vx * b10 + vy * b11 + vz * b12;
    o[2] = dvx * a20 + dvy * a21 + dvz * a22 + vx * b20 + vy * b21 + vz * b22;
    o[3] = -0.5 * (e2 * wx + q2 * dwx + e3 * wy + q3 * dwy + e4 * wz + q4 * dwz);
    o[4] = 0.5 * (e1 * wx + q1 * dwx - e4 * wy - q4 * dwy + e3 * wz + q3 * dwz);
    o[5] = 0.5 * (e4 * wx + q4 * dwx + e1 * wy + q1 * dwy - e2 * wz - q2 * dwz);
    o[6] = 0.5 * (e2 * wy + q2 * dwy - e3 * wx - q3 * dwx + e1 * wz + q1 * dwz);
    o[7] = dvy * wz + vy * dwz - dvz * wy - vz * dwy - CfModel::g0 * b20;
    o[8] = dvz * wx + vz * dwx - dvx * wz - vx * dwz - CfModel::g0 * b21;
    o[9] = dvx * wy + vx * dwy - dvy * wx - vy * dwx - CfModel::g0 * b22;
    o[10] = -((CfModel::Izz - CfModel::Iyy) / CfModel::Ixx) * (dwy * wz + wy * dwz);
    o[11] = -((CfModel::Ixx - CfModel::Izz) / CfModel::Iyy) * (dwx * wz + wx * dwz);
    o[12] = -((CfModel::Iyy - CfModel::Ixx) / CfModel::Izz) * (dwx * wy + wx * dwy);
}

// o += column j of df/du (only rows 9..12 are non-zero; zero at u = 0)
CF_DEV void cf_add_ju_col_hand(const double *u, int j, double *o)
{
    const double uj = (j == 0) ? u[0] : (j == 1) ? u[1] : (j == 2) ? u[2] : u[3];
    const double s10 = (j < 2) ? 1.0 : -1.0;
    const double s11 = (j == 0 || j == 3) ? 1.0 : -1.0;
    const double s12 = (j == 0 || j == 2) ? 1.0 : -1.0;
    o[9] += 2 * CfModel::Ct * uj / CfModel::mq;
    o[10] += -2 * CfModel::Ct * CfModel::arm * s10 * uj / CfModel::Ixx;
    o[11] += -2 * CfModel::Ct * CfModel::arm * s11 * uj / CfModel::Iyy;
    o[12] += -2 * CfModel::Cd * s12 * uj / CfModel::Izz;
}

#endif  // CF_CRAZYFLIE (hand-derived cross-check)

// ---- what the kernels call: the generated model
CF_DEV void cf_ode(const double *x, const double *u, double *f) { cf_ode_gen(x, u, f); }
CF_DEV void cf_jvp_x(const double *x, const double *u, const double *d, double *o) { cf_jvp_x_gen(x, u, d, o); }
CF_DEV void cf_add_ju_col(const double *x, const double *u, int j, double *o) { cf_ju_col_gen(x, u, j, o); }
