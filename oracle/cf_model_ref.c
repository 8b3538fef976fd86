/* TEST INFRASTRUCTURE ONLY (oracle/): analytic Crazyflie ODE and forward VDE,
 * the external functions the reference's ERK integrator calls
 * (acados/acados/sim/sim_erk_integrator.c:631-642,699-703).
 * Model: crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py:34-42,85-101.
 * Hand-derived; cross-checked against finite differences and the oracle
 * restatement in tests/test_oracle_model.py. */
#include "cf_model_ref.h"

static const double g0 = 9.8066, mq = 33e-3, Ixx = 1.395e-5, Iyy = 1.395e-5, Izz = 2.173e-5,
                    Cd = 7.9379e-06, Ct = 3.25e-4, dq = 65e-3;

void cf_ref_ode(const double *x, const double *u, double *f)
{
    const double l = dq / 2;
    double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vbx = x[7], vby = x[8], vbz = x[9],
           wx = x[10], wy = x[11], wz = x[12];
    double w1 = u[0], w2 = u[1], w3 = u[2], w4 = u[3];
    f[0] = vbx * (2 * q1 * q1 + 2 * q2 * q2 - 1) - vby * (2 * q1 * q4 - 2 * q2 * q3) + vbz * (2 * q1 * q3 + 2 * q2 * q4);
    f[1] = vby * (2 * q1 * q1 + 2 * q3 * q3 - 1) + vbx * (2 * q1 * q4 + 2 * q2 * q3) - vbz * (2 * q1 * q2 - 2 * q3 * q4);
    f[2] = vbz * (2 * q1 * q1 + 2 * q4 * q4 - 1) - vbx * (2 * q1 * q3 - 2 * q2 * q4) + vby * (2 * q1 * q2 + 2 * q3 * q4);
    f[3] = -(q2 * wx) / 2 - (q3 * wy) / 2 - (q4 * wz) / 2;
    f[4] = (q1 * wx) / 2 - (q4 * wy) / 2 + (q3 * wz) / 2;
    f[5] = (q4 * wx) / 2 + (q1 * wy) / 2 - (q2 * wz) / 2;
    f[6] = (q2 * wy) / 2 - (q3 * wx) / 2 + (q1 * wz) / 2;
    f[7] = vby * wz - vbz * wy + g0 * (2 * q1 * q3 - 2 * q2 * q4);
    f[8] = vbz * wx - vbx * wz - g0 * (2 * q1 * q2 + 2 * q3 * q4);
    f[9] = vbx * wy - vby * wx - g0 * (2 * q1 * q1 + 2 * q4 * q4 - 1) + (Ct * (w1 * w1 + w2 * w2 + w3 * w3 + w4 * w4)) / mq;
    f[10] = -(Ct * l * (w1 * w1 + w2 * w2 - w3 * w3 - w4 * w4) - Iyy * wy * wz + Izz * wy * wz) / Ixx;
    f[11] = -(Ct * l * (w1 * w1 - w2 * w2 - w3 * w3 + w4 * w4) + Ixx * wx * wz - Izz * wx * wz) / Iyy;
    f[12] = -(Cd * (w1 * w1 - w2 * w2 + w3 * w3 - w4 * w4) - Ixx * wx * wy + Iyy * wx * wy) / Izz;
}

/* directional derivative Jx(x) * d, d in R^13 */
static void jvp_x(const double *x, const double *d, double *o)
{
    double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vx = x[7], vy = x[8], vz = x[9],
           wx = x[10], wy = x[11], wz = x[12];
    double e1 = d[3], e2 = d[4], e3 = d[5], e4 = d[6], dvx = d[7], dvy = d[8], dvz = d[9],
           dwx = d[10], dwy = d[11], dwz = d[12];
    double a00 = 2 * q1 * q1 + 2 * q2 * q2 - 1, a01 = 2 * q2 * q3 - 2 * q1 * q4, a02 = 2 * q1 * q3 + 2 * q2 * q4;
    double a10 = 2 * q1 * q4 + 2 * q2 * q3, a11 = 2 * q1 * q1 + 2 * q3 * q3 - 1, a12 = 2 * q3 * q4 - 2 * q1 * q2;
    double a20 = 2 * q2 * q4 - 2 * q1 * q3, a21 = 2 * q1 * q2 + 2 * q3 * q4, a22 = 2 * q1 * q1 + 2 * q4 * q4 - 1;
    double b00 = 4 * (q1 * e1 + q2 * e2);
    double b01 = 2 * (q2 * e3 + q3 * e2 - q1 * e4 - q4 * e1);
    double b02 = 2 * (q1 * e3 + q3 * e1 + q2 * e4 + q4 * e2);
    double b10 = 2 * (q1 * e4 + q4 * e1 + q2 * e3 + q3 * e2);
    double b11 = 4 * (q1 * e1 + q3 * e3);
    double b12 = 2 * (q3 * e4 + q4 * e3 - q1 * e2 - q2 * e1);
    double b20 = 2 * (q2 * e4 + q4 * e2 - q1 * e3 - q3 * e1);
    double b21 = 2 * (q1 * e2 + q2 * e1 + q3 * e4 + q4 * e3);
    double b22 = 4 * (q1 * e1 + q4 * e4);
    o[0] = dvx * a00 + dvy * a01 + dvz * a02 + vx * b00 + vy * b01 + vz * b02;
    o[1] = dvx * a10 + dvy * a11 + dvz * a12 + vx * b10 + vy * b11 + vz * b12;
    o[2] = dvx * a20 + dvy * a21 + dvz * a22 + vx * b20 + vy * b21 + vz * b22;
    o[3] = -0.5 * (e2 * wx + q2 * dwx + e3 * wy + q3 * dwy + e4 * wz + q4 * dwz);
    o[4] = 0.5 * (e1 * wx + q1 * dwx - e4 * wy - q4 * dwy + e3 * wz + q3 * dwz);
    o[5] = 0.5 * (e4 * wx + q4 * dwx + e1 * wy + q1 * dwy - e2 * wz - q2 * dwz);
    o[6] = 0.5 * (e2 * wy + q2 * dwy - e3 * wx - q3 * dwx + e1 * wz + q1 * dwz);
    o[7] = dvy * wz + vy * dwz - dvz * wy - vz * dwy - g0 * b20;
    o[8] = dvz * wx + vz * dwx - dvx * wz - vx * dwz - g0 * b21;
    o[9] = dvx * wy + vx * dwy - dvy * wx - vy * dwx - g0 * b22;
    o[10] = -((Izz - Iyy) / Ixx) * (dwy * wz + wy * dwz);
    o[11] = -((Ixx - Izz) / Iyy) * (dwx * wz + wx * dwz);
    o[12] = -((Iyy - Ixx) / Izz) * (dwx * wy + wx * dwy);
}

void cf_ref_vde_forw(const double *x, const double *Sx, const double *Su, const double *u,
                     double *f, double *dSx, double *dSu)
{
    const double l = dq / 2;
    static const double s10[4] = {1, 1, -1, -1}, s11[4] = {1, -1, -1, 1}, s12[4] = {1, -1, 1, -1};
    cf_ref_ode(x, u, f);
    for (int j = 0; j < CF_NX; j++) jvp_x(x, Sx + CF_NX * j, dSx + CF_NX * j);
    for (int j = 0; j < CF_NU; j++) {
        double *o = dSu + CF_NX * j;
        jvp_x(x, Su + CF_NX * j, o);
        o[9] += 2 * Ct * u[j] / mq;
        o[10] += -2 * Ct * l * s10[j] * u[j] / Ixx;
        o[11] += -2 * Ct * l * s11[j] * u[j] / Iyy;
        o[12] += -2 * Cd * s12[j] * u[j] / Izz;
    }
}
