// Kernels either side of the RTI step: what the reference's two ROS nodes do per tick around
// acados_solve(), batched so that thousands of simulated vehicles run closed-loop without leaving the device.
//
//   cf_predict_kernel          state predictor of the estimator node: x+ = ERK4(x, u, T), no sensitivities
//                              (crazyflie_controller/src/acados_estimator.cpp:573-593 -> sim_erk,
//                               acados/acados/sim/sim_erk_integrator.c:658-731)
//   cf_predict_sens_kernel     the same with forward sensitivities S_forw = [Sx | Su] (the generated sim solver keeps
//                              sens_forw on: acados_sim_solver.in.c:280-281,386-396)
//   cf_reference_window_kernel per-tick reference update of NMPC::iteration: regulation / tracking / position hold
//                              (crazyflie_controller/src/acados_mpc.cpp:430-516)
//   cf_policy_advance_kernel   iter++ / switch to position hold at the end of the trajectory (:457-487)
//   cf_command_kernel          what the node publishes from the solution: motor speeds truncated to int32
//                              (msg/PropellerSpeedsStamped.msg, :632-640), attitude set-point from x_4
//                              (quatern2euler :384-404), thrust PWM from mean(u_1) (krpm2pwm :421-425), yaw rate
//                              in deg/s (:641-668)
//
// All of them are HBM-bound element-wise passes: one thread per instance (or per output element), coalesced
// global accesses with the instance rows staged through shared memory where a thread needs a whole row.
#pragma once
#include "cf_model.h"

#define CF_POLICY_REGULATION 0   // enum order of the node's `policy` (acados_mpc.cpp:129-133)
#define CF_POLICY_TRACKING 1
#define CF_POLICY_HOLD 2

#define CF_PRED_THREADS 128

// One explicit RK4 step (tableau sim_collocation_utils.c:611-640), same operation order as sim_erk:
// K_s = f(x + (a_s h) K_{s-1}),  x+ = x + sum_s (h b_s) K_s accumulated in stage order.
__device__ __forceinline__ void cf_rk4_step(double *x, const double *u, double h)
{
    double k[CF_NX], xs[CF_NX], acc[CF_NX];
#pragma unroll
    for (int i = 0; i < CF_NX; i++) { xs[i] = x[i]; acc[i] = x[i]; }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
        const double a_next = (s == 2) ? 1.0 : 0.5;
        cf_ode(xs, u, k);
#pragma unroll
        for (int i = 0; i < CF_NX; i++) {
            acc[i] += (h * bw) * k[i];
            xs[i] = x[i] + (a_next * h) * k[i];
        }
    }
#pragma unroll
    for (int i = 0; i < CF_NX; i++) x[i] = acc[i];
}

// xn[i] = ERK4(x[i], u[i], T_i) with n_steps equal steps.  T_b: per-instance horizons [B] or null (T_all for all).
// motors: if non-null the input is taken from this int32 [B][4] array (the truncated speeds the estimator receives,
// acados_estimator.cpp:463-471) instead of u.  x and xn may alias.
__global__ void __launch_bounds__(CF_PRED_THREADS)
cf_predict_kernel(const double *__restrict__ x, const double *__restrict__ u, const int *__restrict__ motors,
                  const double *__restrict__ T_b, double T_all, int n_steps, int B, double *xn)
{
    __shared__ double sx[CF_PRED_THREADS * CF_NX];
    __shared__ double su[CF_PRED_THREADS * CF_NU];
    const int base = blockIdx.x * CF_PRED_THREADS;
    const int n = min(CF_PRED_THREADS, B - base);
    for (int i = threadIdx.x; i < n * CF_NX; i += CF_PRED_THREADS) sx[i] = x[(long) base * CF_NX + i];
    for (int i = threadIdx.x; i < n * CF_NU; i += CF_PRED_THREADS)
        su[i] = motors ? (double) motors[(long) base * CF_NU + i] : u[(long) base * CF_NU + i];
    __syncthreads();
    if (threadIdx.x < n) {
        double xs[CF_NX], uu[CF_NU];
#pragma unroll
        for (int i = 0; i < CF_NX; i++) xs[i] = sx[threadIdx.x * CF_NX + i];
#pragma unroll
        for (int i = 0; i < CF_NU; i++) uu[i] = su[threadIdx.x * CF_NU + i];
        const double T = T_b ? T_b[base + threadIdx.x] : T_all;
        const double h = T / n_steps;
        for (int st = 0; st < n_steps; st++) cf_rk4_step(xs, uu, h);
#pragma unroll
        for (int i = 0; i < CF_NX; i++) sx[threadIdx.x * CF_NX + i] = xs[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * CF_NX; i += CF_PRED_THREADS) xn[(long) base * CF_NX + i] = sx[i];
}

// One warp per instance: lane c < 13 carries column c of Sx, lanes 13..16 the columns of Su (seed [I 0]); every lane
// integrates the nominal state redundantly.  S_forw [B][13*17] column-major, columns [x | u] as sim_out "S_forw".
__global__ void __launch_bounds__(128)
cf_predict_sens_kernel(const double *__restrict__ x, const double *__restrict__ u, const double *__restrict__ T_b,
                       double T_all, int n_steps, int B, double *__restrict__ xn, double *__restrict__ S_forw)
{
    const int inst = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (inst >= B) return;
    double X[CF_NX], uu[CF_NU], S[CF_NX];
#pragma unroll
    for (int i = 0; i < CF_NX; i++) { X[i] = x[(long) inst * CF_NX + i]; S[i] = (lane == i) ? 1.0 : 0.0; }
#pragma unroll
    for (int i = 0; i < CF_NU; i++) uu[i] = u[(long) inst * CF_NU + i];
    const double T = T_b ? T_b[inst] : T_all;
    const double h = T / n_steps;
    for (int st = 0; st < n_steps; st++) {
        double xs[CF_NX], Ss[CF_NX], xa[CF_NX], Sa[CF_NX];
#pragma unroll
        for (int i = 0; i < CF_NX; i++) { xs[i] = X[i]; Ss[i] = S[i]; xa[i] = X[i]; Sa[i] = S[i]; }
#pragma unroll 1
        for (int s = 0; s < 4; s++) {
            const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
            const double a_next = (s == 2) ? 1.0 : 0.5;
            double f[CF_NX], ks[CF_NX];
            cf_ode(xs, uu, f);
            cf_jvp_x(xs, uu, Ss, ks);
            if (lane >= CF_NX && lane < CF_NV) cf_add_ju_col(xs, uu, lane - CF_NX, ks);
#pragma unroll
            for (int i = 0; i < CF_NX; i++) {
                xa[i] += (h * bw) * f[i];
                Sa[i] += (h * bw) * ks[i];
                xs[i] = X[i] + (a_next * h) * f[i];
                Ss[i] = S[i] + (a_next * h) * ks[i];
            }
        }
#pragma unroll
        for (int i = 0; i < CF_NX; i++) { X[i] = xa[i]; S[i] = Sa[i]; }
    }
    if (lane < CF_NV) {
#pragma unroll
        for (int i = 0; i < CF_NX; i++) S_forw[(long) inst * (CF_NX * CF_NV) + lane * CF_NX + i] = S[i];
    }
    if (lane == 17) {
#pragma unroll
        for (int i = 0; i < CF_NX; i++) xn[(long) inst * CF_NX + i] = X[i];
    }
}

// yref[i][k][0:17] for k < N and yref_e[i][0:13] (row k = N) from the instance's policy; one thread per element.
// Tracking past the end of the table leaves the previous window in place for this tick, exactly as the node does
// (the `else policy = Position_Hold; break;` path writes nothing, acados_mpc.cpp:486-487).
__global__ void cf_reference_window_kernel(int B, int N, const int *__restrict__ policy, const int *__restrict__ titer,
                                           const double *__restrict__ setpoint, const double *__restrict__ traj, int n_rows,
                                           double uss, double *__restrict__ yref, double *__restrict__ yref_e)
{
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (long) (N + 1) * CF_NY;
    if (t >= (long) B * per) return;
    const int i = (int) (t / per);
    const int r = (int) (t - (long) i * per);
    const int k = r / CF_NY, e = r - k * CF_NY;
    if (k == N && e >= CF_NX) return;
    const int pol = policy[i];
    double v;
    if (pol == CF_POLICY_TRACKING) {
        const int it = titer[i];
        if (!(traj && it < n_rows - N)) return;
        v = traj[(long) (it + k) * CF_NY + e];
    } else {
        if (e < 3) v = (pol == CF_POLICY_HOLD && traj) ? traj[(long) (n_rows - 1) * CF_NY + e] : setpoint[i * 3 + e];
        else if (e == 3) v = 1.0;
        else if (e < CF_NX) v = 0.0;
        else v = uss;
    }
    if (k < N) yref[((long) i * N + k) * CF_NY + e] = v;
    else yref_e[(long) i * CF_NX + e] = v;
}

__global__ void cf_policy_advance_kernel(int B, int N, int *policy, int *titer, int n_rows)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    if (policy[i] == CF_POLICY_TRACKING) {
        if (titer[i] < n_rows - N) titer[i] += 1;
        else policy[i] = CF_POLICY_HOLD;
    }
}

// motors int32 [B][4]; euler double [B][3] = (phi, theta, psi) of the normalised x_4 quaternion;
// twist double [B][4] = (linear.x = pitch [deg], linear.y = -roll [deg], linear.z = thrust PWM, angular.z = yaw rate [deg/s])
__global__ void cf_command_kernel(int B, int N, const double *__restrict__ x, const double *__restrict__ u, int motors_from_u1,
                                  int *__restrict__ motors, double *__restrict__ euler, double *__restrict__ twist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double *ui = u + (long) i * N * CF_NU;
    const int s1 = N > 1 ? 1 : 0, s4 = N >= 4 ? 4 : N;
    const double *u1 = ui + s1 * CF_NU;
    const double *x4 = x + ((long) i * (N + 1) + s4) * CF_NX;
    const double *um = motors_from_u1 ? u1 : ui;
#pragma unroll
    for (int j = 0; j < CF_NU; j++) motors[i * CF_NU + j] = (int) um[j];   // float64 -> int32 field: truncation
    double qw = x4[3], qx = x4[4], qy = x4[5], qz = x4[6];
    const double nrm = sqrt(qx * qx + qy * qy + qz * qz + qw * qw);        // Eigen::Quaterniond::normalize()
    qw /= nrm; qx /= nrm; qy /= nrm; qz /= nrm;
    const double R11 = 2 * (qw * qw + qx * qx) - 1, R21 = 2 * (qx * qy - qw * qz), R31 = 2 * (qx * qz + qw * qy);
    const double R32 = 2 * (qy * qz - qw * qx), R33 = 2 * (qw * qw + qz * qz) - 1;
    const double phi = atan2(R32, R33), theta = -asin(R31), psi = atan2(R21, R11);
    euler[i * 3 + 0] = phi; euler[i * 3 + 1] = theta; euler[i * 3 + 2] = psi;
    const double pi_node = 3.14159265358979323846;
    const double mean_u1 = (u1[0] + u1[1] + u1[2] + u1[3]) / 4;
    const int pwm = (int) (((mean_u1 * 1000) - 4070.3) / 0.2685);          // krpm2pwm
    twist[i * 4 + 0] = 1.0 * (theta * 180.0 / pi_node);
    twist[i * 4 + 1] = -1.0 * (phi * 180.0 / pi_node);
    twist[i * 4 + 2] = (double) pwm;
    twist[i * 4 + 3] = x4[12] * 180.0 / pi_node;
}
