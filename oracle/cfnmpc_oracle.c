/* TEST INFRASTRUCTURE ONLY -- see cfnmpc_oracle.h for scope, parity status and
 * the reference file:line map.  Plain scalar C, one instance at a time, dense
 * loops, no attempt at speed.  Never part of the product path. */
#include "cfnmpc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NX 13
#define NU 4
#define NV 17

/* ------------------------------------------------------------------ model
 * crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py:34-42,85-97 */
static const double g0 = 9.8066, mq = 33e-3, Ixx = 1.395e-5, Iyy = 1.395e-5, Izz = 2.173e-5,
                    Cd = 7.9379e-06, Ct = 3.25e-4, arm = 65e-3 / 2;

void cfo_default_params(cfo_params *p)
{
    /* generate_c_code.py:61-84,113,133-134 */
    static const double Q[NX] = {120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0};
    for (int i = 0; i < NX; i++) { p->Wdiag[i] = Q[i]; p->WNdiag[i] = 50 * Q[i]; }
    for (int i = 0; i < NU; i++) { p->Wdiag[NX + i] = 0.06; p->lbu[i] = p->lbu0[i] = 0.0; p->ubu[i] = p->ubu0[i] = 22.0; }
    p->has_u0 = 0;
}

void cfo_ode(const double *x, const double *u, double *f)
{
    double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vx = x[7], vy = x[8], vz = x[9];
    double wx = x[10], wy = x[11], wz = x[12];
    double s1 = u[0] * u[0], s2 = u[1] * u[1], s3 = u[2] * u[2], s4 = u[3] * u[3];
    f[0] = vx * (2 * q1 * q1 + 2 * q2 * q2 - 1) - vy * (2 * q1 * q4 - 2 * q2 * q3) + vz * (2 * q1 * q3 + 2 * q2 * q4);
    f[1] = vy * (2 * q1 * q1 + 2 * q3 * q3 - 1) + vx * (2 * q1 * q4 + 2 * q2 * q3) - vz * (2 * q1 * q2 - 2 * q3 * q4);
    f[2] = vz * (2 * q1 * q1 + 2 * q4 * q4 - 1) - vx * (2 * q1 * q3 - 2 * q2 * q4) + vy * (2 * q1 * q2 + 2 * q3 * q4);
    f[3] = -(q2 * wx) / 2 - (q3 * wy) / 2 - (q4 * wz) / 2;
    f[4] = (q1 * wx) / 2 - (q4 * wy) / 2 + (q3 * wz) / 2;
    f[5] = (q4 * wx) / 2 + (q1 * wy) / 2 - (q2 * wz) / 2;
    f[6] = (q2 * wy) / 2 - (q3 * wx) / 2 + (q1 * wz) / 2;
    f[7] = vy * wz - vz * wy + g0 * (2 * q1 * q3 - 2 * q2 * q4);
    f[8] = vz * wx - vx * wz - g0 * (2 * q1 * q2 + 2 * q3 * q4);
    f[9] = vx * wy - vy * wx - g0 * (2 * q1 * q1 + 2 * q4 * q4 - 1) + (Ct * (s1 + s2 + s3 + s4)) / mq;
    f[10] = -(Ct * arm * (s1 + s2 - s3 - s4) - Iyy * wy * wz + Izz * wy * wz) / Ixx;
    f[11] = -(Ct * arm * (s1 - s2 - s3 + s4) + Ixx * wx * wz - Izz * wx * wz) / Iyy;
    f[12] = -(Cd * (s1 - s2 + s3 - s4) - Ixx * wx * wy + Iyy * wx * wy) / Izz;
}

/* dense Jacobians, row-major Jx[13][13], Ju[13][4]; sparsity as SURVEY.md Appendix A.1 */
static void jacobians(const double *x, const double *u, double Jx[NX][NX], double Ju[NX][NU])
{
    double q1 = x[3], q2 = x[4], q3 = x[5], q4 = x[6], vx = x[7], vy = x[8], vz = x[9];
    double wx = x[10], wy = x[11], wz = x[12];
    memset(Jx, 0, sizeof(double) * NX * NX);
    memset(Ju, 0, sizeof(double) * NX * NU);
    /* pdot rows */
    Jx[0][3] = 4 * q1 * vx - 2 * q4 * vy + 2 * q3 * vz;
    Jx[0][4] = 4 * q2 * vx + 2 * q3 * vy + 2 * q4 * vz;
    Jx[0][5] = 2 * q2 * vy + 2 * q1 * vz;
    Jx[0][6] = -2 * q1 * vy + 2 * q2 * vz;
    Jx[0][7] = 2 * q1 * q1 + 2 * q2 * q2 - 1;
    Jx[0][8] = -(2 * q1 * q4 - 2 * q2 * q3);
    Jx[0][9] = 2 * q1 * q3 + 2 * q2 * q4;
    Jx[1][3] = 4 * q1 * vy + 2 * q4 * vx - 2 * q2 * vz;
    Jx[1][4] = 2 * q3 * vx - 2 * q1 * vz;
    Jx[1][5] = 4 * q3 * vy + 2 * q2 * vx + 2 * q4 * vz;
    Jx[1][6] = 2 * q1 * vx + 2 * q3 * vz;
    Jx[1][7] = 2 * q1 * q4 + 2 * q2 * q3;
    Jx[1][8] = 2 * q1 * q1 + 2 * q3 * q3 - 1;
    Jx[1][9] = -(2 * q1 * q2 - 2 * q3 * q4);
    Jx[2][3] = 4 * q1 * vz - 2 * q3 * vx + 2 * q2 * vy;
    Jx[2][4] = 2 * q4 * vx + 2 * q1 * vy;
    Jx[2][5] = -2 * q1 * vx + 2 * q4 * vy;
    Jx[2][6] = 4 * q4 * vz + 2 * q2 * vx + 2 * q3 * vy;
    Jx[2][7] = -(2 * q1 * q3 - 2 * q2 * q4);
    Jx[2][8] = 2 * q1 * q2 + 2 * q3 * q4;
    Jx[2][9] = 2 * q1 * q1 + 2 * q4 * q4 - 1;
    /* qdot rows */
    Jx[3][4] = -wx / 2; Jx[3][5] = -wy / 2; Jx[3][6] = -wz / 2;
    Jx[3][10] = -q2 / 2; Jx[3][11] = -q3 / 2; Jx[3][12] = -q4 / 2;
    Jx[4][3] = wx / 2; Jx[4][5] = wz / 2; Jx[4][6] = -wy / 2;
    Jx[4][10] = q1 / 2; Jx[4][11] = -q4 / 2; Jx[4][12] = q3 / 2;
    Jx[5][3] = wy / 2; Jx[5][4] = -wz / 2; Jx[5][6] = wx / 2;
    Jx[5][10] = q4 / 2; Jx[5][11] = q1 / 2; Jx[5][12] = -q2 / 2;
    Jx[6][3] = wz / 2; Jx[6][4] = wy / 2; Jx[6][5] = -wx / 2;
    Jx[6][10] = -q3 / 2; Jx[6][11] = q2 / 2; Jx[6][12] = q1 / 2;
    /* vdot rows */
    Jx[7][3] = 2 * g0 * q3; Jx[7][4] = -2 * g0 * q4; Jx[7][5] = 2 * g0 * q1; Jx[7][6] = -2 * g0 * q2;
    Jx[7][8] = wz; Jx[7][9] = -wy; Jx[7][11] = -vz; Jx[7][12] = vy;
    Jx[8][3] = -2 * g0 * q2; Jx[8][4] = -2 * g0 * q1; Jx[8][5] = -2 * g0 * q4; Jx[8][6] = -2 * g0 * q3;
    Jx[8][7] = -wz; Jx[8][9] = wx; Jx[8][10] = vz; Jx[8][12] = -vx;
    Jx[9][3] = -4 * g0 * q1; Jx[9][6] = -4 * g0 * q4;
    Jx[9][7] = wy; Jx[9][8] = -wx; Jx[9][10] = -vy; Jx[9][11] = vx;
    /* wdot rows */
    Jx[10][11] = -(Izz - Iyy) * wz / Ixx; Jx[10][12] = -(Izz - Iyy) * wy / Ixx;
    Jx[11][10] = -(Ixx - Izz) * wz / Iyy; Jx[11][12] = -(Ixx - Izz) * wx / Iyy;
    Jx[12][10] = -(Iyy - Ixx) * wy / Izz; Jx[12][11] = -(Iyy - Ixx) * wx / Izz;
    static const double s10[4] = {1, 1, -1, -1}, s11[4] = {1, -1, -1, 1}, s12[4] = {1, -1, 1, -1};
    for (int j = 0; j < NU; j++) {
        Ju[9][j] = 2 * Ct * u[j] / mq;
        Ju[10][j] = -2 * Ct * arm * s10[j] * u[j] / Ixx;
        Ju[11][j] = -2 * Ct * arm * s11[j] * u[j] / Iyy;
        Ju[12][j] = -2 * Cd * s12[j] * u[j] / Izz;
    }
}

/* forward VDE, column-major Sx[13x13], Su[13x4] like the reference's external
 * function (acados_template/casadi_function_generation.py:137-151) */
void cfo_vde(const double *x, const double *Sx, const double *Su, const double *u, double *f, double *dSx, double *dSu)
{
    double Jx[NX][NX], Ju[NX][NU];
    jacobians(x, u, Jx, Ju);
    cfo_ode(x, u, f);
    for (int j = 0; j < NX; j++)
        for (int i = 0; i < NX; i++) {
            double s = 0;
            for (int k = 0; k < NX; k++) s += Jx[i][k] * Sx[k + NX * j];
            dSx[i + NX * j] = s;
        }
    for (int j = 0; j < NU; j++)
        for (int i = 0; i < NX; i++) {
            double s = Ju[i][j];
            for (int k = 0; k < NX; k++) s += Jx[i][k] * Su[k + NX * j];
            dSu[i + NX * j] = s;
        }
}

/* sim_erk_integrator.c:658-731, tableau sim_collocation_utils.c:611-640 */
#define NXX (NX + NX * NX + NX * NU)
void cfo_erk4(const double *x, const double *u, double h, double *xn, double *A, double *B)
{
    static const double a_prev[4] = {0, 0.5, 0.5, 1.0}; /* A[s][s-1], all other entries 0 */
    static const double bw[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    double X[NXX], K[4][NXX], R[NXX];
    memset(X, 0, sizeof X);
    for (int i = 0; i < NX; i++) { X[i] = x[i]; X[NX + i * (NX + 1)] = 1.0; } /* Sx = I, Su = 0 */
    for (int s = 0; s < 4; s++) {
        for (int i = 0; i < NXX; i++) R[i] = X[i];
        if (s > 0) {
            double a = a_prev[s] * h;
            for (int i = 0; i < NXX; i++) R[i] += a * K[s - 1][i];
        }
        cfo_vde(R, R + NX, R + NX + NX * NX, u, K[s], K[s] + NX, K[s] + NX + NX * NX);
    }
    for (int s = 0; s < 4; s++) {
        double b = h * bw[s];
        for (int i = 0; i < NXX; i++) X[i] += b * K[s][i];
    }
    for (int i = 0; i < NX; i++) xn[i] = X[i];
    for (int i = 0; i < NX; i++) {
        for (int j = 0; j < NX; j++) A[i * NX + j] = X[NX + i + NX * j];
        for (int j = 0; j < NU; j++) B[i * NU + j] = X[NX + NX * NX + i + NX * j];
    }
}

void cfo_sim(const double *x, const double *u, double T, int n_steps, double *xn)
{
    static const double a_prev[4] = {0, 0.5, 0.5, 1.0};
    static const double bw[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    double h = T / n_steps, X[NX], K[4][NX], R[NX];
    for (int i = 0; i < NX; i++) X[i] = x[i];
    for (int st = 0; st < n_steps; st++) {
        for (int s = 0; s < 4; s++) {
            for (int i = 0; i < NX; i++) R[i] = X[i];
            if (s > 0) { double a = a_prev[s] * h; for (int i = 0; i < NX; i++) R[i] += a * K[s - 1][i]; }
            cfo_ode(R, u, K[s]);
        }
        for (int s = 0; s < 4; s++) { double b = h * bw[s]; for (int i = 0; i < NX; i++) X[i] += b * K[s][i]; }
    }
    for (int i = 0; i < NX; i++) xn[i] = X[i];
}

/* Non-uniform shooting grid: interval lengths = cost scalings, as crazyflie_acados_update_time_steps sets them
 * (acados_template/c_templates_tera/acados_solver.in.c:133-153).  Global, test use only; n = 0 returns to the
 * uniform grid given by the Ts argument of the calls below. */
#define CFO_MAX_N 4096
static double g_dt[CFO_MAX_N];
static int g_dt_n = 0;
void cfo_set_time_steps(const double *dt, int n)
{
    if (!dt || n <= 0 || n > CFO_MAX_N) { g_dt_n = 0; return; }
    memcpy(g_dt, dt, sizeof(double) * n);
    g_dt_n = n;
}
#define DTK(k) ((k) < g_dt_n ? g_dt[k] : Ts)

/* Input box per stage, tab[n][8] = lbu(4) | ubu(4): ocp_nlp_constraints_model_set addresses one stage at a time
 * (ocp_nlp_constraints_bgh.c:653-674).  Global, test use only; n = 0 returns to the boxes of cfo_params. */
/* cost weights per stage, [N+1][17] (row N: W_e): ocp_nlp_cost_model_set(.., k, "W", ..) addresses one stage at a time
 * (ocp_nlp_cost_ls.c:301-331).  Global in the checker, like the time grid. */
static double g_wst[(CFO_MAX_N + 1) * 17];
static int g_wst_n = 0;
void cfo_set_stage_weights(const double *tab, int n_rows)
{
    g_wst_n = 0;
    if (tab && n_rows > 0 && n_rows <= CFO_MAX_N + 1) { memcpy(g_wst, tab, sizeof(double) * 17 * n_rows); g_wst_n = n_rows; }
}
#define WKS(p, k, idx) ((k) < g_wst_n ? g_wst[17 * (k) + (idx)] : (p)->Wdiag[idx])   /* stages k < N */
#define WK(p, k, N, idx) ((k) < g_wst_n ? g_wst[17 * (k) + (idx)] : ((k) < (N) ? (p)->Wdiag[idx] : (p)->WNdiag[idx]))
/* Full (non-diagonal) weight matrices, one per stage: tab[n_rows][17*17] row-major, symmetric, cost order y = [x;u]; row N
 * is W_e (its leading 13 x 13 block).  ocp_nlp_cost_ls.c:301-331 (any SPD W; the Cholesky factor enters the Hessian,
 * W itself the gradient).  Global, test use only; n_rows = 0 returns to the diagonal weights. */
#define CFO_MAX_DENSE 257
static double g_wd[CFO_MAX_DENSE * 289], g_hd[CFO_MAX_DENSE * 289];   /* W and chol(W) chol(W)' per stage */
static int g_wd_n = 0;
void cfo_set_dense_weights(const double *tab, int n_rows)
{
    g_wd_n = 0;
    if (!tab || n_rows <= 0 || n_rows > CFO_MAX_DENSE) return;
    memcpy(g_wd, tab, sizeof(double) * 289 * n_rows);
    for (int k = 0; k < n_rows; k++) {
        /* hess = (Cyt W_chol)(Cyt W_chol)' (ocp_nlp_cost_ls.c:743-772, blasfeo_dpotrf_l + dtrmm + dsyrk) */
        const int n = (k == n_rows - 1) ? NX : NV;
        const double *W = g_wd + 289 * k;
        double L[17][17];
        memset(L, 0, sizeof L);
        for (int j = 0; j < n; j++) {
            double d = W[j * 17 + j];
            for (int c = 0; c < j; c++) d -= L[j][c] * L[j][c];
            d = d > 0 ? sqrt(d) : 0.0;
            L[j][j] = d;
            for (int i = j + 1; i < n; i++) {
                double v = W[i * 17 + j];
                for (int c = 0; c < j; c++) v -= L[i][c] * L[j][c];
                L[i][j] = d > 0 ? v / d : 0.0;
            }
        }
        double *H = g_hd + 289 * k;
        for (int i = 0; i < 17; i++) for (int j = 0; j < 17; j++) {
            double v = 0;
            if (i < n && j < n) for (int c = 0; c <= (i < j ? i : j); c++) v += L[i][c] * L[j][c];
            H[i * 17 + j] = v;
        }
    }
    g_wd_n = n_rows;
}
/* cost index (y = [x;u]) of stage variable r ([u;x] order) */
#define YIDX(r) ((r) < NU ? NX + (r) : (r) - NU)
static double g_bst[CFO_MAX_N * 8];
static int g_bst_n = 0;
void cfo_set_stage_bounds(const double *tab, int n)
{
    if (!tab || n <= 0 || n > CFO_MAX_N) { g_bst_n = 0; return; }
    memcpy(g_bst, tab, sizeof(double) * 8 * n);
    g_bst_n = n;
}

/* ------------------------------------------------------------ linearisation
 * ocp_nlp_common.c:2157-2292 calling dynamics_cont :755-884, cost_ls :810-916,
 * constraints_bgh :1613-1648.  Output layout = cfref_get_qp. */
void cfo_linearize(int N, double Ts, const cfo_params *p_, const double *x0, const double *yref,
                   const double *yref_e, const double *x, const double *u, double *BAbt, double *b,
                   double *rqz, double *d_lb, double *d_ub)
{
    cfo_params pd;
    if (!p_) { cfo_default_params(&pd); p_ = &pd; }
    int od = 0;
    for (int k = 0; k <= N; k++) {
        const double *xk = x + NX * k;
        if (k < N) {
            const double *uk = u + NU * k;
            double xn[NX], A[NX * NX], B[NX * NU];
            cfo_erk4(xk, uk, DTK(k), xn, A, B);
            if (BAbt) {
                double *M = BAbt + (size_t) k * NV * NX;
                for (int j = 0; j < NU; j++) for (int i = 0; i < NX; i++) M[j * NX + i] = B[i * NU + j];
                for (int j = 0; j < NX; j++) for (int i = 0; i < NX; i++) M[(NU + j) * NX + i] = A[i * NX + j];
            }
            if (b) for (int i = 0; i < NX; i++) b[NX * k + i] = xn[i] - x[NX * (k + 1) + i];
            if (rqz) {
                /* grad = scaling * Cyt * W * (Cy ux - yref), [u;x] order */
                double *g = rqz + NV * k;
                const double *yr = yref + NV * k;
                for (int i = 0; i < NU; i++) g[i] = (WK(p_, k, N, NX + i) * (uk[i] - yr[NX + i])) * DTK(k);
                for (int i = 0; i < NX; i++) g[NU + i] = (WK(p_, k, N, i) * (xk[i] - yr[i])) * DTK(k);
                if (k < g_wd_n) {   /* tmp = W res (symv), grad = scaling * Cyt tmp  (ocp_nlp_cost_ls.c:883-912) */
                    double res[NV];
                    for (int j = 0; j < NX; j++) res[j] = xk[j] - yr[j];
                    for (int j = 0; j < NU; j++) res[NX + j] = uk[j] - yr[NX + j];
                    for (int r = 0; r < NV; r++) {
                        double v = 0;
                        for (int j = 0; j < NV; j++) v += g_wd[289 * k + YIDX(r) * 17 + j] * res[j];
                        g[r] = v * DTK(k);
                    }
                }
            }
            int nb = k == 0 ? NV : NU;
            for (int i = 0; i < NU; i++) {
                /* per-stage bounds: ocp_nlp_constraints_bgh.c:653-674 (model_set copies into stage k only) */
                const int s0 = (k == 0 && p_->has_u0);
                const double lb = k < g_bst_n ? g_bst[8 * k + i] : (s0 ? p_->lbu0[i] : p_->lbu[i]);
                const double ub = k < g_bst_n ? g_bst[8 * k + 4 + i] : (s0 ? p_->ubu0[i] : p_->ubu[i]);
                if (d_lb) d_lb[od + i] = lb - uk[i];
                if (d_ub) d_ub[od + i] = uk[i] - ub;
            }
            if (k == 0)
                for (int i = 0; i < NX; i++) {
                    if (d_lb) d_lb[od + NU + i] = x0[i] - xk[i];
                    if (d_ub) d_ub[od + NU + i] = xk[i] - x0[i];
                }
            od += nb;
        } else if (rqz) {
            double *g = rqz + NV * N;
            for (int i = 0; i < NX; i++) g[i] = WK(p_, N, N, i) * (xk[i] - yref_e[i]);
            if (N < g_wd_n)
                for (int i = 0; i < NX; i++) {
                    double v = 0;
                    for (int j = 0; j < NX; j++) v += g_wd[289 * N + i * 17 + j] * (xk[j] - yref_e[j]);
                    g[i] = v;
                }
        }
    }
}

/* ------------------------------------------------------------------ QP / IPM
 * Reduced QP after x0 elimination (x_ocp_qp_red.c:268-455): stage 0 has nx=0. */
#define NVM 32 /* largest stage the generic routines handle: partially condensed stages have nu = 4 * block size */
#define NUM 19
typedef struct
{
    int nu, nx, nv, nb;
    int dense;            /* != 0: Hessian Hm (partially condensed stage) instead of the diagonal H */
    double M[NVM][NX]; /* [B';A'] rows (nv of them) */
    double b[NX];
    double H[NVM], rq[NVM]; /* diagonal Hessian (+ gradient) */
    double Hm[NVM][NVM];    /* dense symmetric Hessian of a condensed stage (both triangles) */
    double d[2 * NUM];     /* [lb; ub] (HPIPM sign convention) */
    double ux[NVM], pi[NX], lam[2 * NUM], t[2 * NUM];
    double res_g[NVM], res_b[NX], res_d[2 * NUM], res_m[2 * NUM], res_m_bkp[2 * NUM];
    double dux[NVM], dpi[NX], dlam[2 * NUM], dt[2 * NUM];
    double rg2[NVM], rb2[NX], rd2[2 * NUM], rm2[2 * NUM];                /* res_itref */
    double dux2[NVM], dpi2[NX], dlam2[2 * NUM], dt2[2 * NUM];            /* sol_itref */
    double L[NVM + 1][NVM];
    double Pm[NX][NX], pv[NX]; /* classical Riccati: cost-to-go Hessian / gradient of this stage */
    double Gamma[2 * NUM], gamma[2 * NUM], t_inv[2 * NUM], Pb[NX];
} stage;

typedef struct
{
    int N, nc;
    stage *s;
    double mu, alpha, mu_aff, sigma;
    double res_max[4], res2_max[4];
} ipm;

/* HPIPM arguments in effect: BALANCE mode + acados overrides, ocp_qp_hpipm.c:96-108,
 * x_ocp_qp_ipm.c:133-161 */
static const double RES_G_MAX = 1e-6, RES_B_MAX = 1e-8, RES_D_MAX = 1e-8, RES_M_MAX = 1e-8;
static const double ALPHA_MIN = 1e-8, MU0 = 1.0, REG_PRIM = 1e-15, LAM_MIN = 1e-16, T_MIN = 1e-16, TAU_MIN = 1e-16;
static const int ITREF_CORR_MAX = 2;
static int ITER_MAX = 50;
/* qp_iter_max (ocp_qp_hpipm.c:96-108 sets 50); tests lower it to exercise the MAXITER branch of ocp_nlp_sqp_rti.c:651-674 */
void cfo_set_iter_max(int n) { ITER_MAX = (n > 0 && n < 50) ? n : 50; }

/* 0: square-root Riccati (the reference configuration, x_ocp_qp_kkt.c:445-572);
 * 1: HPIPM's classical Riccati (square_root_alg = 0, x_ocp_qp_kkt.c:573-740) -- the variant the CUDA
 * factorisation uses; kept here so that tests can quantify the difference between the two. */
static int g_classical = 0;
void cfo_set_classical_riccati(int on) { g_classical = on; }

/* lower Cholesky of the n x n top of an m x n block, remaining rows solved;
 * non-positive pivot -> 0 (BLASFEO kernel_dgemm_4x4_lib4.c:5701-5714) */
static void potrf_l_mn(int m, int n, double L[NVM + 1][NVM])
{
    for (int j = 0; j < n; j++) {
        double djj = L[j][j];
        for (int c = 0; c < j; c++) djj -= L[j][c] * L[j][c];
        double inv;
        if (djj > 0) { djj = sqrt(djj); inv = 1.0 / djj; } else { djj = 0.0; inv = 0.0; }
        L[j][j] = djj;
        for (int i = j + 1; i < m; i++) {
            double v = L[i][j];
            for (int c = 0; c < j; c++) v -= L[i][c] * L[j][c];
            L[i][j] = v * inv;
        }
    }
}

/* Gamma, gamma: x_core_qp_ipm_aux.c:38-111 (t_lam_min==2 -> plain branch) */
static void compute_Gamma_gamma(ipm *w, int with_Gamma, int itref)
{
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        const double *rd = itref ? s->rd2 : s->res_d, *rm = itref ? s->rm2 : s->res_m;
        for (int i = 0; i < 2 * s->nb; i++) {
            if (with_Gamma) { s->t_inv[i] = 1.0 / s->t[i]; s->Gamma[i] = s->t_inv[i] * s->lam[i]; }
            s->gamma[i] = s->t_inv[i] * (rm[i] - s->lam[i] * rd[i]);
        }
    }
}

/* dlam, dt from dux: x_ocp_qp_kkt.c:741-758, x_core_qp_ipm_aux.c:117-142 */
static void compute_lam_t(ipm *w, int itref)
{
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        const double *rd = itref ? s->rd2 : s->res_d, *rm = itref ? s->rm2 : s->res_m;
        const double *dux = itref ? s->dux2 : s->dux;
        double *dlam = itref ? s->dlam2 : s->dlam, *dt = itref ? s->dt2 : s->dt;
        for (int i = 0; i < s->nb; i++) { dt[i] = dux[i]; dt[s->nb + i] = -dux[i]; }
        for (int i = 0; i < 2 * s->nb; i++) {
            dlam[i] = -s->t_inv[i] * (rm[i] + (s->lam[i] * dt[i]) - (s->lam[i] * rd[i]));
            dt[i] -= rd[i];
        }
    }
}

/* forward substitution shared by factorise-and-solve and solve
 * (x_ocp_qp_kkt.c:536-570 / :1250-1290).  On entry dux[k][0:nu] holds the
 * u-part rhs (already negated), pl[k] the x-part "l" vector of stage k. */
static void ric_forward(ipm *w, double (*pl)[NX], int itref, int l_is_scaled)
{
    int N = w->N;
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        double *dux = itref ? s->dux2 : s->dux;
        const double *rb = itref ? s->rb2 : s->res_b;
        /* TRSV_LTN_MN(nv, nu): du = Luu^-T (du - Lxu' dx) */
        for (int j = s->nu - 1; j >= 0; j--) {
            double v = dux[j];
            for (int i = j + 1; i < s->nv; i++) v -= s->L[i][j] * dux[i];
            dux[j] = v * (1.0 / s->L[j][j]);
        }
        if (k == N) break;
        stage *n = &w->s[k + 1];
        double *ndux = itref ? n->dux2 : n->dux, *dpi = itref ? s->dpi2 : s->dpi;
        /* dx+ = [B A] dux + res_b */
        for (int c = 0; c < NX; c++) {
            double v = 0;
            for (int r = 0; r < s->nv; r++) v += s->M[r][c] * dux[r];
            ndux[n->nu + c] = v + rb[c];
        }
        /* dpi: x_ocp_qp_kkt.c:545-547 (factorise) / :1262-1266 (solve) */
        if (g_classical) { /* dpi = P dx+ + p   (GEMV_N, x_ocp_qp_kkt.c:712,729) */
            for (int r = 0; r < NX; r++) {
                double v = pl[k + 1][r];
                for (int c = 0; c < NX; c++) v += n->Pm[r][c] * ndux[n->nu + c];
                dpi[r] = v;
            }
            continue;
        }
        double tmp[NX];
        for (int c = 0; c < NX; c++) {
            double v = 0;
            for (int r = c; r < NX; r++) v += n->L[n->nu + r][n->nu + c] * ndux[n->nu + r];
            tmp[c] = l_is_scaled ? v + pl[k + 1][c] : v; /* fact: Lxx (Lxx' dx + l~) */
        }
        for (int r = 0; r < NX; r++) {
            double v = 0;
            for (int c = 0; c <= r; c++) v += n->L[n->nu + r][n->nu + c] * tmp[c];
            dpi[r] = l_is_scaled ? v : v + pl[k + 1][r]; /* solve: p + Lxx Lxx' dx */
        }
    }
}

/* OCP_QP_FACT_SOLVE_KKT_STEP, square-root branch: x_ocp_qp_kkt.c:445-572 */
static void fact_solve_kkt_step(ipm *w)
{
    int N = w->N;
    double(*pl)[NX] = malloc(sizeof(double[NX]) * (N + 1));
    compute_Gamma_gamma(w, 1, 0);
    for (int k = N; k >= 0 && g_classical; k--) {
        stage *s = &w->s[k];
        double AL[NVM + 1][NX];
        if (k < N) {
            stage *n = &w->s[k + 1];
            for (int r = 0; r <= s->nv; r++) /* AL = [B';A';res_b'] * P(k+1)   (GEMM_NT, :621) */
                for (int c = 0; c < NX; c++) {
                    double v = 0;
                    for (int j = 0; j < NX; j++) v += (r < s->nv ? s->M[r][j] : s->res_b[j]) * n->Pm[j][c];
                    AL[r][c] = v;
                }
            for (int c = 0; c < NX; c++) { s->Pb[c] = AL[s->nv][c]; AL[s->nv][c] += n->pv[c]; }
        }
        memset(s->L, 0, sizeof s->L);
        for (int i = 0; i < s->nv; i++) { s->L[i][i] = s->H[i] + REG_PRIM; s->L[s->nv][i] = s->res_g[i]; }
        if (s->dense)
            for (int i = 0; i < s->nv; i++)
                for (int j = 0; j <= i; j++) s->L[i][j] = s->Hm[i][j] + (i == j ? REG_PRIM : 0.0);
        for (int i = 0; i < s->nb; i++) {
            s->L[i][i] += s->Gamma[i] + s->Gamma[s->nb + i];
            s->L[s->nv][i] += s->gamma[i] - s->gamma[s->nb + i];
        }
        if (k < N)
            for (int r = 0; r <= s->nv; r++) /* SYRK_LN_MN(AL, BAbt) :652 */
                for (int c = 0; c <= r && c < s->nv; c++) {
                    double v = 0;
                    for (int j = 0; j < NX; j++) v += AL[r][j] * s->M[c][j];
                    s->L[r][c] += v;
                }
        potrf_l_mn(s->nv + 1, s->nu, s->L); /* only the nu input columns are factorised :653 */
        for (int r = 0; r < s->nx; r++) /* P = Sxx - Ls Ls' (SYRK -1, :655), symmetrised (TRTR_L) */
            for (int c = 0; c <= r; c++) {
                double v = s->L[s->nu + r][s->nu + c];
                for (int j = 0; j < s->nu; j++) v -= s->L[s->nu + r][j] * s->L[s->nu + c][j];
                s->Pm[r][c] = s->Pm[c][r] = v;
            }
        for (int c = 0; c < s->nx; c++) {
            double v = s->L[s->nv][s->nu + c];
            for (int j = 0; j < s->nu; j++) v -= s->L[s->nv][j] * s->L[s->nu + c][j];
            s->pv[c] = v;
        }
    }
    for (int k = N; k >= 0 && !g_classical; k--) {
        stage *s = &w->s[k];
        double AL[NVM + 1][NX];
        if (k < N) {
            stage *n = &w->s[k + 1];
            /* AL = [B';A';res_b'] * Lxx(k+1)   (TRMM_RLNN) */
            for (int r = 0; r <= s->nv; r++)
                for (int c = 0; c < NX; c++) {
                    double v = 0;
                    for (int j = c; j < NX; j++)
                        v += (r < s->nv ? s->M[r][j] : s->res_b[j]) * n->L[n->nu + j][n->nu + c];
                    AL[r][c] = v;
                }
            /* Pb = Lxx * AL[last]'  (TRMV_LNN) */
            for (int r = 0; r < NX; r++) {
                double v = 0;
                for (int c = 0; c <= r; c++) v += n->L[n->nu + r][n->nu + c] * AL[s->nv][c];
                s->Pb[r] = v;
            }
            for (int c = 0; c < NX; c++) AL[s->nv][c] += n->L[n->nv][n->nu + c];
        }
        memset(s->L, 0, sizeof s->L);
        for (int i = 0; i < s->nv; i++) { s->L[i][i] = s->H[i] + REG_PRIM; s->L[s->nv][i] = s->res_g[i]; }
        if (s->dense)
            for (int i = 0; i < s->nv; i++)
                for (int j = 0; j <= i; j++) s->L[i][j] = s->Hm[i][j] + (i == j ? REG_PRIM : 0.0);
        for (int i = 0; i < s->nb; i++) {
            s->L[i][i] += s->Gamma[i] + s->Gamma[s->nb + i];
            s->L[s->nv][i] += s->gamma[i] - s->gamma[s->nb + i];
        }
        if (k < N)
            for (int r = 0; r <= s->nv; r++)
                for (int c = 0; c <= r && c < s->nv; c++) {
                    double v = 0;
                    for (int j = 0; j < NX; j++) v += AL[r][j] * AL[c][j];
                    s->L[r][c] += v;
                }
        potrf_l_mn(s->nv + 1, s->nv, s->L);
    }
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < s->nu; i++) s->dux[i] = -s->L[s->nv][i];
        for (int i = 0; i < s->nx; i++) pl[k][i] = g_classical ? s->pv[i] : s->L[s->nv][s->nu + i];
    }
    ric_forward(w, pl, 0, 1);
    compute_lam_t(w, 0);
    free(pl);
}

/* OCP_QP_SOLVE_KKT_STEP, square-root branch: x_ocp_qp_kkt.c:1147-1292.
 * itref=0: rhs = res_*, sol = d*   (qp_step / sol_step)
 * itref=1: rhs = r*2,   sol = d*2  (qp_itref / sol_itref), use_Pb = 0 */
static void solve_kkt_step(ipm *w, int use_Pb, int itref)
{
    int N = w->N;
    double(*pl)[NX] = malloc(sizeof(double[NX]) * (N + 1));
    compute_Gamma_gamma(w, 0, itref);
    for (int k = N; k >= 0; k--) {
        stage *s = &w->s[k];
        double *dux = itref ? s->dux2 : s->dux;
        const double *rg = itref ? s->rg2 : s->res_g, *rb = itref ? s->rb2 : s->res_b;
        for (int i = 0; i < s->nv; i++) dux[i] = rg[i];
        for (int i = 0; i < s->nb; i++) dux[i] += s->gamma[i] - s->gamma[s->nb + i];
        if (k < N) {
            stage *n = &w->s[k + 1];
            const double *ndux = itref ? n->dux2 : n->dux;
            double tmp[NX];
            if (use_Pb) {
                for (int i = 0; i < NX; i++) tmp[i] = ndux[n->nu + i] + s->Pb[i];
            } else if (g_classical) {
                for (int r = 0; r < NX; r++) {
                    double v = ndux[n->nu + r];
                    for (int c = 0; c < NX; c++) v += n->Pm[r][c] * rb[c];
                    tmp[r] = v;
                }
            } else {
                double t2[NX];
                for (int c = 0; c < NX; c++) {
                    double v = 0;
                    for (int r = c; r < NX; r++) v += n->L[n->nu + r][n->nu + c] * rb[r];
                    t2[c] = v;
                }
                for (int r = 0; r < NX; r++) {
                    double v = 0;
                    for (int c = 0; c <= r; c++) v += n->L[n->nu + r][n->nu + c] * t2[c];
                    tmp[r] = v + ndux[n->nu + r];
                }
            }
            for (int r = 0; r < s->nv; r++) {
                double v = 0;
                for (int c = 0; c < NX; c++) v += s->M[r][c] * tmp[c];
                dux[r] += v;
            }
        }
        /* TRSV_LNN_MN(nv, nu) */
        for (int i = 0; i < s->nu; i++) {
            double v = dux[i];
            for (int c = 0; c < i; c++) v -= s->L[i][c] * dux[c];
            dux[i] = v * (1.0 / s->L[i][i]);
        }
        for (int i = s->nu; i < s->nv; i++) {
            double v = dux[i];
            for (int c = 0; c < s->nu; c++) v -= s->L[i][c] * dux[c];
            dux[i] = v;
        }
    }
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        double *dux = itref ? s->dux2 : s->dux;
        for (int i = 0; i < s->nx; i++) pl[k][i] = dux[s->nu + i];
        for (int i = 0; i < s->nu; i++) dux[i] = -dux[i];
    }
    ric_forward(w, pl, itref, 0);
    compute_lam_t(w, itref);
    free(pl);
}

/* OCP_QP_RES_COMPUTE: x_ocp_qp_res.c:334-470 */
static void res_compute(ipm *w)
{
    int N = w->N;
    double mu = 0;
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < s->nv; i++) {
            double v = s->H[i] * s->ux[i];
            if (s->dense) { v = 0; for (int j = 0; j < s->nv; j++) v += s->Hm[i][j] * s->ux[j]; }   /* SYMV_L */
            s->res_g[i] = v + s->rq[i];
        }
        if (k > 0) for (int i = 0; i < s->nx; i++) s->res_g[s->nu + i] -= w->s[k - 1].pi[i];
        for (int i = 0; i < s->nb; i++) {
            s->res_g[i] += s->lam[s->nb + i] - s->lam[i];
            s->res_d[i] = s->d[i] + s->t[i] - s->ux[i];
            s->res_d[s->nb + i] = s->d[s->nb + i] + s->t[s->nb + i] + s->ux[i];
        }
        if (k < N) {
            stage *n = &w->s[k + 1];
            for (int c = 0; c < NX; c++) {
                double v = s->b[c] - n->ux[n->nu + c];
                for (int r = 0; r < s->nv; r++) v += s->M[r][c] * s->ux[r];
                s->res_b[c] = v;
            }
            for (int r = 0; r < s->nv; r++) {
                double v = 0;
                for (int c = 0; c < NX; c++) v += s->M[r][c] * s->pi[c];
                s->res_g[r] += v;
            }
        }
        for (int i = 0; i < 2 * s->nb; i++) { s->res_m[i] = s->lam[i] * s->t[i]; mu += s->res_m[i]; }
    }
    w->mu = mu * (1.0 / w->nc);
}

/* OCP_QP_RES_COMPUTE_LIN: x_ocp_qp_res.c:474-598; rhs = qp_step (= current res_*),
 * solution = current step, result in r*2 */
static void res_compute_lin(ipm *w)
{
    int N = w->N;
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < s->nv; i++) {
            double v = s->H[i] * s->dux[i];
            if (s->dense) { v = 0; for (int j = 0; j < s->nv; j++) v += s->Hm[i][j] * s->dux[j]; }
            s->rg2[i] = v + s->res_g[i];
        }
        if (k > 0) for (int i = 0; i < s->nx; i++) s->rg2[s->nu + i] -= w->s[k - 1].dpi[i];
        for (int i = 0; i < s->nb; i++) {
            s->rg2[i] += s->dlam[s->nb + i] - s->dlam[i];
            s->rd2[i] = s->res_d[i] + s->dt[i] - s->dux[i];
            s->rd2[s->nb + i] = s->res_d[s->nb + i] + s->dt[s->nb + i] + s->dux[i];
        }
        if (k < N) {
            stage *n = &w->s[k + 1];
            for (int c = 0; c < NX; c++) {
                double v = s->res_b[c] - n->dux[n->nu + c];
                for (int r = 0; r < s->nv; r++) v += s->M[r][c] * s->dux[r];
                s->rb2[c] = v;
            }
            for (int r = 0; r < s->nv; r++) {
                double v = 0;
                for (int c = 0; c < NX; c++) v += s->M[r][c] * s->dpi[c];
                s->rg2[r] += v;
            }
        }
        for (int i = 0; i < 2 * s->nb; i++)
            s->rm2[i] = s->res_m[i] + s->lam[i] * s->dt[i] + s->dlam[i] * s->t[i];
    }
}

static void inf_norms(ipm *w, int itref, double *out)
{
    double g = 0, b = 0, d = 0, m = 0;
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        const double *rg = itref ? s->rg2 : s->res_g, *rb = itref ? s->rb2 : s->res_b;
        const double *rd = itref ? s->rd2 : s->res_d, *rm = itref ? s->rm2 : s->res_m;
        for (int i = 0; i < s->nv; i++) g = fmax(g, fabs(rg[i]));
        if (k < w->N) for (int i = 0; i < NX; i++) b = fmax(b, fabs(rb[i]));
        for (int i = 0; i < 2 * s->nb; i++) { d = fmax(d, fabs(rd[i])); m = fmax(m, fabs(rm[i])); }
    }
    out[0] = g; out[1] = b; out[2] = d; out[3] = m;
}

/* COMPUTE_ALPHA_QP: x_core_qp_ipm_aux.c:146-216 (single step length, split_step 0) */
static void compute_alpha(ipm *w)
{
    double ap = -1.0, ad = -1.0;
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < 2 * s->nb; i++) {
            if (ad * s->dlam[i] > s->lam[i]) ad = s->lam[i] / s->dlam[i];
            if (ap * s->dt[i] > s->t[i]) ap = s->t[i] / s->dt[i];
        }
    }
    double a = ap > ad ? ap : ad;
    w->alpha = -a;
}

static void compute_mu_aff(ipm *w)
{
    double mu = 0;
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < 2 * s->nb; i++)
            mu += (s->lam[i] + w->alpha * s->dlam[i]) * (s->t[i] + w->alpha * s->dt[i]);
    }
    w->mu_aff = mu * (1.0 / w->nc);
}

static int itref_converged(const ipm *w)
{
    const double *n = w->res2_max, *r = w->res_max;
    return (n[0] < RES_G_MAX || n[0] < 1e-3 * r[0]) && (n[1] < RES_B_MAX || n[1] < 1e-3 * r[1]) &&
           (n[2] < RES_D_MAX || n[2] < 1e-3 * r[2]) && (n[3] < RES_M_MAX || n[3] < 1e-3 * r[3]);
}

/* OCP_QP_IPM_DELTA_STEP: x_ocp_qp_ipm.c:1943-2405 */
static void delta_step(ipm *w, cfo_info *info)
{
    int N = w->N;
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < 2 * s->nb; i++) { s->res_m_bkp[i] = s->res_m[i]; s->res_m[i] = s->res_m_bkp[i] - TAU_MIN; }
    }
    fact_solve_kkt_step(w);
    /* lq_fact==1: linear-system residual decides on an LQ refactorisation (:2011-2059).
     * Never taken on this OCP; detected and counted, not restated. */
    res_compute_lin(w);
    inf_norms(w, 1, w->res2_max);
    if ((w->res2_max[0] == 0.0 && isnan(w->s[0].rg2[0])) || w->res2_max[0] > 1e-5 || w->res2_max[1] > 1e-5 ||
        w->res2_max[2] > 1e-5 || w->res2_max[3] > 1e-5)
        info->n_lq_flag++;

    compute_alpha(w);
    compute_mu_aff(w);
    double tmp = w->mu_aff / w->mu;
    w->sigma = tmp * tmp * tmp;
    double sigma_mu = w->sigma * w->mu;
    sigma_mu = sigma_mu > TAU_MIN ? sigma_mu : TAU_MIN;
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < 2 * s->nb; i++) s->res_m[i] = s->res_m_bkp[i] + s->dt[i] * s->dlam[i] - sigma_mu;
    }
    solve_kkt_step(w, 1, 0);
    compute_alpha(w);
    /* conditional predictor-corrector (:2230-2273) */
    double mu_aff0 = w->mu_aff;
    compute_mu_aff(w);
    if (w->mu_aff > 2.0 * mu_aff0) {
        for (int k = 0; k <= N; k++) {
            stage *s = &w->s[k];
            for (int i = 0; i < 2 * s->nb; i++) s->res_m[i] = s->res_m_bkp[i] - sigma_mu;
        }
        solve_kkt_step(w, 1, 0);
        compute_alpha(w);
    }
    /* iterative refinement on the corrector (:2275-2366) */
    int refined = 0, it;
    for (it = 0; it < ITREF_CORR_MAX; it++) {
        res_compute_lin(w);
        inf_norms(w, 1, w->res2_max);
        if (itref_converged(w)) break;
        solve_kkt_step(w, 0, 1);
        refined = 1;
        info->n_itref++;
        for (int k = 0; k <= N; k++) {
            stage *s = &w->s[k];
            for (int i = 0; i < s->nv; i++) s->dux[i] += s->dux2[i];
            if (k < N) for (int i = 0; i < NX; i++) s->dpi[i] += s->dpi2[i];
            for (int i = 0; i < 2 * s->nb; i++) { s->dlam[i] += s->dlam2[i]; s->dt[i] += s->dt2[i]; }
        }
    }
    if (refined) compute_alpha(w);
    /* UPDATE_VAR_QP: x_core_qp_ipm_aux.c:220-325 */
    double alpha = w->alpha;
    if (alpha < 1.0) alpha = alpha * ((1.0 - alpha) * 0.99 + alpha * 0.9999999);
    for (int k = 0; k <= N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < s->nv; i++) s->ux[i] += alpha * s->dux[i];
        if (k < N) for (int i = 0; i < NX; i++) s->pi[i] += alpha * s->dpi[i];
        for (int i = 0; i < 2 * s->nb; i++) {
            s->lam[i] += alpha * s->dlam[i];
            s->lam[i] = s->lam[i] <= LAM_MIN ? LAM_MIN : s->lam[i];
            s->t[i] += alpha * s->dt[i];
            s->t[i] = s->t[i] <= T_MIN ? T_MIN : s->t[i];
        }
    }
}

/* OCP_QP_INIT_VAR, scheme 1: x_ocp_qp_ipm.c:1491-1530,1636-1769 */
static void init_var(ipm *w)
{
    const double thr0 = 1e-1;
    for (int k = 0; k <= w->N; k++) {
        stage *s = &w->s[k];
        for (int i = 0; i < s->nv; i++) s->ux[i] = 0.0;
        for (int i = 0; i < NX; i++) s->pi[i] = 0.0;
        for (int i = 0; i < s->nb; i++) {
            double *tl = &s->t[i], *tu = &s->t[s->nb + i];
            *tl = s->ux[i] - s->d[i];
            *tu = -s->ux[i] - s->d[s->nb + i];
            if (*tl < thr0) {
                if (*tu < thr0) {
                    s->ux[i] = 0.5 * (s->d[i] - s->d[s->nb + i]);
                    *tl = thr0;
                    *tu = thr0;
                } else {
                    *tl = thr0;
                    s->ux[i] = s->d[i] + thr0;
                }
            } else if (*tu < thr0) {
                *tu = thr0;
                s->ux[i] = -s->d[s->nb + i] - thr0;
            }
        }
        for (int i = 0; i < 2 * s->nb; i++) s->lam[i] = MU0 / s->t[i];
    }
}

/* OCP_QP_IPM_SOLVE, delta formulation: x_ocp_qp_ipm.c:2409-2759 */
static void ipm_solve(ipm *w, cfo_info *info)
{
    init_var(w);
    w->alpha = 1.0;
    res_compute(w);
    inf_norms(w, 0, w->res_max);
    int kk;
    for (kk = 0; kk < ITER_MAX && w->alpha > ALPHA_MIN &&
                 (w->res_max[0] > RES_G_MAX || w->res_max[1] > RES_B_MAX || w->res_max[2] > RES_D_MAX ||
                  fabs(w->res_max[3] - TAU_MIN) > RES_M_MAX);
         kk++) {
        delta_step(w, info);
        res_compute(w);
        inf_norms(w, 0, w->res_max);
    }
    info->qp_iter = kk;
    if (kk == ITER_MAX) info->qp_status = 1;
    else if (w->alpha <= ALPHA_MIN) info->qp_status = 2;
    else if (isnan(w->mu)) info->qp_status = 3;
    else info->qp_status = 0;
    for (int i = 0; i < 4; i++) info->res[i] = w->res_max[i];
}

/* Multipliers of the iterate after the next cfo_rti calls with this N, in the layout of oracle/ref_harness.c:
 * cfref_get_multipliers (pi[N][13]; lam / t: stage 0 2*17 = [lbu lbx | ubu ubx], stages 1..N-1 2*4).  Global, test use
 * only; N = 0 switches the recording off. */
static int g_mult_N = 0;
static double g_mult_pi[CFO_MAX_N * NX], g_mult_lam[2 * NV + 2 * NU * CFO_MAX_N], g_mult_t[2 * NV + 2 * NU * CFO_MAX_N];
void cfo_record_multipliers(int N) { g_mult_N = (N > 0 && N <= CFO_MAX_N) ? N : 0; }
void cfo_last_multipliers(double *pi, double *lam, double *t)
{
    const int N = g_mult_N;
    if (pi) memcpy(pi, g_mult_pi, sizeof(double) * N * NX);
    if (lam) memcpy(lam, g_mult_lam, sizeof(double) * (2 * NV + 2 * NU * (N - 1)));
    if (t) memcpy(t, g_mult_t, sizeof(double) * (2 * NV + 2 * NU * (N - 1)));
}

int cfo_rti(int N, double Ts, const cfo_params *p_, const double *x0, const double *yref,
            const double *yref_e, double *x, double *u, cfo_info *info_, double *dux_out, double *dpi_out)
{
    cfo_params pd;
    cfo_info li;
    cfo_info *info = info_ ? info_ : &li;
    memset(info, 0, sizeof *info);
    if (!p_) { cfo_default_params(&pd); p_ = &pd; }

    /* --- preparation + QP vectors */
    double *BAbt = malloc(sizeof(double) * N * NV * NX), *b = malloc(sizeof(double) * N * NX);
    double *rqz = malloc(sizeof(double) * (N * NV + NX));
    double *dl = malloc(sizeof(double) * (NV + NU * N)), *du = malloc(sizeof(double) * (NV + NU * N));
    cfo_linearize(N, Ts, p_, x0, yref, yref_e, x, u, BAbt, b, rqz, dl, du);

    /* --- reduce: eliminate the 13 equality-bounded x0 (x_ocp_qp_red.c:268-455) */
    ipm w;
    w.N = N;
    w.nc = 2 * NU * N;
    w.s = calloc(N + 1, sizeof(stage));
    double xbar[NX];
    for (int i = 0; i < NX; i++) xbar[i] = dl[NU + i]; /* lower-bound entry is the value (:310-315) */
    for (int k = 0; k <= N; k++) {
        stage *s = &w.s[k];
        s->nu = k < N ? NU : 0;
        s->nx = k == 0 ? 0 : NX;
        s->nv = s->nu + s->nx;
        s->nb = k < N ? NU : 0;
        /* Hessian: scaling * (sqrt(w))^2, cost_ls.c:739-772 */
        for (int i = 0; i < s->nu; i++) { double r = sqrt(WKS(p_, k, NX + i)); s->H[i] = DTK(k) * (r * r); }
        for (int i = 0; i < s->nx; i++) {
            double r = sqrt(WK(p_, k, N, i));
            s->H[s->nu + i] = (k < N ? DTK(k) : 1.0) * (r * r);
        }
        const double *g = rqz + NV * k;
        if (k == 0) {
            for (int i = 0; i < NU; i++) s->rq[i] = g[i]; /* diagonal RSQ: no cross term from xbar (:354) */
        } else {
            for (int i = 0; i < s->nv; i++) s->rq[i] = g[i];
        }
        if (k < g_wd_n) {
            /* full weight matrix: dense stage Hessian scaling * (Cyt W_chol)(Cyt W_chol)'; the eliminated x_0 leaves the
             * cross term S xbar in the stage-0 gradient (SYMV_L with the masked vector, x_ocp_qp_red.c:354) */
            const double *Hd = g_hd + 289 * k;
            const double sc = k < N ? DTK(k) : 1.0;
            const int off = k == N ? NU : 0;   /* first [u;x] index among the stage's variables */
            s->dense = 1;
            for (int i = 0; i < s->nv; i++)
                for (int j = 0; j < s->nv; j++) s->Hm[i][j] = sc * Hd[YIDX(off + i) * 17 + YIDX(off + j)];
            if (k == 0)
                for (int i = 0; i < NU; i++) {
                    double v = 0;
                    for (int j = 0; j < NX; j++) v += (sc * Hd[YIDX(i) * 17 + j]) * xbar[j];
                    s->rq[i] = v + g[i];
                }
        }
        if (k < N) {
            const double *M = BAbt + (size_t) k * NV * NX;
            if (k == 0) {
                for (int r = 0; r < NU; r++) for (int c = 0; c < NX; c++) s->M[r][c] = M[r * NX + c];
                for (int c = 0; c < NX; c++) {
                    double v = 0; /* b0 += A0 xbar  (GEMV_T over the masked vector, :330) */
                    for (int r = 0; r < NX; r++) v += M[(NU + r) * NX + c] * xbar[r];
                    s->b[c] = v + b[c];
                }
            } else {
                for (int r = 0; r < NV; r++) for (int c = 0; c < NX; c++) s->M[r][c] = M[r * NX + c];
                for (int c = 0; c < NX; c++) s->b[c] = b[NX * k + c];
            }
            int od = k == 0 ? 0 : NV + NU * (k - 1);
            for (int i = 0; i < NU; i++) { s->d[i] = dl[od + i]; s->d[NU + i] = du[od + i]; }
        }
    }

    ipm_solve(&w, info);

    /* --- restore (x_ocp_qp_red.c:723-871) + primal update (ocp_nlp_common.c:2900-2952) */
    if (dux_out) {
        for (int i = 0; i < NU; i++) dux_out[i] = w.s[0].ux[i];
        for (int i = 0; i < NX; i++) dux_out[NU + i] = xbar[i];
        for (int k = 1; k <= N; k++) for (int i = 0; i < w.s[k].nv; i++) dux_out[NV * k + i] = w.s[k].ux[i];
    }
    if (dpi_out) for (int k = 0; k < N; k++) for (int i = 0; i < NX; i++) dpi_out[NX * k + i] = w.s[k].pi[i];
    int status = 0;
    if (info->qp_status == 0 || info->qp_status == 1) {
        if (g_mult_N == N) {
            /* full-step duals (ocp_nlp_common.c:2917-2925) in the layout ocp_nlp_out_get hands them out; the multipliers
             * of the eliminated x_0 = x0 from the stationarity of the unreduced stage 0 (x_ocp_qp_red.c:820-840):
             * tmp = q_0 + Q_0 xbar + A_0' pi_0; tmp >= 0 -> lower-bound multiplier, else minus the upper-bound one */
            int ol = 0;
            for (int k = 0; k < N; k++) {
                for (int i = 0; i < NX; i++) g_mult_pi[NX * k + i] = w.s[k].pi[i];
                const int nb = k == 0 ? NV : NU;
                for (int i = 0; i < NU; i++) {
                    g_mult_lam[ol + i] = w.s[k].lam[i]; g_mult_lam[ol + nb + i] = w.s[k].lam[NU + i];
                    g_mult_t[ol + i] = w.s[k].t[i]; g_mult_t[ol + nb + i] = w.s[k].t[NU + i];
                }
                if (k == 0) {
                    const double *M = BAbt, *g = rqz;
                    for (int i = 0; i < NX; i++) {
                        double r = sqrt(WK(p_, 0, N, i));
                        double tmp = g[NU + i] + (DTK(0) * (r * r)) * xbar[i];
                        for (int c = 0; c < NX; c++) tmp += M[(NU + i) * NX + c] * w.s[0].pi[c];
                        g_mult_lam[NU + i] = LAM_MIN; g_mult_lam[NV + NU + i] = LAM_MIN;
                        g_mult_t[NU + i] = T_MIN; g_mult_t[NV + NU + i] = T_MIN;
                        if (tmp >= 0) g_mult_lam[NU + i] = tmp;
                        else g_mult_lam[NV + NU + i] = -tmp;
                    }
                }
                ol += 2 * nb;
            }
        }
        for (int i = 0; i < NU; i++) u[i] += w.s[0].ux[i];
        for (int i = 0; i < NX; i++) x[i] += xbar[i];
        for (int k = 1; k <= N; k++) {
            if (k < N) for (int i = 0; i < NU; i++) u[NU * k + i] += w.s[k].ux[i];
            for (int i = 0; i < NX; i++) x[NX * k + i] += w.s[k].ux[w.s[k].nu + i];
        }
    } else {
        status = 4; /* ACADOS_QP_FAILURE, iterate untouched (ocp_nlp_sqp_rti.c:651-664) */
    }
    free(w.s); free(BAbt); free(b); free(rqz); free(dl); free(du);
    return status;
}

/* ------------------------------------------------------------------ partial condensing
 * d_part_cond_qp_cond / d_part_cond_qp_expand_sol (external/hpipm/cond/x_part_cond.c:505-560,658-742) with the block
 * sizes of PART_COND_QP_COMPUTE_BLOCK_SIZE (:36-54): the first N mod N2 blocks hold floor(N/N2)+1 stages, the others
 * floor(N/N2); the terminal stage stays alone.  A block [k0, k0+bs) becomes ONE stage with state x_k0 and inputs
 * [u_{k0+bs-1}; ...; u_k0] (last stage first, x_cond_aux.c:251-273).  Restated mathematically (the products HPIPM forms
 * in COND_BABT :36-100 / COND_RSQRQ :208-470 / COND_DCTD :840 differ from these sums by rounding only):
 *   x_{k0+j} = G_j' [ubar; xbar; 1],  G_0 = [0; I; 0],  G_{j+1} = G_j A_j' + [E_j B_j'; 0; b_j']
 *   H2 = blkdiag(R_j.., Q_0) + sum_{j>=1} G_j Q_j G_j',  rq2 = [r_j..; q_0] + sum_{j>=1} G_j (q_j + Q_j c_j)
 * where c_j is the constant row of G_j.  Only input boxes exist, so the inequality data are just re-ordered. */
void cfo_block_sizes(int N, int N2, int *bs)
{
    int bs0 = N / N2, ii = 0;
    for (; ii < N - N2 * bs0; ii++) bs[ii] = bs0 + 1;
    for (; ii < N2; ii++) bs[ii] = bs0;
    bs[N2] = 0;
}

static void condense_block(const stage *src, int bs, stage *dst)
{
    const int nu2 = NU * bs, nx2 = src[0].nx, nv2 = nu2 + nx2;
    static double G[2][NVM + 1][NX];
    memset(dst, 0, sizeof *dst);
    dst->nu = nu2; dst->nx = nx2; dst->nv = nv2; dst->nb = nu2; dst->dense = 1;
    memset(G, 0, sizeof G);
    for (int i = 0; i < nx2; i++) G[0][nu2 + i][i] = 1.0;
    for (int i = 0; i < nx2; i++) { dst->Hm[nu2 + i][nu2 + i] = src[0].H[NU + i]; dst->rq[nu2 + i] = src[0].rq[NU + i]; }
    for (int j = 0; j < bs; j++) {
        const stage *sj = &src[j];
        const int off = NU * (bs - 1 - j);
        double (*Gc)[NX] = G[j & 1], (*Gn)[NX] = G[(j + 1) & 1];
        for (int e = 0; e < NU; e++) {
            dst->Hm[off + e][off + e] = sj->H[e];
            dst->rq[off + e] = sj->rq[e];
            dst->d[off + e] = sj->d[e];
            dst->d[nu2 + off + e] = sj->d[NU + e];
        }
        if (j >= 1) {   /* cost of x_{k0+j}: 1/2 x'Qx + q'x with x = G_j' z */
            for (int r = 0; r < nv2; r++) {
                for (int c = 0; c <= r; c++) {
                    double v = 0;
                    for (int i = 0; i < NX; i++) v += Gc[r][i] * sj->H[NU + i] * Gc[c][i];
                    dst->Hm[r][c] += v;
                }
                double g = 0;
                for (int i = 0; i < NX; i++) g += Gc[r][i] * (sj->rq[NU + i] + sj->H[NU + i] * Gc[nv2][i]);
                dst->rq[r] += g;
            }
        }
        /* G_{j+1} = G_j A_j' + [E_j B_j' ; 0 ; b_j'] */
        for (int r = 0; r <= nv2; r++)
            for (int c = 0; c < NX; c++) {
                double v = 0;
                if (sj->nx > 0) for (int i = 0; i < NX; i++) v += Gc[r][i] * sj->M[NU + i][c];
                Gn[r][c] = v;
            }
        for (int e = 0; e < NU; e++) for (int c = 0; c < NX; c++) Gn[off + e][c] += sj->M[e][c];
        for (int c = 0; c < NX; c++) Gn[nv2][c] += sj->b[c];
    }
    for (int r = 0; r < nv2; r++) {
        for (int c = 0; c < NX; c++) dst->M[r][c] = G[bs & 1][r][c];
        for (int c = r + 1; c < nv2; c++) dst->Hm[r][c] = dst->Hm[c][r];   /* filled below the diagonal above */
    }
    for (int r = 0; r < nv2; r++) for (int c = 0; c < r; c++) dst->Hm[c][r] = dst->Hm[r][c];
    for (int c = 0; c < NX; c++) dst->b[c] = G[bs & 1][nv2][c];
}

/* uncondensed, x0-eliminated stages of one RTI step (shared by cfo_rti and cfo_rti_pcond) */
static void build_reduced_stages(int N, double Ts, const cfo_params *p_, const double *BAbt, const double *b, const double *rqz,
                                 const double *dl, const double *du, stage *st, double *xbar)
{
    (void) Ts;
    for (int i = 0; i < NX; i++) xbar[i] = dl[NU + i]; /* lower-bound entry is the value (x_ocp_qp_red.c:310-315) */
    for (int k = 0; k <= N; k++) {
        stage *s = &st[k];
        s->nu = k < N ? NU : 0;
        s->nx = k == 0 ? 0 : NX;
        s->nv = s->nu + s->nx;
        s->nb = k < N ? NU : 0;
        for (int i = 0; i < s->nu; i++) { double r = sqrt(WKS(p_, k, NX + i)); s->H[i] = DTK(k) * (r * r); }
        for (int i = 0; i < s->nx; i++) {
            double r = sqrt(WK(p_, k, N, i));
            s->H[s->nu + i] = (k < N ? DTK(k) : 1.0) * (r * r);
        }
        const double *g = rqz + NV * k;
        if (k == 0) { for (int i = 0; i < NU; i++) s->rq[i] = g[i]; }
        else { for (int i = 0; i < s->nv; i++) s->rq[i] = g[i]; }
        if (k < N) {
            const double *M = BAbt + (size_t) k * NV * NX;
            if (k == 0) {
                for (int r = 0; r < NU; r++) for (int c = 0; c < NX; c++) s->M[r][c] = M[r * NX + c];
                for (int c = 0; c < NX; c++) {
                    double v = 0;
                    for (int r = 0; r < NX; r++) v += M[(NU + r) * NX + c] * xbar[r];
                    s->b[c] = v + b[c];
                }
            } else {
                for (int r = 0; r < NV; r++) for (int c = 0; c < NX; c++) s->M[r][c] = M[r * NX + c];
                for (int c = 0; c < NX; c++) s->b[c] = b[NX * k + c];
            }
            int od = k == 0 ? 0 : NV + NU * (k - 1);
            for (int i = 0; i < NU; i++) { s->d[i] = dl[od + i]; s->d[NU + i] = du[od + i]; }
        }
    }
}

/* One RTI step with the QP partially condensed to N2 stages before the interior-point solve (qp_cond_N = N2 < N in the
 * reference: ocp_qp_partial_condensing.c:457-576, ocp_qp_xcond_solver.c:487-533).  N2 >= N or <= 0: plain cfo_rti. */
int cfo_rti_pcond(int N, double Ts, const cfo_params *p_, int N2, const double *x0, const double *yref,
                  const double *yref_e, double *x, double *u, cfo_info *info_)
{
    if (N2 <= 0 || N2 >= N) return cfo_rti(N, Ts, p_, x0, yref, yref_e, x, u, info_, NULL, NULL);
    cfo_params pd;
    cfo_info li;
    cfo_info *info = info_ ? info_ : &li;
    memset(info, 0, sizeof *info);
    if (!p_) { cfo_default_params(&pd); p_ = &pd; }
    if ((N + N2 - 1) / N2 * NU + NX > NVM) return -1;   /* block too large for the generic stage */
    double *BAbt = malloc(sizeof(double) * N * NV * NX), *b = malloc(sizeof(double) * N * NX);
    double *rqz = malloc(sizeof(double) * (N * NV + NX));
    double *dl = malloc(sizeof(double) * (NV + NU * N)), *du = malloc(sizeof(double) * (NV + NU * N));
    cfo_linearize(N, Ts, p_, x0, yref, yref_e, x, u, BAbt, b, rqz, dl, du);
    stage *st = calloc(N + 1, sizeof(stage));
    double xbar[NX];
    build_reduced_stages(N, Ts, p_, BAbt, b, rqz, dl, du, st, xbar);

    int *bs = malloc(sizeof(int) * (N2 + 1));
    cfo_block_sizes(N, N2, bs);
    ipm w;
    w.N = N2;
    w.nc = 2 * NU * N;
    w.s = calloc(N2 + 1, sizeof(stage));
    int k0 = 0;
    for (int i = 0; i < N2; i++) { condense_block(st + k0, bs[i], &w.s[i]); k0 += bs[i]; }
    w.s[N2] = st[N];
    ipm_solve(&w, info);

    int status = 0;
    if (info->qp_status == 0 || info->qp_status == 1) {
        /* expansion (EXPAND_SOL, x_cond_aux.c:1820-): inner states by the dynamics, then the full step */
        k0 = 0;
        for (int i = 0; i < N2; i++) {
            const stage *c = &w.s[i];
            double xs[NX], xn[NX];
            for (int r = 0; r < NX; r++) xs[r] = c->nx ? c->ux[c->nu + r] : 0.0;
            for (int j = 0; j < bs[i]; j++) {
                const stage *sj = &st[k0 + j];
                const double *uj = c->ux + NU * (bs[i] - 1 - j);
                if (k0 + j == 0) { for (int r = 0; r < NX; r++) x[r] += xbar[r]; }
                else for (int r = 0; r < NX; r++) x[NX * (k0 + j) + r] += xs[r];
                for (int e = 0; e < NU; e++) u[NU * (k0 + j) + e] += uj[e];
                for (int cc = 0; cc < NX; cc++) {
                    double v = sj->b[cc];
                    for (int e = 0; e < NU; e++) v += sj->M[e][cc] * uj[e];
                    if (sj->nx) for (int r = 0; r < NX; r++) v += sj->M[NU + r][cc] * xs[r];
                    xn[cc] = v;
                }
                memcpy(xs, xn, sizeof xs);
            }
            k0 += bs[i];
        }
        for (int r = 0; r < NX; r++) x[NX * N + r] += w.s[N2].ux[r];
    } else {
        status = 4;
    }
    free(w.s); free(st); free(bs); free(BAbt); free(b); free(rqz); free(dl); free(du);
    return status;
}

/* Split real-time iteration (rti_phase 1 then 2, ocp_nlp_sqp_rti.c:495-683): the preparation linearises around the
 * iterate -- the measured state does not enter it -- and the feedback evaluates the bound vectors, i.e. the stage-0
 * equality x_0 = x0, with the data of the moment (ocp_nlp_approximate_qp_vectors_sqp, ocp_nlp_common.c:2258-2292).  The
 * pair therefore equals one full step with the feedback-time x0 (yref is consumed by the preparation and is the same for
 * both); oracle/ref_harness.c:cfref_rti_split runs the reference's own two phases and the tests pin this equality. */
int cfo_rti_split(int N, double Ts, const cfo_params *p, const double *x0_prep, const double *x0_fb, const double *yref,
                  const double *yref_e, double *x, double *u, cfo_info *info)
{
    (void) x0_prep;
    return cfo_rti(N, Ts, p, x0_fb, yref, yref_e, x, u, info, NULL, NULL);
}

void cfo_batch(int N, double Ts, const cfo_params *p, int n_rti, int n, const double *x0,
               const double *yref, const double *yref_e, double *x, double *u, int *status, int *qp_iter)
{
    for (int i = 0; i < n; i++) {
        cfo_info info;
        int st = 0;
        for (int r = 0; r < n_rti; r++)
            st = cfo_rti(N, Ts, p, x0 + NX * (size_t) i, yref + (size_t) N * NV * i, yref_e + NX * (size_t) i,
                         x + (size_t) (N + 1) * NX * i, u + (size_t) N * NU * i, &info, NULL, NULL);
        if (status) status[i] = st;
        if (qp_iter) qp_iter[i] = info.qp_iter;
    }
}

void cfo_batch_pcond(int N, double Ts, const cfo_params *p, int N2, int n_rti, int n, const double *x0,
                     const double *yref, const double *yref_e, double *x, double *u, int *status, int *qp_iter)
{
    for (int i = 0; i < n; i++) {
        cfo_info info;
        int st = 0;
        for (int r = 0; r < n_rti; r++)
            st = cfo_rti_pcond(N, Ts, p, N2, x0 + NX * (size_t) i, yref + (size_t) N * NV * i, yref_e + NX * (size_t) i,
                               x + (size_t) (N + 1) * NX * i, u + (size_t) N * NU * i, &info);
        if (status) status[i] = st;
        if (qp_iter) qp_iter[i] = info.qp_iter;
    }
}
