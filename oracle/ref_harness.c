/* TEST INFRASTRUCTURE ONLY (oracle/_ref): drives the reference's OWN acados /
 * HPIPM / BLASFEO code (compiled from /root/reference by oracle/Makefile) on
 * the Crazyflie OCP, so that the product and the plain-C oracle can be checked
 * against the real reference and so that bench.py has a "reference" CPU arm.
 *
 * The generated glue the reference normally gets from CasADi + Tera
 * (c_generated_code/, not in the tree) is replaced by the call sequence below,
 * which follows the template section by section:
 *   acados_template/c_templates_tera/acados_solver.in.c
 *     plan :159-199, dims :204-380, nlp_in :879-1569, opts :2028-2317,
 *     nlp_out :2323-2352, precompute :2381-2435
 * with the values of crazyflie_controller/scripts/crazyflie_full_model/generate_c_code.py:41-146.
 * The model enters as an external_function_generic whose first member is the
 * evaluate pointer (acados/utils/external_function_generic.h:75-83).
 *
 * Never linked into, or called from, the product library. */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "acados/ocp_nlp/ocp_nlp_common.h"
#include "acados/ocp_qp/ocp_qp_hpipm.h"
#include "acados/ocp_qp/ocp_qp_xcond_solver.h"
#include "acados/utils/external_function_generic.h"
#include "acados_c/external_function_interface.h"
#include "acados_c/ocp_nlp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "hpipm_d_ocp_qp_ipm.h"

/* Second model ("other nx, nu compile without touching kernels"): -DCFREF_MODEL_PENDULUM builds the same
 * harness for the pendulum on a cart (nx = 4, nu = 1) with the CasADi-generated external functions the reference ships
 * in acados/examples/c/pendulum_model/ (compiled from there by oracle/Makefile) and the OCP data of
 * acados/examples/acados_python/tests/test_ocp_setting.py:150-205 (N, Tf, Q, R, Fmax, x0 are the caller's / those values). */
#if defined(CFREF_MODEL_PENDULUM)
#define NX 4
#define NU 1
#define NY 5
int pendulum_ode_expl_vde_forw(const double **arg, double **res, int *iw, double *w, void *mem);
int pendulum_ode_expl_ode_fun(const double **arg, double **res, int *iw, double *w, void *mem);
/* The checked-in C functions order the state as [x1, v1, theta, dtheta] (export_pendulum_ode_model.m:62), the Python model
 * the spec is generated from as [x1, theta, v1, dtheta] (pendulum_model.py:50-55): same dynamics, states 1 and 2 swapped.
 * The wrappers present the Python order (the swap is its own inverse; matrices are column-major). */
static const int PP[4] = {0, 2, 1, 3};
static void cf_ref_vde_forw(const double *x, const double *Sx, const double *Su, const double *u, double *f, double *dSx, double *dSu)
{
    double xc[4], Sxc[16], Suc[4], fc[4], dSxc[16], dSuc[4];
    for (int i = 0; i < 4; i++) {
        xc[PP[i]] = x[i];
        Suc[PP[i]] = Su[i];
        for (int j = 0; j < 4; j++) Sxc[PP[i] + 4 * PP[j]] = Sx[i + 4 * j];
    }
    const double *arg[4] = {xc, Sxc, Suc, u};
    double *res[3] = {fc, dSxc, dSuc};
    int iw[64];
    double w[512];
    pendulum_ode_expl_vde_forw(arg, res, iw, w, 0);
    for (int i = 0; i < 4; i++) {
        f[i] = fc[PP[i]];
        dSu[i] = dSuc[PP[i]];
        for (int j = 0; j < 4; j++) dSx[i + 4 * j] = dSxc[PP[i] + 4 * PP[j]];
    }
}
static void cf_ref_ode(const double *x, const double *u, double *f)
{
    double xc[4], fc[4];
    for (int i = 0; i < 4; i++) xc[PP[i]] = x[i];
    const double *arg[2] = {xc, u};
    double *res[1] = {fc};
    int iw[64];
    double w[512];
    pendulum_ode_expl_ode_fun(arg, res, iw, w, 0);
    for (int i = 0; i < 4; i++) f[i] = fc[PP[i]];
}
#else
#include "cf_model_ref.h"

#define NX 13
#define NU 4
#define NY 17
#endif
#define NV (NX + NU)

typedef struct
{
    int N;
    double Ts;
    ocp_nlp_plan_t *plan;
    ocp_nlp_config *config;
    ocp_nlp_dims *dims;
    ocp_nlp_in *in;
    ocp_nlp_out *out;
    void *opts;
    ocp_nlp_solver *solver;
    external_function_generic vde, ode;
} cfref;

static void vde_eval(void *self, ext_fun_arg_t *tin, void **in, ext_fun_arg_t *tout, void **out)
{
    (void) self; (void) tin; (void) tout;
    cf_ref_vde_forw((const double *) in[0], (const double *) in[1], (const double *) in[2],
                    (const double *) in[3], (double *) out[0], (double *) out[1], (double *) out[2]);
}
static void ode_eval(void *self, ext_fun_arg_t *tin, void **in, ext_fun_arg_t *tout, void **out)
{
    (void) self; (void) tin; (void) tout;
    cf_ref_ode((const double *) in[0], (const double *) in[1], (double *) out[0]);
}

void cfref_destroy(void *h_)
{
    cfref *h = h_;
    if (!h) return;
    if (h->solver) ocp_nlp_solver_destroy(h->solver);
    if (h->out) ocp_nlp_out_destroy(h->out);
    if (h->opts) ocp_nlp_solver_opts_destroy(h->opts);
    if (h->in) ocp_nlp_in_destroy(h->in);
    if (h->dims) ocp_nlp_dims_destroy(h->dims);
    if (h->config) ocp_nlp_config_destroy(h->config);
    if (h->plan) ocp_nlp_plan_destroy(h->plan);
    free(h);
}

/* cond_N <= 0 means "what generate_c_code.py yields": qp_cond_N = N. */
void *cfref_create_dt(int N, const double *dt, int cond_N);
void *cfref_create(int N, double Ts, int cond_N)
{
    double *dt = malloc(sizeof(double) * N);
    for (int i = 0; i < N; i++) dt[i] = Ts;
    void *h = cfref_create_dt(N, dt, cond_N);
    free(dt);
    return h;
}

/* The same with one time step per shooting interval, as crazyflie_acados_create_with_discretization does it
 * (acados_template/c_templates_tera/acados_solver.in.c:133-153,854-892): "Ts" and the cost "scaling" of interval i. */
void *cfref_create_dt(int N, const double *dt, int cond_N)
{
    const double Ts = dt[0];
    cfref *h = calloc(1, sizeof(cfref));
    h->N = N;
    h->Ts = Ts;
    if (cond_N <= 0 || cond_N > N) cond_N = N;

    ocp_nlp_plan_t *plan = h->plan = ocp_nlp_plan_create(N);
    plan->nlp_solver = SQP_RTI;
    plan->ocp_qp_solver_plan.qp_solver = PARTIAL_CONDENSING_HPIPM;
    plan->regularization = NO_REGULARIZE;
    for (int i = 0; i <= N; i++) { plan->nlp_cost[i] = LINEAR_LS; plan->nlp_constraints[i] = BGH; }
    for (int i = 0; i < N; i++) { plan->nlp_dynamics[i] = CONTINUOUS_MODEL; plan->sim_solver_plan[i].sim_solver = ERK; }
    ocp_nlp_config *config = h->config = ocp_nlp_config_create(*plan);

    int *nx = calloc(N + 1, sizeof(int)), *nu = calloc(N + 1, sizeof(int)), *zz = calloc(N + 1, sizeof(int));
    for (int i = 0; i <= N; i++) { nx[i] = NX; nu[i] = i < N ? NU : 0; }
    ocp_nlp_dims *dims = h->dims = ocp_nlp_dims_create(config);
    ocp_nlp_dims_set_opt_vars(config, dims, "nx", nx);
    ocp_nlp_dims_set_opt_vars(config, dims, "nu", nu);
    ocp_nlp_dims_set_opt_vars(config, dims, "nz", zz);
    ocp_nlp_dims_set_opt_vars(config, dims, "ns", zz);
    ocp_nlp_dims_set_opt_vars(config, dims, "np", zz);
    for (int i = 0; i <= N; i++) {
        int nbx = i == 0 ? NX : 0, nbu = i < N ? NU : 0, z = 0, nbxe = nbx, ny = i < N ? NY : NX;
        ocp_nlp_dims_set_constraints(config, dims, i, "nbx", &nbx);
        ocp_nlp_dims_set_constraints(config, dims, i, "nbu", &nbu);
        ocp_nlp_dims_set_constraints(config, dims, i, "nsbx", &z);
        ocp_nlp_dims_set_constraints(config, dims, i, "nsbu", &z);
        ocp_nlp_dims_set_constraints(config, dims, i, "ng", &z);
        ocp_nlp_dims_set_constraints(config, dims, i, "nsg", &z);
        ocp_nlp_dims_set_constraints(config, dims, i, "nbxe", &nbxe);
        ocp_nlp_dims_set_constraints(config, dims, i, "nh", &z);
        ocp_nlp_dims_set_constraints(config, dims, i, "nsh", &z);
        ocp_nlp_dims_set_cost(config, dims, i, "ny", &ny);
    }
    free(nx); free(nu); free(zz);

    h->vde.evaluate = &vde_eval;
    h->ode.evaluate = &ode_eval;
    ocp_nlp_in *in = h->in = ocp_nlp_in_create(config, dims);
    for (int i = 0; i < N; i++) {
        double dti = dt[i];
        ocp_nlp_in_set(config, dims, in, i, "Ts", &dti);
        ocp_nlp_cost_model_set(config, dims, in, i, "scaling", &dti);
        ocp_nlp_dynamics_model_set(config, dims, in, i, "expl_vde_forw", &h->vde);
        ocp_nlp_dynamics_model_set(config, dims, in, i, "expl_ode_fun", &h->ode);
    }

#if defined(CFREF_MODEL_PENDULUM)
    /* test_ocp_setting.py:160-163,199-205: Q = 2 diag(1e3, 1e3, 1e-2, 1e-2), R = 2 diag(1e-2), W_e = Q, yref = 0 */
    double Qd[NX] = {2e3, 2e3, 2e-2, 2e-2};
    const double Rd = 2e-2, WNscale = 1.0;
    double yref[NY] = {0, 0, 0, 0, 0};
#else
    /* generate_c_code.py:50-129 */
    const double g0 = 9.8066, mq = 33e-3, Ct = 3.25e-4;
    double hov = sqrt((mq * g0) / (4 * Ct));
    double Qd[NX] = {120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0};
    const double Rd = 0.06, WNscale = 50.0;
    double yref[NY] = {0, 0, 0.5, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, hov, hov, hov, hov};
#endif
    double W[NY * NY] = {0}, WN[NX * NX] = {0}, Vx[NY * NX] = {0}, Vu[NY * NU] = {0}, VxN[NX * NX] = {0};
    for (int i = 0; i < NX; i++) { W[i + NY * i] = Qd[i]; WN[i + NX * i] = WNscale * Qd[i]; Vx[i + NY * i] = 1; VxN[i + NX * i] = 1; }
    for (int i = 0; i < NU; i++) { W[(NX + i) + NY * (NX + i)] = Rd; Vu[(NX + i) + NY * i] = 1; }
    for (int i = 0; i < N; i++) {
        ocp_nlp_cost_model_set(config, dims, in, i, "W", W);
        ocp_nlp_cost_model_set(config, dims, in, i, "Vx", Vx);
        ocp_nlp_cost_model_set(config, dims, in, i, "Vu", Vu);
        ocp_nlp_cost_model_set(config, dims, in, i, "yref", yref);
    }
    ocp_nlp_cost_model_set(config, dims, in, N, "W", WN);
    ocp_nlp_cost_model_set(config, dims, in, N, "Vx", VxN);
    ocp_nlp_cost_model_set(config, dims, in, N, "yref", yref);

    int idxbx0[NX], idxbu[NU];
    for (int i = 0; i < NX; i++) idxbx0[i] = i;
    for (int i = 0; i < NU; i++) idxbu[i] = i;
#if defined(CFREF_MODEL_PENDULUM)
    double x0[NX] = {0, 3.14159265358979323846, 0, 0};
#else
    double x0[NX] = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
    ocp_nlp_constraints_model_set(config, dims, in, 0, "idxbx", idxbx0);
    ocp_nlp_constraints_model_set(config, dims, in, 0, "lbx", x0);
    ocp_nlp_constraints_model_set(config, dims, in, 0, "ubx", x0);
    ocp_nlp_constraints_model_set(config, dims, in, 0, "idxbxe", idxbx0);
#if defined(CFREF_MODEL_PENDULUM)
    double lbu[NU] = {-80}, ubu[NU] = {80};
#else
    double lbu[NU] = {0, 0, 0, 0}, ubu[NU] = {22, 22, 22, 22};
#endif
    for (int i = 0; i < N; i++) {
        ocp_nlp_constraints_model_set(config, dims, in, i, "idxbu", idxbu);
        ocp_nlp_constraints_model_set(config, dims, in, i, "lbu", lbu);
        ocp_nlp_constraints_model_set(config, dims, in, i, "ubu", ubu);
    }

    void *opts = h->opts = ocp_nlp_solver_opts_create(config, dims);
    int num_stages = 4, num_steps = 1, iter_max = 50, one = 1, zero = 0, rti_level = 4;
    double step_length = 1.0, lm = 0.0;
    ocp_nlp_solver_opts_set(config, opts, "globalization", "fixed_step");
    ocp_nlp_solver_opts_set(config, opts, "full_step_dual", &zero);
    for (int i = 0; i < N; i++) {
        ocp_nlp_solver_opts_set_at_stage(config, opts, i, "dynamics_num_steps", &num_steps);
        ocp_nlp_solver_opts_set_at_stage(config, opts, i, "dynamics_num_stages", &num_stages);
    }
    ocp_nlp_solver_opts_set(config, opts, "step_length", &step_length);
    ocp_nlp_solver_opts_set(config, opts, "levenberg_marquardt", &lm);
    ocp_nlp_solver_opts_set(config, opts, "qp_cond_N", &cond_N);
    ocp_nlp_solver_opts_set(config, opts, "qp_hpipm_mode", "BALANCE");
    ocp_nlp_solver_opts_set(config, opts, "as_rti_iter", &one);
    ocp_nlp_solver_opts_set(config, opts, "as_rti_level", &rti_level);
    ocp_nlp_solver_opts_set(config, opts, "rti_log_residuals", &zero);
    ocp_nlp_solver_opts_set(config, opts, "qp_iter_max", &iter_max);
    ocp_nlp_solver_opts_set(config, opts, "print_level", &zero);
    ocp_nlp_solver_opts_set(config, opts, "qp_cond_ric_alg", &one);
    ocp_nlp_solver_opts_set(config, opts, "qp_ric_alg", &one);

    ocp_nlp_out *out = h->out = ocp_nlp_out_create(config, dims);
    double u_init[NU] = {0};
    for (int i = 0; i <= N; i++) {
        ocp_nlp_out_set(config, dims, out, i, "x", x0);
        if (i < N) ocp_nlp_out_set(config, dims, out, i, "u", u_init);
    }
    h->solver = ocp_nlp_solver_create(config, dims, opts);
    if (ocp_nlp_precompute(h->solver, in, out) != 0) { cfref_destroy(h); return NULL; }
    return h;
}

/* Runtime weights (diagonals, cost order y=[x;u]) -- the node's SET_WEIGHTS path,
 * crazyflie_controller/src/acados_mpc.cpp:596-602. */
void cfref_set_weights(void *h_, const double *Wdiag, const double *WNdiag)
{
    cfref *h = h_;
    double W[NY * NY] = {0}, WN[NX * NX] = {0};
    for (int i = 0; i < NY; i++) W[i + NY * i] = Wdiag[i];
    for (int i = 0; i < NX; i++) WN[i + NX * i] = WNdiag[i];
    for (int i = 0; i < h->N; i++) ocp_nlp_cost_model_set(h->config, h->dims, h->in, i, "W", W);
    ocp_nlp_cost_model_set(h->config, h->dims, h->in, h->N, "W", WN);
}

void cfref_set_input_bounds(void *h_, const double *lbu, const double *ubu)
{
    cfref *h = h_;
    for (int i = 0; i < h->N; i++) {
        ocp_nlp_constraints_model_set(h->config, h->dims, h->in, i, "lbu", (void *) lbu);
        ocp_nlp_constraints_model_set(h->config, h->dims, h->in, i, "ubu", (void *) ubu);
    }
}

/* input box of ONE stage, as ocp_nlp_constraints_model_set addresses it (ocp_nlp_constraints_bgh.c:653-674) */
void cfref_set_input_bounds_at(void *h_, int stage, const double *lbu, const double *ubu)
{
    cfref *h = h_;
    ocp_nlp_constraints_model_set(h->config, h->dims, h->in, stage, "lbu", (void *) lbu);
    ocp_nlp_constraints_model_set(h->config, h->dims, h->in, stage, "ubu", (void *) ubu);
}

/* stage-0 input box only: what the node's FIXED_U0 branch does (acados_mpc.cpp:604-608) */
void cfref_set_input_bounds_stage0(void *h_, const double *lbu0, const double *ubu0)
{
    cfref *h = h_;
    ocp_nlp_constraints_model_set(h->config, h->dims, h->in, 0, "lbu", (void *) lbu0);
    ocp_nlp_constraints_model_set(h->config, h->dims, h->in, 0, "ubu", (void *) ubu0);
}

/* One RTI step exactly as NMPC::iteration does it (acados_mpc.cpp:581-625):
 * set x0 as lbx=ubx, set yref per stage, solve, read the iterate back.
 * x[(N+1)*13], u[N*4] are in/out (iterate before -> iterate after).
 * times[5] = time_tot, time_lin, time_qp_sol, time_qp_solver_call, time_qp_xcond. */
int cfref_rti(void *h_, const double *x0, const double *yref, const double *yref_e,
              double *x, double *u, int *qp_iter, int *qp_status, double *times)
{
    cfref *h = h_;
    int N = h->N;
    ocp_nlp_config *c = h->config;
    ocp_nlp_dims *d = h->dims;
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "lbx", (void *) x0);
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "ubx", (void *) x0);
    for (int k = 0; k < N; k++) ocp_nlp_cost_model_set(c, d, h->in, k, "yref", (void *) (yref + NY * k));
    ocp_nlp_cost_model_set(c, d, h->in, N, "yref", (void *) yref_e);
    for (int k = 0; k <= N; k++) {
        ocp_nlp_out_set(c, d, h->out, k, "x", x + NX * k);
        if (k < N) ocp_nlp_out_set(c, d, h->out, k, "u", u + NU * k);
    }
    int status = ocp_nlp_solve(h->solver, h->in, h->out);
    for (int k = 0; k <= N; k++) {
        ocp_nlp_out_get(c, d, h->out, k, "x", x + NX * k);
        if (k < N) ocp_nlp_out_get(c, d, h->out, k, "u", u + NU * k);
    }
    if (qp_iter) ocp_nlp_get(c, h->solver, "qp_iter", qp_iter);
    if (qp_status) ocp_nlp_get(c, h->solver, "qp_status", qp_status);
    if (times) {
        ocp_nlp_get(c, h->solver, "time_tot", times + 0);
        ocp_nlp_get(c, h->solver, "time_lin", times + 1);
        ocp_nlp_get(c, h->solver, "time_qp_sol", times + 2);
        ocp_nlp_get(c, h->solver, "time_qp_solver_call", times + 3);
        ocp_nlp_get(c, h->solver, "time_qp_xcond", times + 4);
    }
    return status;
}

/* The real-time iteration as the reference's two phases: rti_phase 1 (PREPARATION) with x0_prep in place, then the new
 * measurement x0_fb and rti_phase 2 (FEEDBACK); ocp_nlp_sqp_rti.c:189-198,1213-1237.  Leaves rti_phase at 0. */
int cfref_rti_split(void *h_, const double *x0_prep, const double *x0_fb, const double *yref, const double *yref_e,
                    double *x, double *u, int *qp_iter, int *qp_status)
{
    cfref *h = h_;
    int N = h->N, phase;
    ocp_nlp_config *c = h->config;
    ocp_nlp_dims *d = h->dims;
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "lbx", (void *) x0_prep);
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "ubx", (void *) x0_prep);
    for (int k = 0; k < N; k++) ocp_nlp_cost_model_set(c, d, h->in, k, "yref", (void *) (yref + NY * k));
    ocp_nlp_cost_model_set(c, d, h->in, N, "yref", (void *) yref_e);
    for (int k = 0; k <= N; k++) {
        ocp_nlp_out_set(c, d, h->out, k, "x", x + NX * k);
        if (k < N) ocp_nlp_out_set(c, d, h->out, k, "u", u + NU * k);
    }
    phase = 1;
    ocp_nlp_solver_opts_set(c, h->opts, "rti_phase", &phase);
    ocp_nlp_solve(h->solver, h->in, h->out);
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "lbx", (void *) x0_fb);
    ocp_nlp_constraints_model_set(c, d, h->in, 0, "ubx", (void *) x0_fb);
    phase = 2;
    ocp_nlp_solver_opts_set(c, h->opts, "rti_phase", &phase);
    int status = ocp_nlp_solve(h->solver, h->in, h->out);
    phase = 0;
    ocp_nlp_solver_opts_set(c, h->opts, "rti_phase", &phase);
    for (int k = 0; k <= N; k++) {
        ocp_nlp_out_get(c, d, h->out, k, "x", x + NX * k);
        if (k < N) ocp_nlp_out_get(c, d, h->out, k, "u", u + NU * k);
    }
    if (qp_iter) ocp_nlp_get(c, h->solver, "qp_iter", qp_iter);
    if (qp_status) ocp_nlp_get(c, h->solver, "qp_status", qp_status);
    return status;
}

/* QP data of the last solve, unpacked to plain row-major arrays:
 *  BAbt[N][17][13] rows = [u;x] of stage k, cols = x_{k+1};  b[N][13];
 *  rqz[(N)*17+13]; d_lb/d_ub: stage0 17 each ([u;x] order), stages 1..N-1 4 each;
 *  dux[(N)*17+13], dpi[N*13] = QP solution (step).  Any pointer may be NULL. */
void cfref_get_qp(void *h_, double *BAbt, double *b, double *rqz, double *d_lb, double *d_ub,
                  double *dux, double *dpi)
{
    cfref *h = h_;
    int N = h->N;
    ocp_nlp_memory *m;
    ocp_nlp_get(h->config, h->solver, "nlp_mem", &m);
    struct d_ocp_qp *qp = m->qp_in;
    struct d_ocp_qp_sol *sol = m->qp_out;
    double tmp[NV * NX];
    int od = 0;
    for (int k = 0; k <= N; k++) {
        int nv = k < N ? NV : NX;
        if (k < N && BAbt) {
            blasfeo_unpack_dmat(NV, NX, qp->BAbt + k, 0, 0, tmp, NV); /* col-major */
            for (int r = 0; r < NV; r++)
                for (int cc = 0; cc < NX; cc++) BAbt[(k * NV + r) * NX + cc] = tmp[r + NV * cc];
        }
        if (k < N && b) blasfeo_unpack_dvec(NX, qp->b + k, 0, b + NX * k, 1);
        if (rqz) blasfeo_unpack_dvec(nv, qp->rqz + k, 0, rqz + NV * k, 1);
        if (dux) blasfeo_unpack_dvec(nv, sol->ux + k, 0, dux + NV * k, 1);
        if (k < N && dpi) blasfeo_unpack_dvec(NX, sol->pi + k, 0, dpi + NX * k, 1);
        int nb = k == 0 ? NV : (k < N ? NU : 0);
        if (d_lb) blasfeo_unpack_dvec(nb, qp->d + k, 0, d_lb + od, 1);
        if (d_ub) blasfeo_unpack_dvec(nb, qp->d + k, nb, d_ub + od, 1);
        od += nb;
    }
}

/* HPIPM per-iteration statistics table of the last solve
 * (external/hpipm/ocp_qp/x_ocp_qp_ipm.c:2181-2200,2706-2714); returns rows. */
int cfref_get_ipm_stat(void *h_, double *stat, int max_rows, int *stat_m)
{
    cfref *h = h_;
    ocp_nlp_memory *m;
    ocp_nlp_get(h->config, h->solver, "nlp_mem", &m);
    ocp_qp_xcond_solver_memory *xm = (ocp_qp_xcond_solver_memory *) m->qp_solver_mem;
    ocp_qp_hpipm_memory *hm = (ocp_qp_hpipm_memory *) xm->solver_memory;
    struct d_ocp_qp_ipm_ws *w = hm->hpipm_workspace;
    int rows = w->iter + 1;
    if (rows > w->stat_max) rows = w->stat_max;
    if (rows > max_rows) rows = max_rows;
    if (stat_m) *stat_m = w->stat_m;
    if (stat) memcpy(stat, w->stat, sizeof(double) * rows * w->stat_m);
    return rows;
}

/* ---- batch driver: one solver object per worker thread, static slices -----
 * (the reference has no batch API; solver objects are independent).
 * Layouts: x0[n][13], yref[n][N*17], yref_e[n][13], x[n][(N+1)*13] in/out,
 * u[n][N*4] in/out, status[n], qp_iter[n], time_tot[n] (acados' own timer). */
typedef struct
{
    int N, cond_N, lo, hi, n_rti;
    double Ts;
    const double *x0, *yref, *yref_e;
    double *x, *u, *time_tot;
    int *status, *qp_iter;
    int fail;
} slice;

static void *slice_run(void *arg)
{
    slice *s = arg;
    void *h = cfref_create(s->N, s->Ts, s->cond_N);
    if (!h) { s->fail = 1; return NULL; }
    int N = s->N;
    for (int i = s->lo; i < s->hi; i++) {
        double t[5], tt = 0;
        int st = 0, it = 0;
        for (int r = 0; r < s->n_rti; r++) {
            st = cfref_rti(h, s->x0 + NX * (size_t) i, s->yref + (size_t) N * NY * i, s->yref_e + NX * (size_t) i,
                           s->x + (size_t) (N + 1) * NX * i, s->u + (size_t) N * NU * i, &it, NULL, t);
            tt += t[0];
        }
        if (s->status) s->status[i] = st;
        if (s->qp_iter) s->qp_iter[i] = it;
        if (s->time_tot) s->time_tot[i] = tt;
    }
    cfref_destroy(h);
    return NULL;
}

int cfref_batch(int N, double Ts, int cond_N, int nthreads, int n_rti, int n,
                const double *x0, const double *yref, const double *yref_e,
                double *x, double *u, int *status, int *qp_iter, double *time_tot)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n > 0 ? n : 1;
    pthread_t *th = calloc(nthreads, sizeof(pthread_t));
    slice *sl = calloc(nthreads, sizeof(slice));
    int fail = 0;
    for (int t = 0; t < nthreads; t++) {
        sl[t] = (slice){N, cond_N, (int) ((long) n * t / nthreads), (int) ((long) n * (t + 1) / nthreads), n_rti,
                        Ts, x0, yref, yref_e, x, u, time_tot, status, qp_iter, 0};
        if (nthreads == 1) slice_run(&sl[t]);
        else pthread_create(&th[t], NULL, slice_run, &sl[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        if (nthreads > 1) pthread_join(th[t], NULL);
        fail |= sl[t].fail;
    }
    free(th); free(sl);
    return fail;
}

/* ------------------------------------------------------------------ state predictor
 * The reference's own ERK integrator driven the way acados_sim_solver_crazyflie.c does it
 * (c_templates_tera/acados_sim_solver.in.c:236-396): plan ERK, dims nx/nu, 4 stages, num_steps, sens_forw,
 * S_forw seeded with [I 0], then per call sim_in_set "T","x","u" -> sim_solve -> sim_out_get "xn"
 * (crazyflie_controller/src/acados_estimator.cpp:573-593). */
#include "acados_c/sim_interface.h"

typedef struct
{
    sim_config *config;
    void *dims, *opts;
    sim_in *in;
    sim_out *out;
    sim_solver *solver;
    external_function_generic vde, ode;
    int sens_forw;
} cfsim;

void cfref_sim_destroy(void *h_)
{
    cfsim *h = h_;
    if (!h) return;
    if (h->solver) sim_solver_destroy(h->solver);
    if (h->in) sim_in_destroy(h->in);
    if (h->out) sim_out_destroy(h->out);
    if (h->opts) sim_opts_destroy(h->opts);
    if (h->dims) sim_dims_destroy(h->dims);
    if (h->config) sim_config_destroy(h->config);
    free(h);
}

void *cfref_sim_create(int num_steps, int sens_forw)
{
    cfsim *h = calloc(1, sizeof *h);
    sim_solver_plan_t plan;
    plan.sim_solver = ERK;
    h->config = sim_config_create(plan);
    h->dims = sim_dims_create(h->config);
    int nx = NX, nu = NU, nz = 0;
    sim_dims_set(h->config, h->dims, "nx", &nx);
    sim_dims_set(h->config, h->dims, "nu", &nu);
    sim_dims_set(h->config, h->dims, "nz", &nz);
    h->opts = sim_opts_create(h->config, h->dims);
    int ns = 4;
    bool sf = sens_forw != 0, no = false;
    sim_opts_set(h->config, h->opts, "num_stages", &ns);
    sim_opts_set(h->config, h->opts, "num_steps", &num_steps);
    sim_opts_set(h->config, h->opts, "sens_forw", &sf);
    sim_opts_set(h->config, h->opts, "sens_adj", &no);
    sim_opts_set(h->config, h->opts, "sens_hess", &no);
    h->sens_forw = sens_forw;
    h->in = sim_in_create(h->config, h->dims);
    h->out = sim_out_create(h->config, h->dims);
    h->vde.evaluate = &vde_eval;
    h->ode.evaluate = &ode_eval;
    h->config->model_set(h->in->model, "expl_vde_forw", &h->vde);
    h->config->model_set(h->in->model, "expl_ode_fun", &h->ode);
    h->solver = sim_solver_create(h->config, h->dims, h->opts);
    double S[NX * (NX + NU)] = {0};
    for (int i = 0; i < NX; i++) S[i + NX * i] = 1.0;
    sim_in_set(h->config, h->dims, h->in, "S_forw", S);
    double T = 0.015;
    sim_in_set(h->config, h->dims, h->in, "T", &T);
    if (sim_precompute(h->solver, h->in, h->out)) { cfref_sim_destroy(h); return NULL; }
    return h;
}

/* xn[13]; S_forw[13*17] column-major, columns [x(13) | u(4)] (may be NULL) */
int cfref_sim_solve(void *h_, const double *x, const double *u, double T, double *xn, double *S_forw)
{
    cfsim *h = h_;
    sim_in_set(h->config, h->dims, h->in, "T", &T);
    sim_in_set(h->config, h->dims, h->in, "x", (void *) x);
    sim_in_set(h->config, h->dims, h->in, "u", (void *) u);
    int status = sim_solve(h->solver, h->in, h->out);
    sim_out_get(h->config, h->dims, h->out, "xn", xn);
    if (S_forw && h->sens_forw) sim_out_get(h->config, h->dims, h->out, "S_forw", S_forw);
    return status;
}

/* ------------------------------------------------------------------ round-2 additions (edge-case parity)
 * Solver option by name, integer valued ("qp_iter_max", "qp_cond_N", ...): ocp_nlp_solver_opts_set,
 * acados/interfaces/acados_c/ocp_nlp_interface.h:313. */
void cfref_set_opt_int(void *h_, const char *name, int value)
{
    cfref *h = h_;
    ocp_nlp_solver_opts_set(h->config, h->opts, name, &value);
}

/* Multipliers of the iterate after the last solve, as ocp_nlp_out_get hands them out
 * (acados_c/ocp_nlp_interface.c:549-600): pi[N][13]; lam: stage 0 2*17, stages 1..N-1 2*4 each ([lower | upper]). */
void cfref_get_multipliers(void *h_, double *pi, double *lam)
{
    cfref *h = h_;
    int N = h->N, ol = 0;
    for (int k = 0; k < N; k++) {
        if (pi) ocp_nlp_out_get(h->config, h->dims, h->out, k, "pi", pi + NX * k);
        int nb = k == 0 ? NX + NU : NU;
        if (lam) ocp_nlp_out_get(h->config, h->dims, h->out, k, "lam", lam + ol);
        ol += 2 * nb;
    }
}

/* Cost weight of ONE stage as a full column-major matrix (17x17 for k < N, 13x13 for k = N):
 * ocp_nlp_cost_model_set(.., k, "W", ..), ocp_nlp_cost_ls.c:301-331. */
void cfref_set_W_at(void *h_, int stage, const double *W)
{
    cfref *h = h_;
    ocp_nlp_cost_model_set(h->config, h->dims, h->in, stage, "W", (void *) W);
}

/* Solver statistics by name (double or int valued): ocp_nlp_get, ocp_nlp_sqp_rti.c:1361-1425 */
void cfref_get_stat(void *h_, const char *name, void *value)
{
    cfref *h = h_;
    ocp_nlp_get(h->config, h->solver, name, value);
}

/* The acados objects behind a handle, for the node harness (tests/dropin/refglue/ref_node_glue.c), which must fill the
 * process globals the reference's ROS node defines (crazyflie_controller/src/acados_mpc.cpp:76-84). */
void cfref_export(void *h_, void **plan, void **config, void **dims, void **in, void **out, void **opts, void **solver)
{
    cfref *h = h_;
    *plan = h->plan; *config = h->config; *dims = h->dims; *in = h->in; *out = h->out; *opts = h->opts; *solver = h->solver;
}

/* sizes this build of the harness was compiled for */
void cfref_dims(int *nx, int *nu) { *nx = NX; *nu = NU; }
