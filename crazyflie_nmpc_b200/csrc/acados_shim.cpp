// Single-instance drop-in surfaces (include/acados_solver_crazyflie.h, include/acados_c/ocp_nlp_interface.h):
// the entry points `crazyflie_controller`'s NMPC node binds, implemented as a batch of ONE instance over the
// batch C-ABI (include/cfnmpc.h).  Host logic only; every solve is a GPU kernel launch.
//
// Reference behaviour mirrored here:
//   setters copy caller arrays, getters copy out        ocp_nlp_cost_ls.c:357-361, ocp_nlp_interface.c:549-557
//   the iterate persists between solves                  SURVEY.md fact 5
//   initial iterate x_k = lbx_0 (= [0,0,0,1,0..]), u = 0 acados_solver.in.c:2323-2352
//   status codes                                         acados/utils/types.h:75-83
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/acados_solver_crazyflie.h"
#include "../../include/cfnmpc.h"
#define CF_DEV static inline   // host-only use of the generated OCP description (defaults below)
#include "cf_spec_generated.h"
static_assert(CF_SPEC_N == CRAZYFLIE_N, "include/acados_solver_crazyflie.h and the generated spec disagree on N");

struct crazyflie_solver_capsule
{
    cfnmpc_batch *batch = nullptr;
    int N = 0;
    double Ts = 0.015;
    std::vector<double> x0, yref, yref_e, x, u;
    bool in_dirty = true, iterate_dirty = true;
    int status = 0, qp_iter = 0, qp_status = 0, cond_N = 0;
    int rti_phase = 0;   // 0 preparation + feedback, 1 preparation, 2 feedback (ocp_nlp_sqp_rti.c:189-198)
    std::vector<double> bnd;   // input box per stage, [N][8] = lbu | ubu (ocp_nlp_constraints_bgh.c:653-674)
    std::vector<double> wst;   // weight diagonals per stage, [N+1][17] (ocp_nlp_cost_ls.c:301-331); uploaded once stages differ
    bool wst_used = false;
    std::vector<double> wdense;      // full weight matrices per stage, [N+1][17][17] row-major (filled on the first "W" set)
    std::vector<char> wdense_nd;     // stage carries an off-diagonal entry
    bool wdense_used = false;
    // multipliers of the iterate as ocp_nlp_out carries them (zero until the first feedback; acados_c/ocp_nlp_interface.c:576-590)
    std::vector<double> pi, lam, t, lam_x0;   // [N][13], [N][8], [N][8], [13] (signed, see include/cfnmpc.h)
    bool mult_on = false, mult_valid = false;
    double time_tot = 0.0, time_lin = 0.0, time_qp_sol = 0.0;
    double stat[4] = {0, 0, 0, 0};   // SQP_RTI statistics table: stat_m = 2 rows x stat_n = 2 (qp_status, qp_iter), ocp_nlp_sqp_rti.c:264-265,648-649
    ocp_nlp_plan_t plan;
    ocp_nlp_config config;
    ocp_nlp_dims dims;
    ocp_nlp_in in;
    ocp_nlp_out out;
    ocp_nlp_solver solver;
    int opts_dummy = 0;
};

static void reset_iterate(crazyflie_solver_capsule *c)
{
    const int N = c->N;
    c->x.assign((size_t) (N + 1) * 13, 0.0);
    for (int k = 0; k <= N; k++) c->x[(size_t) k * 13 + 3] = 1.0;
    c->u.assign((size_t) N * 4, 0.0);
    c->iterate_dirty = true;
}

extern "C" {

crazyflie_solver_capsule *crazyflie_acados_create_capsule(void) { return new crazyflie_solver_capsule(); }

int crazyflie_acados_free_capsule(crazyflie_solver_capsule *c)
{
    delete c;
    return 0;
}

int crazyflie_acados_free(crazyflie_solver_capsule *c)
{
    if (!c) return 1;
    if (c->batch) cfnmpc_batch_destroy(c->batch);
    c->batch = nullptr;
    return 0;
}

int crazyflie_acados_create_with_discretization(crazyflie_solver_capsule *c, int N, double *new_time_steps)
{
    if (!c || N < 1) return 1;
    double Ts = CF_SPEC_TF / CF_SPEC_N;  // Tf / N of generate_c_code.py:41-42
    if (new_time_steps) {
        Ts = new_time_steps[0];
        for (int i = 0; i < N; i++)
            if (!(new_time_steps[i] > 0.0)) {
                fprintf(stderr, "crazyflie_acados_create_with_discretization: time steps must be positive\n");
                return 1;
            }
    } else if (N != CRAZYFLIE_N) {
        fprintf(stderr, "crazyflie_acados_create_with_discretization: new_time_steps is required when N != %d\n", CRAZYFLIE_N);
        return 1;
    }
    crazyflie_acados_free(c);
    if (cfnmpc_batch_create(1, N, Ts, 0, &c->batch) != CFNMPC_OK) {
        fprintf(stderr, "crazyflie_acados_create: %s\n", cfnmpc_last_error());
        c->batch = nullptr;
        return 1;
    }
    // one step per shooting interval = its cost scaling (acados_solver.in.c:133-153,879-892)
    if (new_time_steps && cfnmpc_batch_set(c->batch, "time_steps", new_time_steps, 0) != CFNMPC_OK) {
        fprintf(stderr, "crazyflie_acados_create: %s\n", cfnmpc_last_error());
        crazyflie_acados_free(c);
        return 1;
    }
    c->N = N;
    c->Ts = Ts;
    c->rti_phase = 0;
    c->pi.assign((size_t) N * 13, 0.0); c->lam.assign((size_t) N * 8, 0.0); c->t.assign((size_t) N * 8, 0.0); c->lam_x0.assign(13, 0.0);
    c->mult_on = cfnmpc_batch_set_option(c->batch, "multipliers", 1) == CFNMPC_OK;
    c->mult_valid = false;
    c->wst.assign((size_t) (N + 1) * 17, 0.0);
    for (int k = 0; k <= N; k++)
        for (int i = 0; i < 17; i++) c->wst[(size_t) k * 17 + i] = k < N ? CfSpec::W[i] : (i < 13 ? CfSpec::W_e[i] : 0.0);
    c->wst_used = false;
    c->wdense.clear(); c->wdense_nd.clear(); c->wdense_used = false;
    c->bnd.assign((size_t) N * 8, 0.0);   // generate_c_code.py:133-134
    for (int k = 0; k < N; k++)
        for (int i = 0; i < 4; i++) { c->bnd[(size_t) k * 8 + i] = CfSpec::lbu[i]; c->bnd[(size_t) k * 8 + 4 + i] = CfSpec::ubu[i]; }
    // generate_c_code.py:128-129 reference, :135 x0 (through tools/gen_spec.py)
    const double *y = CfSpec::yref;
    c->x0.assign(CfSpec::x0, CfSpec::x0 + 13);
    c->yref.resize((size_t) N * 17);
    for (int k = 0; k < N; k++) memcpy(&c->yref[(size_t) k * 17], y, 17 * sizeof(double));
    c->yref_e.assign(CfSpec::yref_e, CfSpec::yref_e + 13);
    reset_iterate(c);
    c->in_dirty = true;
    c->plan.N = N;
    c->config.N = N; c->config.capsule = c;
    c->dims.N = N; c->dims.nx = 13; c->dims.nu = 4; c->dims.ny = 17; c->dims.ny_e = 13; c->dims.capsule = c;
    c->in.capsule = c;
    c->out.capsule = c; c->out.inf_norm_res = 0.0; c->out.total_time = 0.0;
    c->solver.capsule = c;
    c->cond_N = N;
    return 0;
}

int crazyflie_acados_create(crazyflie_solver_capsule *c) { return crazyflie_acados_create_with_discretization(c, CRAZYFLIE_N, nullptr); }

int crazyflie_acados_update_time_steps(crazyflie_solver_capsule *c, int N, double *new_time_steps)
{
    if (!c || !c->batch || N != c->N || !new_time_steps) return 1;
    // acados_solver.in.c:133-153: "Ts" and the cost "scaling" of every interval; inputs, weights and iterate stay
    if (cfnmpc_batch_set(c->batch, "time_steps", new_time_steps, 0) != CFNMPC_OK) {
        fprintf(stderr, "crazyflie_acados_update_time_steps: %s\n", cfnmpc_last_error());
        return 1;
    }
    c->Ts = new_time_steps[0];
    return 0;
}

int crazyflie_acados_update_qp_solver_cond_N(crazyflie_solver_capsule *c, int cond_N)
{
    if (!c || cond_N < 1) return 1;
    c->cond_N = cond_N;
    // blocks of up to 3 stages run on the partially condensed kernel (cf_pcond_warp.h); for coarser condensing the QP is
    // solved uncondensed: same solution to the interior-point tolerances, stated on stderr once
    // (the condensed feedback does not recover the multipliers of the eliminated stages: ocp_nlp_out_get "pi"/"lam"/"t"
    // then keep the values of the last uncondensed step)
    if (c->batch) {
        if (cond_N < c->N) { cfnmpc_batch_set_option(c->batch, "multipliers", 0); c->mult_on = false; }
    }
    if (c->batch && cfnmpc_batch_set_option(c->batch, "qp_cond_N", cond_N) != CFNMPC_OK) {
        fprintf(stderr, "crazyflie_acados_update_qp_solver_cond_N: %s; solving uncondensed\n", cfnmpc_last_error());
        cfnmpc_batch_set_option(c->batch, "qp_cond_N", 0);
        cond_N = c->N;
    }
    if (c->batch && cond_N >= c->N && !c->mult_on) c->mult_on = cfnmpc_batch_set_option(c->batch, "multipliers", 1) == CFNMPC_OK;
    return 0;
}

int crazyflie_acados_reset(crazyflie_solver_capsule *c, int)
{
    if (!c || !c->batch) return 1;
    // acados_solver.in.c:2467-2520 zeroes the iterate
    c->x.assign(c->x.size(), 0.0);
    c->u.assign(c->u.size(), 0.0);
    c->iterate_dirty = true;
    return 0;
}

int crazyflie_acados_solve(crazyflie_solver_capsule *c)
{
    if (!c || !c->batch) return ACADOS_QP_FAILURE;
    cfnmpc_batch *b = c->batch;
    int rc = 0;
    if (c->in_dirty) {
        rc |= cfnmpc_batch_set(b, "x0", c->x0.data(), 0);
        rc |= cfnmpc_batch_set(b, "yref", c->yref.data(), 0);
        rc |= cfnmpc_batch_set(b, "yref_e", c->yref_e.data(), 0);
        c->in_dirty = false;
    }
    if (c->iterate_dirty) {
        rc |= cfnmpc_batch_set(b, "x", c->x.data(), 0);
        rc |= cfnmpc_batch_set(b, "u", c->u.data(), 0);
        c->iterate_dirty = false;
    }
    // ocp_nlp_sqp_rti.c:1213-1237
    if (c->rti_phase == 1) {
        rc |= cfnmpc_batch_prepare(b);
        rc |= cfnmpc_batch_sync(b);
        double ms = 0.0;
        rc |= cfnmpc_batch_last_solve_ms(b, &ms);
        if (rc) {
            fprintf(stderr, "crazyflie_acados_solve (preparation): %s\n", cfnmpc_last_error());
            return ACADOS_QP_FAILURE;
        }
        c->time_tot = ms * 1e-3;
        c->out.total_time = c->time_tot;
        return c->status;   // mem->status is left as the last feedback set it
    }
    rc |= (c->rti_phase == 2) ? cfnmpc_batch_feedback(b) : cfnmpc_batch_solve(b, 1);
    rc |= cfnmpc_batch_get(b, "x_all", 0, c->x.data(), 0);
    rc |= cfnmpc_batch_get(b, "u_all", 0, c->u.data(), 0);
    rc |= cfnmpc_batch_get(b, "status", 0, &c->status, 0);
    rc |= cfnmpc_batch_get(b, "qp_iter", 0, &c->qp_iter, 0);
    rc |= cfnmpc_batch_get(b, "qp_status", 0, &c->qp_status, 0);
    if (c->mult_on && !rc && c->status == ACADOS_SUCCESS) {   // a failed QP leaves the duals untouched, like the primal iterate
        rc |= cfnmpc_batch_get(b, "pi_all", 0, c->pi.data(), 0);
        rc |= cfnmpc_batch_get(b, "lam_all", 0, c->lam.data(), 0);
        rc |= cfnmpc_batch_get(b, "t_all", 0, c->t.data(), 0);
        rc |= cfnmpc_batch_get(b, "lam_x0", 0, c->lam_x0.data(), 0);
        c->mult_valid = true;
    }
    double ms = 0.0, ph[2] = {0.0, 0.0};
    rc |= cfnmpc_batch_last_solve_ms(b, &ms);
    rc |= cfnmpc_batch_last_phase_ms(b, ph);
    if (rc) {
        fprintf(stderr, "crazyflie_acados_solve: %s\n", cfnmpc_last_error());
        return ACADOS_QP_FAILURE;
    }
    c->time_tot = ms * 1e-3;
    // device time of the two launches of the step: linearisation (preparation kernel) and QP solution (feedback kernel);
    // a fused or feedback-only step reports everything as time_qp_sol (mem->time_lin / time_qp_sol, ocp_nlp_sqp_rti.c:1361-1384)
    c->time_lin = ph[0] * 1e-3;
    c->time_qp_sol = ph[1] * 1e-3;
    c->stat[0] = c->qp_status; c->stat[1] = c->qp_iter;
    c->out.total_time = c->time_tot;
    return c->status;
}

void crazyflie_acados_print_stats(crazyflie_solver_capsule *c)
{
    if (!c) return;
    printf("iter\tqp_stat\tqp_iter\n%d\t%d\t%d\n", 1, c->qp_status, c->qp_iter);
}

ocp_nlp_in *crazyflie_acados_get_nlp_in(crazyflie_solver_capsule *c) { return &c->in; }
ocp_nlp_out *crazyflie_acados_get_nlp_out(crazyflie_solver_capsule *c) { return &c->out; }
ocp_nlp_solver *crazyflie_acados_get_nlp_solver(crazyflie_solver_capsule *c) { return &c->solver; }
ocp_nlp_config *crazyflie_acados_get_nlp_config(crazyflie_solver_capsule *c) { return &c->config; }
void *crazyflie_acados_get_nlp_opts(crazyflie_solver_capsule *c) { return &c->opts_dummy; }
ocp_nlp_dims *crazyflie_acados_get_nlp_dims(crazyflie_solver_capsule *c) { return &c->dims; }
ocp_nlp_plan_t *crazyflie_acados_get_nlp_plan(crazyflie_solver_capsule *c) { return &c->plan; }

// ------------------------------------------------------------------ acados_c subset
int ocp_nlp_constraints_model_set(ocp_nlp_config *, ocp_nlp_dims *, ocp_nlp_in *in, int stage, const char *field, void *value)
{
    if (!in || !in->capsule || !field || !value) return 1;
    crazyflie_solver_capsule *c = in->capsule;
    if (!strcmp(field, "lbx") || !strcmp(field, "ubx")) {
        if (stage != 0) return 1;
        memcpy(c->x0.data(), value, 13 * sizeof(double));  // lbx_0 = ubx_0 = measured state (acados_mpc.cpp:581-582)
        c->in_dirty = true;
        return 0;
    }
    if (!strcmp(field, "lbu") || !strcmp(field, "ubu")) {
        // one stage at a time, as in the reference (ocp_nlp_constraints_bgh.c:653-674; the node's FIXED_U0 branch pins
        // stage 0 this way, acados_mpc.cpp:604-608)
        if (stage < 0 || stage >= c->N || !c->batch) return 1;
        memcpy(&c->bnd[(size_t) stage * 8 + (field[0] == 'l' ? 0 : 4)], value, 4 * sizeof(double));
        return cfnmpc_batch_set(c->batch, "bounds_stage", c->bnd.data(), 0) != CFNMPC_OK;
    }
    return 1;
}

int ocp_nlp_cost_model_set(ocp_nlp_config *, ocp_nlp_dims *, ocp_nlp_in *in, int stage, const char *field, void *value)
{
    if (!in || !in->capsule || !field || !value) return 1;
    crazyflie_solver_capsule *c = in->capsule;
    if (stage < 0 || stage > c->N) return 1;
    if (!strcmp(field, "yref") || !strcmp(field, "y_ref")) {
        if (stage < c->N) memcpy(&c->yref[(size_t) stage * 17], value, 17 * sizeof(double));
        else memcpy(c->yref_e.data(), value, 13 * sizeof(double));
        c->in_dirty = true;
        return 0;
    }
    if (!strcmp(field, "W")) {
        // column-major ny x ny as the node fills it (acados_mpc.cpp:526-542)
        const int n = stage < c->N ? 17 : 13;
        const double *W = static_cast<const double *>(value);
        double d[17];
        bool diag = true;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) {
                if (i == j) d[i] = W[i + n * j];
                else if (W[i + n * j] != 0.0) diag = false;
            }
        if (!c->batch) return 1;
        // Any SPD matrix is accepted, as in the reference (ocp_nlp_cost_ls.c:301-331).  The full matrices of all stages are
        // kept; once one of them has an off-diagonal entry the table goes to the dense-Hessian kernel ("W_dense_table"),
        // and back to the diagonal paths when every stage is diagonal again.
        if (c->wdense.empty()) {
            c->wdense.assign((size_t) (c->N + 1) * 289, 0.0);
            for (int k = 0; k <= c->N; k++)
                for (int i = 0; i < (k < c->N ? 17 : 13); i++) c->wdense[(size_t) k * 289 + i * 18] = c->wst[(size_t) k * 17 + i];
            c->wdense_nd.assign((size_t) c->N + 1, 0);
        }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) c->wdense[(size_t) stage * 289 + i * 17 + j] = 0.5 * (W[i + n * j] + W[j + n * i]);
        c->wdense_nd[stage] = diag ? 0 : 1;
        bool any_nd = false;
        for (char f : c->wdense_nd) any_nd = any_nd || f;
        if (any_nd) {
            if (c->mult_on) { cfnmpc_batch_set_option(c->batch, "multipliers", 0); c->mult_on = false; }
            if (cfnmpc_batch_set(c->batch, "W_dense_table", c->wdense.data(), 0) != CFNMPC_OK) {
                fprintf(stderr, "ocp_nlp_cost_model_set W: %s\n", cfnmpc_last_error());
                return 1;
            }
            c->wdense_used = true;
            memcpy(&c->wst[(size_t) stage * 17], d, n * sizeof(double));
            return 0;
        }
        if (c->wdense_used) {
            cfnmpc_batch_clear(c->batch, "W_dense_table");
            c->wdense_used = false;
            if (c->cond_N >= c->N && !c->mult_on) c->mult_on = cfnmpc_batch_set_option(c->batch, "multipliers", 1) == CFNMPC_OK;
        }
        // the reference sets the weight of ONE stage (ocp_nlp_cost_ls.c:301-331).  While all stages < N carry the same
        // diagonal (the node's SET_WEIGHTS loop, acados_mpc.cpp:596-602) the solver-wide value and the fast kernels are
        // used; as soon as stages differ the per-stage table takes over.
        memcpy(&c->wst[(size_t) stage * 17], d, n * sizeof(double));
        bool same = true;
        for (int k = 1; k < c->N && same; k++) same = !memcmp(&c->wst[0], &c->wst[(size_t) k * 17], 17 * sizeof(double));
        if (same) {
            if (c->wst_used) { cfnmpc_batch_clear(c->batch, "W_stage"); c->wst_used = false; }
            int rc = cfnmpc_batch_set(c->batch, "W", &c->wst[0], 0);
            rc |= cfnmpc_batch_set(c->batch, "W_e", &c->wst[(size_t) c->N * 17], 0);
            return rc != CFNMPC_OK;
        }
        c->wst_used = true;
        return cfnmpc_batch_set(c->batch, "W_stage", c->wst.data(), 0) != CFNMPC_OK;
    }
    return 1;
}

void ocp_nlp_out_set(ocp_nlp_config *, ocp_nlp_dims *, ocp_nlp_out *out, int stage, const char *field, void *value)
{
    if (!out || !out->capsule || !field || !value) return;
    crazyflie_solver_capsule *c = out->capsule;
    if (!strcmp(field, "x") && stage >= 0 && stage <= c->N) memcpy(&c->x[(size_t) stage * 13], value, 13 * sizeof(double));
    else if (!strcmp(field, "u") && stage >= 0 && stage < c->N) memcpy(&c->u[(size_t) stage * 4], value, 4 * sizeof(double));
    else return;
    c->iterate_dirty = true;
}

void ocp_nlp_out_get(ocp_nlp_config *, ocp_nlp_dims *, ocp_nlp_out *out, int stage, const char *field, void *value)
{
    if (!out || !out->capsule || !field || !value) return;
    crazyflie_solver_capsule *c = out->capsule;
    if (!strcmp(field, "x") && stage >= 0 && stage <= c->N) memcpy(value, &c->x[(size_t) stage * 13], 13 * sizeof(double));
    else if (!strcmp(field, "u") && stage >= 0 && stage < c->N) memcpy(value, &c->u[(size_t) stage * 4], 4 * sizeof(double));
    else if (!strcmp(field, "pi") && stage >= 0 && stage < c->N) memcpy(value, &c->pi[(size_t) stage * 13], 13 * sizeof(double));
    else if ((!strcmp(field, "lam") || !strcmp(field, "t")) && stage >= 0 && stage < c->N) {
        // 2 * ni[stage] values, [lower | upper], inputs before states: stage 0 carries the 13 bounds lbx_0 = ubx_0 = x0 whose
        // multipliers the reference restores after the elimination (x_ocp_qp_red.c:796-840: slacks t_min, multipliers
        // lam_min except the side the stationarity condition selects)
        const bool is_lam = field[0] == 'l';
        const double *src = &(is_lam ? c->lam : c->t)[(size_t) stage * 8];
        double *v = static_cast<double *>(value);
        if (stage > 0) memcpy(v, src, 8 * sizeof(double));
        else {
            const double fill = c->mult_valid ? 1e-16 : 0.0;
            for (int i = 0; i < 4; i++) { v[i] = src[i]; v[17 + i] = src[4 + i]; }
            for (int i = 0; i < 13; i++) {
                v[4 + i] = fill; v[21 + i] = fill;
                if (is_lam && c->mult_valid) {
                    if (c->lam_x0[i] >= 0) v[4 + i] = c->lam_x0[i];
                    else v[21 + i] = -c->lam_x0[i];
                }
            }
        }
    }
}

void ocp_nlp_solver_opts_set(ocp_nlp_config *config, void *, const char *field, void *value)
{
    if (!config || !config->capsule || !field || !value) return;
    if (!strcmp(field, "qp_cond_N")) crazyflie_acados_update_qp_solver_cond_N(config->capsule, *static_cast<int *>(value));
    else if (!strcmp(field, "rti_phase")) {
        const int v = *static_cast<int *>(value);
        if (v < 0 || v > 2) fprintf(stderr, "ocp_nlp_solver_opts_set: invalid value %d for rti_phase (0, 1, 2)\n", v);
        else config->capsule->rti_phase = v;
    }
}

int ocp_nlp_solve(ocp_nlp_solver *solver, ocp_nlp_in *, ocp_nlp_out *)
{
    if (!solver || !solver->capsule) return ACADOS_QP_FAILURE;
    return crazyflie_acados_solve(solver->capsule);
}

void ocp_nlp_get(ocp_nlp_config *, ocp_nlp_solver *solver, const char *field, void *ret)
{
    if (!solver || !solver->capsule || !field || !ret) return;
    crazyflie_solver_capsule *c = solver->capsule;
    if (!strcmp(field, "time_tot") || !strcmp(field, "tot_time")) *static_cast<double *>(ret) = c->time_tot;
    else if (!strcmp(field, "qp_iter")) *static_cast<int *>(ret) = c->qp_iter;
    else if (!strcmp(field, "sqp_iter")) *static_cast<int *>(ret) = 1;
    else if (!strcmp(field, "status")) *static_cast<int *>(ret) = c->status;
    else if (!strcmp(field, "qp_status")) *static_cast<int *>(ret) = c->qp_status;
    // ocp_nlp_sqp_rti.c:1361-1425
    else if (!strcmp(field, "time_lin")) *static_cast<double *>(ret) = c->time_lin;
    else if (!strcmp(field, "time_qp_sol") || !strcmp(field, "time_qp")) *static_cast<double *>(ret) = c->time_qp_sol;
    else if (!strcmp(field, "time_qp_solver") || !strcmp(field, "time_qp_solver_call")) *static_cast<double *>(ret) = c->time_qp_sol;
    else if (!strcmp(field, "time_qp_xcond") || !strcmp(field, "time_reg") || !strcmp(field, "time_glob") || !strcmp(field, "time_sim") ||
             !strcmp(field, "time_sim_ad") || !strcmp(field, "time_sim_la") || !strcmp(field, "time_solution_sensitivities"))
        *static_cast<double *>(ret) = 0.0;   // not separated from the two kernel times above
    else if (!strcmp(field, "stat")) *static_cast<double **>(ret) = c->stat;
    else if (!strcmp(field, "stat_m")) *static_cast<int *>(ret) = 2;
    else if (!strcmp(field, "stat_n")) *static_cast<int *>(ret) = 2;
    else if (!strcmp(field, "statistics")) {
        // n_row = min(stat_m, sqp_iter + 1) = 1 row, column-major [iter | qp_status | qp_iter]
        double *v = static_cast<double *>(ret);
        v[0] = 0; v[1] = c->stat[0]; v[2] = c->stat[1];
    }
}

// ------------------------------------------------------------------ legacy surface (process globals)
__attribute__((weak)) ocp_nlp_in *nlp_in;
__attribute__((weak)) ocp_nlp_out *nlp_out;
__attribute__((weak)) ocp_nlp_solver *nlp_solver;
__attribute__((weak)) void *nlp_opts;
__attribute__((weak)) ocp_nlp_plan *nlp_solver_plan;
__attribute__((weak)) ocp_nlp_config *nlp_config;
__attribute__((weak)) ocp_nlp_dims *nlp_dims;
static crazyflie_solver_capsule *g_capsule = nullptr;

int acados_create(void)
{
    if (g_capsule) acados_free();
    g_capsule = crazyflie_acados_create_capsule();
    if (crazyflie_acados_create(g_capsule)) {
        crazyflie_acados_free_capsule(g_capsule);
        g_capsule = nullptr;
        return 1;
    }
    nlp_in = &g_capsule->in; nlp_out = &g_capsule->out; nlp_solver = &g_capsule->solver;
    nlp_opts = &g_capsule->opts_dummy; nlp_solver_plan = &g_capsule->plan; nlp_config = &g_capsule->config;
    nlp_dims = &g_capsule->dims;
    return 0;
}

int acados_solve(void) { return g_capsule ? crazyflie_acados_solve(g_capsule) : ACADOS_QP_FAILURE; }

int acados_free(void)
{
    if (!g_capsule) return 0;
    crazyflie_acados_free(g_capsule);
    crazyflie_acados_free_capsule(g_capsule);
    g_capsule = nullptr;
    nlp_in = nullptr; nlp_out = nullptr; nlp_solver = nullptr; nlp_opts = nullptr; nlp_solver_plan = nullptr;
    nlp_config = nullptr; nlp_dims = nullptr;
    return 0;
}

}  // extern "C"

// ================================================================== state predictor surfaces
// include/acados_sim_solver_crazyflie.h + include/acados_c/sim_interface.h: one-instance wrappers over cfnmpc_sim.
#include "../../include/acados_sim_solver_crazyflie.h"

struct crazyflie_sim_solver_capsule
{
    cfnmpc_sim *sim = nullptr;
    double x[13] = {0}, u[4] = {0}, T = 0.015, xn[13] = {0}, S[13 * 17] = {0};
    bool sens_forw = true;   // the generated sim solver keeps forward sensitivities on (acados_sim_solver.in.c:280-281)
    bool have_S = false;
    sim_config config;
    sim_in in;
    sim_out out;
    sim_opts opts;
    sim_solver solver;
    int dims_dummy = 0;
};

extern "C" {

crazyflie_sim_solver_capsule *crazyflie_acados_sim_solver_create_capsule(void) { return new crazyflie_sim_solver_capsule(); }

int crazyflie_acados_sim_free_capsule_solver(crazyflie_sim_solver_capsule *c)
{
    if (!c) return 1;
    if (c->sim) cfnmpc_sim_destroy(c->sim);
    c->sim = nullptr;
    return 0;
}

int crazyflie_acados_sim_solver_free_capsule(crazyflie_sim_solver_capsule *c)
{
    if (c) crazyflie_acados_sim_free_capsule_solver(c);
    delete c;
    return 0;
}

int crazyflie_acados_sim_create_capsule_solver(crazyflie_sim_solver_capsule *c)
{
    if (!c) return 1;
    crazyflie_acados_sim_free_capsule_solver(c);
    if (cfnmpc_sim_create(1, 0, &c->sim) != CFNMPC_OK) {
        fprintf(stderr, "crazyflie_acados_sim_create: %s\n", cfnmpc_last_error());
        c->sim = nullptr;
        return 1;
    }
    cfnmpc_sim_opts_set(c->sim, "sens_forw", c->sens_forw ? 1 : 0);
    c->config.capsule = c; c->in.capsule = c; c->out.capsule = c; c->opts.capsule = c; c->solver.capsule = c;
    c->T = 0.015;  // Tsim = the shooting interval (acados_sim_solver.in.c:306-307)
    return 0;
}

int crazyflie_acados_sim_solve_capsule(crazyflie_sim_solver_capsule *c)
{
    if (!c || !c->sim) return ACADOS_QP_FAILURE;
    int rc = cfnmpc_sim_set(c->sim, "x", c->x, 0);
    rc |= cfnmpc_sim_set(c->sim, "u", c->u, 0);
    rc |= cfnmpc_sim_set(c->sim, "T", &c->T, 0);
    rc |= cfnmpc_sim_solve(c->sim);
    rc |= cfnmpc_sim_get(c->sim, "xn", c->xn, 0);
    c->have_S = false;
    if (!rc && c->sens_forw) { rc |= cfnmpc_sim_get(c->sim, "S_forw", c->S, 0); c->have_S = !rc; }
    if (rc) {
        fprintf(stderr, "crazyflie_acados_sim_solve: %s\n", cfnmpc_last_error());
        return ACADOS_QP_FAILURE;
    }
    for (int i = 0; i < 13; i++) if (c->xn[i] != c->xn[i]) return ACADOS_NAN_DETECTED;  // sim_erk reports NaNs the same way
    return ACADOS_SUCCESS;
}

sim_config *crazyflie_acados_get_sim_config(crazyflie_sim_solver_capsule *c) { return &c->config; }
sim_in *crazyflie_acados_get_sim_in(crazyflie_sim_solver_capsule *c) { return &c->in; }
sim_out *crazyflie_acados_get_sim_out(crazyflie_sim_solver_capsule *c) { return &c->out; }
void *crazyflie_acados_get_sim_dims(crazyflie_sim_solver_capsule *c) { return &c->dims_dummy; }
sim_opts *crazyflie_acados_get_sim_opts(crazyflie_sim_solver_capsule *c) { return &c->opts; }
sim_solver *crazyflie_acados_get_sim_solver(crazyflie_sim_solver_capsule *c) { return &c->solver; }

int sim_in_set(void *, void *, sim_in *in, const char *field, void *value)
{
    if (!in || !in->capsule || !field || !value) return 1;
    crazyflie_sim_solver_capsule *c = in->capsule;
    if (!strcmp(field, "T")) c->T = *static_cast<double *>(value);
    else if (!strcmp(field, "x")) memcpy(c->x, value, sizeof c->x);
    else if (!strcmp(field, "u")) memcpy(c->u, value, sizeof c->u);
    else return 1;
    return 0;
}

int sim_out_get(void *, void *, sim_out *out, const char *field, void *value)
{
    if (!out || !out->capsule || !field || !value) return 1;
    crazyflie_sim_solver_capsule *c = out->capsule;
    if (!strcmp(field, "xn") || !strcmp(field, "x")) memcpy(value, c->xn, sizeof c->xn);
    else if (!strcmp(field, "S_forw") && c->have_S) memcpy(value, c->S, sizeof c->S);
    else return 1;
    return 0;
}

void sim_opts_set(sim_config *config, void *, const char *field, void *value)
{
    if (!config || !config->capsule || !field || !value) return;
    crazyflie_sim_solver_capsule *c = config->capsule;
    if (!strcmp(field, "sens_forw")) {
        c->sens_forw = *static_cast<bool *>(value);
        if (c->sim) cfnmpc_sim_opts_set(c->sim, "sens_forw", c->sens_forw ? 1 : 0);
    } else if (!strcmp(field, "num_steps") || !strcmp(field, "num_stages")) {
        if (c->sim && cfnmpc_sim_opts_set(c->sim, field, *static_cast<int *>(value)) != CFNMPC_OK)
            fprintf(stderr, "sim_opts_set: %s\n", cfnmpc_last_error());
    }
}

int sim_solve(sim_solver *solver, sim_in *, sim_out *)
{
    if (!solver || !solver->capsule) return ACADOS_QP_FAILURE;
    return crazyflie_acados_sim_solve_capsule(solver->capsule);
}

__attribute__((weak)) sim_config *crazyflie_sim_config;
__attribute__((weak)) void *crazyflie_sim_dims;
__attribute__((weak)) sim_in *crazyflie_sim_in;
__attribute__((weak)) sim_out *crazyflie_sim_out;
__attribute__((weak)) sim_opts *crazyflie_sim_opts;
__attribute__((weak)) sim_solver *crazyflie_sim_solver;
static crazyflie_sim_solver_capsule *g_sim_capsule = nullptr;

int crazyflie_acados_sim_free_legacy(void)
{
    if (!g_sim_capsule) return 0;
    crazyflie_acados_sim_solver_free_capsule(g_sim_capsule);
    g_sim_capsule = nullptr;
    crazyflie_sim_config = nullptr; crazyflie_sim_dims = nullptr; crazyflie_sim_in = nullptr; crazyflie_sim_out = nullptr;
    crazyflie_sim_opts = nullptr; crazyflie_sim_solver = nullptr;
    return 0;
}

int crazyflie_acados_sim_create_legacy(void)
{
    crazyflie_acados_sim_free_legacy();
    g_sim_capsule = crazyflie_acados_sim_solver_create_capsule();
    if (crazyflie_acados_sim_create_capsule_solver(g_sim_capsule)) {
        crazyflie_acados_sim_solver_free_capsule(g_sim_capsule);
        g_sim_capsule = nullptr;
        return 1;
    }
    crazyflie_sim_config = &g_sim_capsule->config; crazyflie_sim_dims = &g_sim_capsule->dims_dummy;
    crazyflie_sim_in = &g_sim_capsule->in; crazyflie_sim_out = &g_sim_capsule->out;
    crazyflie_sim_opts = &g_sim_capsule->opts; crazyflie_sim_solver = &g_sim_capsule->solver;
    return 0;
}

int crazyflie_acados_sim_solve_legacy(void) { return g_sim_capsule ? crazyflie_acados_sim_solve_capsule(g_sim_capsule) : ACADOS_QP_FAILURE; }

}  // extern "C"
