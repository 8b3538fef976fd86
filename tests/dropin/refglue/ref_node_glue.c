/* TEST INFRASTRUCTURE ONLY.  acados_create / acados_solve / acados_free for the node harness's REFERENCE backend: the
 * generated glue that is not in the reference tree, written over oracle/ref_harness.c (cfref_create builds the OCP the
 * way acados_solver.in.c does) and the node's own process globals (acados_mpc.cpp:76-84).  acados_solve() optionally logs
 * what the node handed to the solver and what came back:
 *   per tick doubles: x0(13) from lbx_0, yref (N*17), yref_e (13), status, u0(4), u1(4), x4(13), qp_iter */
#include <stdio.h>
#include <stdlib.h>

#include "acados/ocp_nlp/ocp_nlp_common.h"
#include "acados/ocp_nlp/ocp_nlp_constraints_bgh.h"
#include "acados/ocp_nlp/ocp_nlp_cost_ls.h"
#include "acados_c/external_function_interface.h"
#include "acados_c/ocp_nlp_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "acados_solver_crazyflie.h"

extern ocp_nlp_in *nlp_in;
extern ocp_nlp_out *nlp_out;
extern ocp_nlp_solver *nlp_solver;
extern void *nlp_opts;
extern ocp_nlp_plan *nlp_solver_plan;
extern ocp_nlp_config *nlp_config;
extern ocp_nlp_dims *nlp_dims;

void *cfref_create(int N, double Ts, int cond_N);
void cfref_destroy(void *h);
void cfref_export(void *h, void **plan, void **config, void **dims, void **in, void **out, void **opts, void **solver);

#define GN 50
static void *g_h;
static FILE *g_log;

void cf_glue_set_log(const char *path) { g_log = fopen(path, "wb"); }

int acados_create(void)
{
    g_h = cfref_create(GN, 0.015, 0);
    if (!g_h) return 1;
    cfref_export(g_h, (void **) &nlp_solver_plan, (void **) &nlp_config, (void **) &nlp_dims, (void **) &nlp_in, (void **) &nlp_out,
                 &nlp_opts, (void **) &nlp_solver);
    return 0;
}

int acados_solve(void)
{
    double v[GN * 17 + 64];
    if (g_log) {
        ocp_nlp_constraints_bgh_model *cm = nlp_in->constraints[0];
        blasfeo_unpack_dvec(13, &cm->d, 4, v, 1);   /* lower bounds [lbu(4); lbx(13)] (ocp_nlp_constraints_bgh.c:647-666) */
        fwrite(v, 8, 13, g_log);
        for (int k = 0; k < GN; k++) {
            ocp_nlp_cost_ls_model *m = nlp_in->cost[k];
            blasfeo_unpack_dvec(17, &m->y_ref, 0, v + 17 * k, 1);
        }
        fwrite(v, 8, GN * 17, g_log);
        ocp_nlp_cost_ls_model *m = nlp_in->cost[GN];
        blasfeo_unpack_dvec(13, &m->y_ref, 0, v, 1);
        fwrite(v, 8, 13, g_log);
    }
    int status = ocp_nlp_solve(nlp_solver, nlp_in, nlp_out);
    if (g_log) {
        int qi = 0;
        v[0] = status;
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 0, "u", v + 1);
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 1, "u", v + 5);
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 4, "x", v + 9);
        ocp_nlp_get(nlp_config, nlp_solver, "qp_iter", &qi);
        v[22] = qi;
        fwrite(v, 8, 23, g_log);
        fflush(g_log);
    }
    return status;
}

int acados_free(void)
{
    cfref_destroy(g_h);
    g_h = NULL;
    if (g_log) { fclose(g_log); g_log = NULL; }
    return 0;
}
