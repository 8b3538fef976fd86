// The per-tick call sequence of the reference's ROS node, NMPC::iteration()
// (crazyflie_controller/src/acados_mpc.cpp:427-718), written against the drop-in headers exactly as the
// node includes them (:61-73) and using the globals the node defines (:76-84).  Compiling and linking this
// file against include/ + libcfnmpc.so is the source-compatibility check; on a GPU box it prints u0, u1, x4
// of a few consecutive ticks for the parity test (tests/test_gpu_dropin.py).
#include <cmath>
#include <cstdio>
#include <cstdlib>

// acados
#include "acados/utils/print.h"
#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/external_function_interface.h"
#include "acados/ocp_nlp/ocp_nlp_constraints_bgh.h"
#include "acados/ocp_nlp/ocp_nlp_cost_ls.h"
// blasfeo
#include "blasfeo/include/blasfeo_d_aux.h"
#include "blasfeo/include/blasfeo_d_aux_ext_dep.h"
// crazyflie specific
#include "crazyflie_model/crazyflie_model.h"
#include "acados_solver_crazyflie.h"

// global data (defined by the application, as in the node)
ocp_nlp_in *nlp_in;
ocp_nlp_out *nlp_out;
ocp_nlp_solver *nlp_solver;
void *nlp_opts;
ocp_nlp_plan *nlp_solver_plan;
ocp_nlp_config *nlp_config;
ocp_nlp_dims *nlp_dims;
external_function_param_casadi *forw_vde_casadi;

#define N 50
#define NX 13
#define NU 4
#define NY 17
#define NYN 13

int main(int argc, char **argv)
{
    const int ticks = argc > 1 ? atoi(argv[1]) : 3;
    int status = acados_create();
    if (status) {
        fprintf(stderr, "acados_create() returned status %d. Exiting.\n", status);
        return 3;
    }
    // regulation reference as the node builds it (:435-454): set-point (0,0,0.4), float hover speed, g0 = 9.80665
    const float mq = 33e-3f, Ct = 3.25e-4f;
    const double g0 = 9.80665;
    const float uss = sqrt((mq * g0) / (4 * Ct));   // as the node: float mq, Ct, uss; g0 a double macro (acados_mpc.cpp:107,189,253)
    double x0[NX] = {0.1, -0.05, 0.3, 1, 0, 0, 0, 0.1, 0, -0.1, 0, 0, 0};
    double yref_sign[(N + 1) * NY];
    for (int k = 0; k <= N; k++) {
        double *y = yref_sign + k * NY;
        for (int i = 0; i < NY; i++) y[i] = 0.0;
        y[2] = 0.40; y[3] = 1.0;
        for (int i = 0; i < NU; i++) y[NX + i] = uss;
    }
    for (int t = 0; t < ticks; t++) {
        ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, 0, "lbx", x0);
        ocp_nlp_constraints_model_set(nlp_config, nlp_dims, nlp_in, 0, "ubx", x0);
        for (int k = 0; k < N + 1; k++) ocp_nlp_cost_model_set(nlp_config, nlp_dims, nlp_in, k, "yref", yref_sign + k * NY);
        int acados_status = acados_solve();
        double u0[NU], u1[NU], x4[NX];
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 0, "u", u0);
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 1, "u", u1);
        ocp_nlp_out_get(nlp_config, nlp_dims, nlp_out, 4, "x", x4);
        printf("tick %d status %d time %.6f", t, acados_status, nlp_out->total_time);
        for (int i = 0; i < NU; i++) printf(" %.15e", u0[i]);
        for (int i = 0; i < NU; i++) printf(" %.15e", u1[i]);
        for (int i = 0; i < NX; i++) printf(" %.15e", x4[i]);
        printf("\n");
    }
    status = acados_free();
    return status;
}
