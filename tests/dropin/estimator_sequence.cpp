// The state-prediction call sequence of the reference's estimator node
// (crazyflie_controller/src/acados_estimator.cpp:60-76 includes, :237 create, :573-593 per message), written
// against the drop-in headers.  Prints xn of a few predictions for the parity test.
#include <cstdio>
#include <cstdlib>

#include "acados/utils/print.h"
#include "acados_c/ocp_nlp_interface.h"
#include "acados_c/external_function_interface.h"
#include "blasfeo/include/blasfeo_d_aux.h"
#include "blasfeo/include/blasfeo_d_aux_ext_dep.h"
#include "crazyflie_model/crazyflie_model.h"
#include "acados_solver_crazyflie.h"
#include "acados_sim_solver_crazyflie.h"

#define NX 13
#define NU 4

external_function_param_casadi *forw_vde_casadi;

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 3;
    int status = crazyflie_acados_sim_create();
    if (status) {
        fprintf(stderr, "acados_sim_create() returned status %d. Exiting.\n", status);
        return 3;
    }
    double x0[NX] = {0.1, -0.05, 0.3, 0.995, 0.05, -0.03, 0.02, 0.1, 0.0, -0.1, 0.2, -0.1, 0.3};
    double u0[NU] = {15, 16, 14, 17};   // int32 motor speeds as received from /crazyflie/acados_motvel
    for (int t = 0; t < n; t++) {
        double delay = 0.015 + 0.005 * t;
        double xn[NX];
        sim_in_set(crazyflie_sim_config, crazyflie_sim_dims, crazyflie_sim_in, "T", &delay);
        sim_in_set(crazyflie_sim_config, crazyflie_sim_dims, crazyflie_sim_in, "x", x0);
        sim_in_set(crazyflie_sim_config, crazyflie_sim_dims, crazyflie_sim_in, "u", u0);
        int sim_acados_status = crazyflie_acados_sim_solve();
        sim_out_get(crazyflie_sim_config, crazyflie_sim_dims, crazyflie_sim_out, "xn", xn);
        printf("pred %d status %d T %.6f", t, sim_acados_status, delay);
        for (int i = 0; i < NX; i++) printf(" %.17e", xn[i]);
        printf("\n");
        for (int i = 0; i < NX; i++) x0[i] = xn[i];
    }
    return crazyflie_acados_sim_free();
}
