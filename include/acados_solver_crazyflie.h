/* Drop-in for the generated header acados_solver_crazyflie.h.
 *
 * Surface B (capsule API of the vendored acados):
 *   acados/interfaces/acados_template/acados_template/c_templates_tera/acados_solver.in.h:237-275
 * Surface A (the 2020 template the ROS node was written against): un-prefixed
 *   acados_create / acados_solve / acados_free and process-global objects,
 *   crazyflie_controller/src/acados_mpc.cpp:76-84,225,418,611.
 * Both are one-instance wrappers over the batch C-ABI in cfnmpc.h; the solve runs on the GPU.
 */
#ifndef ACADOS_SOLVER_crazyflie_H_
#define ACADOS_SOLVER_crazyflie_H_

#include "acados_c/ocp_nlp_interface.h"

#define CRAZYFLIE_NX 13
#define CRAZYFLIE_NU 4
#define CRAZYFLIE_NY 17
#define CRAZYFLIE_NYN 13
#define CRAZYFLIE_N 50

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crazyflie_solver_capsule crazyflie_solver_capsule;

crazyflie_solver_capsule *crazyflie_acados_create_capsule(void);
int crazyflie_acados_free_capsule(crazyflie_solver_capsule *capsule);
int crazyflie_acados_create(crazyflie_solver_capsule *capsule);
/* other horizons / grids: n_time_steps shooting intervals of lengths new_time_steps[i] > 0, each also the scaling of its
 * stage cost (acados_solver.in.c:133-153,879-892); NULL keeps Ts = 0.015 s and needs n_time_steps == CRAZYFLIE_N
 * (:2381-2391).  crazyflie_acados_update_time_steps changes the grid of an existing solver (same N). */
int crazyflie_acados_create_with_discretization(crazyflie_solver_capsule *capsule, int n_time_steps, double *new_time_steps);
int crazyflie_acados_update_time_steps(crazyflie_solver_capsule *capsule, int N, double *new_time_steps);
int crazyflie_acados_update_qp_solver_cond_N(crazyflie_solver_capsule *capsule, int qp_solver_cond_N);
int crazyflie_acados_reset(crazyflie_solver_capsule *capsule, int reset_qp_solver_mem);
int crazyflie_acados_solve(crazyflie_solver_capsule *capsule);
int crazyflie_acados_free(crazyflie_solver_capsule *capsule);
void crazyflie_acados_print_stats(crazyflie_solver_capsule *capsule);

ocp_nlp_in *crazyflie_acados_get_nlp_in(crazyflie_solver_capsule *capsule);
ocp_nlp_out *crazyflie_acados_get_nlp_out(crazyflie_solver_capsule *capsule);
ocp_nlp_solver *crazyflie_acados_get_nlp_solver(crazyflie_solver_capsule *capsule);
ocp_nlp_config *crazyflie_acados_get_nlp_config(crazyflie_solver_capsule *capsule);
void *crazyflie_acados_get_nlp_opts(crazyflie_solver_capsule *capsule);
ocp_nlp_dims *crazyflie_acados_get_nlp_dims(crazyflie_solver_capsule *capsule);
ocp_nlp_plan_t *crazyflie_acados_get_nlp_plan(crazyflie_solver_capsule *capsule);

/* ---- legacy surface: the node defines these globals itself (acados_mpc.cpp:76-84); the library
 * provides weak definitions so that other hosts (ctypes, tests) can link too. */
extern ocp_nlp_in *nlp_in;
extern ocp_nlp_out *nlp_out;
extern ocp_nlp_solver *nlp_solver;
extern void *nlp_opts;
extern ocp_nlp_plan *nlp_solver_plan;
extern ocp_nlp_config *nlp_config;
extern ocp_nlp_dims *nlp_dims;

int acados_create(void);
int acados_solve(void);
int acados_free(void);

#ifdef __cplusplus
}
#endif
#endif
