/* Drop-in subset of the reference's acados_c/sim_interface.h for the Crazyflie state predictor.
 *
 * Same names, argument order and meaning as acados/interfaces/acados_c/sim_interface.h:96 (sim_in_set),
 * :105 (sim_out_get), :113 (sim_opts_set), :127 (sim_solve) -- the calls
 * crazyflie_controller/src/acados_estimator.cpp:573-593 makes.  The objects are handles into a cfnmpc_sim batch of
 * one instance (include/cfnmpc.h); the integration runs on the GPU.  Unknown fields return non-zero instead of
 * exiting the process.
 */
#ifndef ACADOS_C_SIM_INTERFACE_H_
#define ACADOS_C_SIM_INTERFACE_H_

#ifdef __cplusplus
extern "C" {
#endif

struct crazyflie_sim_solver_capsule;

typedef struct sim_config { struct crazyflie_sim_solver_capsule *capsule; } sim_config;
typedef struct sim_in { struct crazyflie_sim_solver_capsule *capsule; } sim_in;
typedef struct sim_out { struct crazyflie_sim_solver_capsule *capsule; } sim_out;
typedef struct sim_opts { struct crazyflie_sim_solver_capsule *capsule; } sim_opts;
typedef struct sim_solver { struct crazyflie_sim_solver_capsule *capsule; } sim_solver;

/* field: "T" (double: horizon [s]), "x" (13 doubles), "u" (4 doubles) */
int sim_in_set(void *config, void *dims, sim_in *in, const char *field, void *value);
/* field: "xn" | "x" (13 doubles), "S_forw" (13 x 17 column-major, columns [x | u]; needs sens_forw) */
int sim_out_get(void *config, void *dims, sim_out *out, const char *field, void *value);
/* field: "num_steps" (int), "num_stages" (int, must be 4), "sens_forw" (bool) */
void sim_opts_set(sim_config *config, void *opts, const char *field, void *value);
int sim_solve(sim_solver *solver, sim_in *in, sim_out *out);

#ifdef __cplusplus
}
#endif
#endif
