/* Drop-in for the generated header acados_sim_solver_crazyflie.h (the estimator node's state predictor).
 *
 * Capsule API of the vendored acados: c_templates_tera/acados_sim_solver.in.h:81-95.
 * Legacy API of the 2020 template the estimator was written against: un-prefixed-capsule
 *   crazyflie_acados_sim_create() / _solve() / _free() and the process globals crazyflie_sim_config, _dims, _in,
 *   _out, _opts, _solver (crazyflie_controller/src/acados_estimator.cpp:237,573-593).
 * In C++ the two spellings are overloads; in C only the capsule API is declared (use the legacy names through
 * crazyflie_acados_sim_create_legacy & co).  Both drive a cfnmpc_sim batch of one instance (include/cfnmpc.h).
 */
#ifndef ACADOS_SIM_crazyflie_H_
#define ACADOS_SIM_crazyflie_H_

#include "acados_c/sim_interface.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crazyflie_sim_solver_capsule crazyflie_sim_solver_capsule;

crazyflie_sim_solver_capsule *crazyflie_acados_sim_solver_create_capsule(void);
int crazyflie_acados_sim_solver_free_capsule(crazyflie_sim_solver_capsule *capsule);
int crazyflie_acados_sim_create_capsule_solver(crazyflie_sim_solver_capsule *capsule);
int crazyflie_acados_sim_solve_capsule(crazyflie_sim_solver_capsule *capsule);
int crazyflie_acados_sim_free_capsule_solver(crazyflie_sim_solver_capsule *capsule);

sim_config *crazyflie_acados_get_sim_config(crazyflie_sim_solver_capsule *capsule);
sim_in *crazyflie_acados_get_sim_in(crazyflie_sim_solver_capsule *capsule);
sim_out *crazyflie_acados_get_sim_out(crazyflie_sim_solver_capsule *capsule);
void *crazyflie_acados_get_sim_dims(crazyflie_sim_solver_capsule *capsule);
sim_opts *crazyflie_acados_get_sim_opts(crazyflie_sim_solver_capsule *capsule);
sim_solver *crazyflie_acados_get_sim_solver(crazyflie_sim_solver_capsule *capsule);

/* legacy surface (process globals; weak definitions in the library, the 2020 generated code defined them) */
extern sim_config *crazyflie_sim_config;
extern void *crazyflie_sim_dims;
extern sim_in *crazyflie_sim_in;
extern sim_out *crazyflie_sim_out;
extern sim_opts *crazyflie_sim_opts;
extern sim_solver *crazyflie_sim_solver;
int crazyflie_acados_sim_create_legacy(void);
int crazyflie_acados_sim_solve_legacy(void);
int crazyflie_acados_sim_free_legacy(void);

#ifdef __cplusplus
}
/* the names the reference's two templates use, as C++ overloads */
static inline int crazyflie_acados_sim_create(void) { return crazyflie_acados_sim_create_legacy(); }
static inline int crazyflie_acados_sim_solve(void) { return crazyflie_acados_sim_solve_legacy(); }
static inline int crazyflie_acados_sim_free(void) { return crazyflie_acados_sim_free_legacy(); }
static inline int crazyflie_acados_sim_create(crazyflie_sim_solver_capsule *c) { return crazyflie_acados_sim_create_capsule_solver(c); }
static inline int crazyflie_acados_sim_solve(crazyflie_sim_solver_capsule *c) { return crazyflie_acados_sim_solve_capsule(c); }
static inline int crazyflie_acados_sim_free(crazyflie_sim_solver_capsule *c) { return crazyflie_acados_sim_free_capsule_solver(c); }
#endif
#endif
