/* Drop-in subset of the reference's acados_c/ocp_nlp_interface.h for the Crazyflie solver.
 *
 * Same function names, argument order and meaning as
 *   acados/interfaces/acados_c/ocp_nlp_interface.h:221-234 (cost / constraints setters),
 *   :262,274-275 (ocp_nlp_out_set / ocp_nlp_out_get), :313 (opts), :352 (ocp_nlp_solve), :408 (ocp_nlp_get)
 * which are the entry points `crazyflie_controller/src/acados_mpc.cpp:581-625` binds.  The
 * objects are handles into a cfnmpc batch of one instance (include/cfnmpc.h); their layouts
 * are private except for the two fields of ocp_nlp_out the node reads directly (:615-616).
 * Differences, all on error paths: unknown fields return a non-zero code (setters) or leave
 * the destination untouched (getters) instead of printing and calling exit(1).
 */
#ifndef ACADOS_C_OCP_NLP_INTERFACE_H_
#define ACADOS_C_OCP_NLP_INTERFACE_H_

#include "acados/utils/types.h"

#ifdef __cplusplus
extern "C" {
#endif

struct crazyflie_solver_capsule;

typedef struct ocp_nlp_plan_t
{
    int N;
} ocp_nlp_plan_t;
typedef ocp_nlp_plan_t ocp_nlp_plan; /* 2020 spelling used by the node (acados_mpc.cpp:80) */

typedef struct ocp_nlp_config
{
    int N;
    struct crazyflie_solver_capsule *capsule;
} ocp_nlp_config;

typedef struct ocp_nlp_dims
{
    int N;
    int nx, nu, ny, ny_e;
    struct crazyflie_solver_capsule *capsule;
} ocp_nlp_dims;

typedef struct ocp_nlp_in
{
    struct crazyflie_solver_capsule *capsule;
} ocp_nlp_in;

typedef struct ocp_nlp_out
{
    double inf_norm_res; /* read by the node (:615); like the reference's SQP_RTI this is never written by a solve */
    double total_time;   /* read by the node (:616): time of the last solve [s] */
    struct crazyflie_solver_capsule *capsule;
} ocp_nlp_out;

typedef struct ocp_nlp_solver
{
    struct crazyflie_solver_capsule *capsule;
} ocp_nlp_solver;

/* field: "lbx" | "ubx" (stage 0, 13 doubles: the measured state), "lbu" | "ubu" (4 doubles; stage 0 has its own box
 * -- the node's FIXED_U0 branch, acados_mpc.cpp:604-608 -- stages 1..N-1 share one, the last value set wins).
 * Returns 0, or 1 for an unknown field / bad stage. */
int ocp_nlp_constraints_model_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_in *in, int stage,
                                  const char *field, void *value);
/* field: "yref" | "y_ref" (17 doubles for stage < N, 13 at stage N), "W" (column-major ny x ny; must be diagonal,
 * applies to all stages < N, or to the terminal cost at stage N). */
int ocp_nlp_cost_model_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_in *in, int stage, const char *field,
                           void *value);
/* field: "x" (13 doubles) | "u" (4 doubles) of the iterate at `stage` */
void ocp_nlp_out_set(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_out *out, int stage, const char *field,
                     void *value);
void ocp_nlp_out_get(ocp_nlp_config *config, ocp_nlp_dims *dims, ocp_nlp_out *out, int stage, const char *field,
                     void *value);
/* field: "rti_phase" (int 0 = preparation + feedback, 1 = preparation, 2 = feedback: ocp_nlp_sqp_rti.c:189-198,
 * 1213-1237; the next ocp_nlp_solve / acados_solve runs that phase), "qp_cond_N" (accepted; results do not depend on it),
 * "print_level" (ignored) */
void ocp_nlp_solver_opts_set(ocp_nlp_config *config, void *opts_, const char *field, void *value);
/* One RTI step (ocp_nlp_sqp_rti.c:1232-1237). Returns the acados status. */
int ocp_nlp_solve(ocp_nlp_solver *solver, ocp_nlp_in *nlp_in, ocp_nlp_out *nlp_out);
/* field: "time_tot" (double, s) | "qp_iter" | "sqp_iter" | "status" | "qp_status" (int) */
void ocp_nlp_get(ocp_nlp_config *config, ocp_nlp_solver *solver, const char *field, void *return_value_);

#ifdef __cplusplus
}
#endif
#endif
