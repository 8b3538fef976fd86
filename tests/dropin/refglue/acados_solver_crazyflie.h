/* Reference backend of the node harness only (tests/dropin/build_node.py, node_ref): the three entry points of the
 * 2020 generated solver that the node calls (acados_mpc.cpp:225,418,611); implemented in ref_node_glue.c on top of the
 * reference's own acados build (oracle/_ref/libcfref.so). */
#ifndef CF_REFGLUE_ACADOS_SOLVER_CRAZYFLIE_H
#define CF_REFGLUE_ACADOS_SOLVER_CRAZYFLIE_H
#include "acados_c/ocp_nlp_interface.h"
/* the node was written against the 2020 acados API, where the plan type was called ocp_nlp_plan; the vendored acados
 * calls it ocp_nlp_plan_t (acados/interfaces/acados_c/ocp_nlp_interface.h:105-140) */
typedef ocp_nlp_plan_t ocp_nlp_plan;
#ifdef __cplusplus
extern "C" {
#endif
int acados_create(void);
int acados_solve(void);
int acados_free(void);
#ifdef __cplusplus
}
#endif
#endif
