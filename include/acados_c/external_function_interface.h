/* Source-compatibility stub: the node includes this header (acados_mpc.cpp:63) and declares one
 * `external_function_param_casadi *` global (:84) that it never uses.  The model is compiled into the
 * CUDA kernels (crazyflie_nmpc_b200/csrc/cf_model.h); there are no external functions to manage. */
#ifndef ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H_
#define ACADOS_C_EXTERNAL_FUNCTION_INTERFACE_H_
typedef struct external_function_param_casadi external_function_param_casadi;
typedef struct external_function_casadi external_function_casadi;
#endif
