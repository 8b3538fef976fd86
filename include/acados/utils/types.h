/* Status codes of the drop-in solver: same names and values as the reference,
 * acados/acados/utils/types.h:75-83. */
#ifndef ACADOS_UTILS_TYPES_H_
#define ACADOS_UTILS_TYPES_H_
enum return_values
{
    ACADOS_SUCCESS,
    ACADOS_NAN_DETECTED,
    ACADOS_MAXITER,
    ACADOS_MINSTEP,
    ACADOS_QP_FAILURE,
    ACADOS_READY,
};
#endif
