/* TEST INFRASTRUCTURE ONLY (oracle/): Crazyflie ODE + forward variational
 * equations handed to the reference's ERK integrator as external functions.
 *
 * The reference generates these with CasADi (not available here):
 *   crazyflie_controller/scripts/crazyflie_full_model/export_ode_model.py:34-42,85-101
 *   acados_template/casadi_function_generation.py:137-151  (expl_vde_forw)
 * This is a hand-derived analytic restatement of the same functions. */
#ifndef CF_MODEL_REF_H
#define CF_MODEL_REF_H

#define CF_NX 13
#define CF_NU 4

/* f(x,u) -> xdot[13] */
void cf_ref_ode(const double *x, const double *u, double *xdot);
/* (x, Sx[13x13 colmaj], Su[13x4 colmaj], u) -> (f, Jx*Sx, Ju + Jx*Su) */
void cf_ref_vde_forw(const double *x, const double *Sx, const double *Su, const double *u,
                     double *f, double *dSx, double *dSu);

#endif
