/* Source-compatibility stub.  The reference node includes this path
 * (crazyflie_controller/src/acados_mpc.cpp:61-73) but uses nothing from it on the NMPC path;
 * the drop-in library has no BLASFEO / module-level objects to expose. */
#ifndef CFNMPC_STUB_BLASFEO_INCLUDE_BLASFEO_D_AUX_H
#define CFNMPC_STUB_BLASFEO_INCLUDE_BLASFEO_D_AUX_H
#endif
