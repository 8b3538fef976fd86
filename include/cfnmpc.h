/* cfnmpc.h -- batch C-ABI of the B200-native Crazyflie NMPC solver.
 *
 * One handle owns B independent instances of the OCP that
 * `crazyflie_controller` solves (N shooting intervals, nx = 13, nu = 4) and
 * advances all of them by one real-time-iteration SQP step per
 * cfnmpc_batch_solve() on one GPU, one warp per instance.
 *
 * Reference interface this replaces.  The reference has no batch notion; per
 * instance and per control tick its ROS node does
 *   ocp_nlp_constraints_model_set(.., 0, "lbx"/"ubx", x0)   crazyflie_controller/src/acados_mpc.cpp:581-582
 *   ocp_nlp_cost_model_set(.., k, "yref", ..) k = 0..N      :590-594
 *   acados_solve()                                           :611
 *   ocp_nlp_out_get(.., 0, "u"), (1,"u"), (4,"x")            :619-625
 * against acados/interfaces/acados_c/ocp_nlp_interface.h:221-234,274-275,352.
 * The calls below are the same operations with a leading batch dimension:
 * cfnmpc_batch_set("x0"|"yref"|...) <-> the two *_model_set calls,
 * cfnmpc_batch_solve <-> acados_solve / ocp_nlp_solve,
 * cfnmpc_batch_get("u"|"x", stage) <-> ocp_nlp_out_get.
 * The single-instance surfaces (acados_solver_crazyflie.h, acados_c/...) are thin
 * B = 1 wrappers over this file.
 *
 * Conventions (same as the reference): plain pointers and sizes,
 * fp64, vectors contiguous, the library copies in setters and copies out in
 * getters and owns all solver memory.  No call aborts the process: errors are
 * negative return values and cfnmpc_last_error() has the text.  Not re-entrant
 * per handle; different handles may be used from different threads.
 */
#ifndef CFNMPC_H
#define CFNMPC_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFNMPC_NX 13
#define CFNMPC_NU 4
#define CFNMPC_NY 17

/* return codes */
#define CFNMPC_OK 0
#define CFNMPC_EINVAL (-1)  /* bad argument / unknown field */
#define CFNMPC_ECUDA (-2)   /* CUDA runtime error (no device, launch failure, out of memory) */
#define CFNMPC_ESTATE (-3)  /* call sequence error */

/* per-instance status values written by a solve: acados/acados/utils/types.h:75-83 */
#define CFNMPC_SUCCESS 0
#define CFNMPC_NAN_DETECTED 1
#define CFNMPC_MAXITER 2
#define CFNMPC_MINSTEP 3
#define CFNMPC_QP_FAILURE 4

typedef struct cfnmpc_batch cfnmpc_batch;

/* Create a solver for `batch` instances with horizon N and grid step Ts [s]
 * on CUDA device `device`.  Weights and bounds start at the values of
 * crazyflie_controller/scripts/crazyflie_full_model/generate_c_code.py:61-136; the
 * iterate starts as the generated solver's does (x_k = [0,0,0,1,0..], u_k = 0,
 * acados_solver.in.c:2323-2352). */
int cfnmpc_batch_create(int batch, int N, double Ts, int device, cfnmpc_batch **out);
int cfnmpc_batch_destroy(cfnmpc_batch *h);

/* Run all work of this handle on an existing CUDA stream (a cudaStream_t passed as
 * void*); NULL returns to the handle's own stream. */
int cfnmpc_batch_set_stream(cfnmpc_batch *h, void *cuda_stream);

/* Copy caller data into the solver.  `src_on_device` != 0 means `src` is a device
 * pointer on the handle's GPU.  Fields and shapes (row-major, instance-major):
 *   "x0"      double [B][13]        measured state -> lbx_0 = ubx_0
 *   "yref"    double [B][N][17]     stage references, y = [x(13); u(4)]
 *   "yref_e"  double [B][13]        terminal reference
 *   "x"       double [B][N+1][13]   iterate (states)
 *   "u"       double [B][N][4]      iterate (inputs)
 *   "W"       double [17]           diagonal of the stage weight matrix (all instances)
 *   "W_e"     double [13]           diagonal of the terminal weight matrix
 *   "lbu","ubu" double [4]          input bounds, stages 0..N-1
 *   "lbu0","ubu0" double [4]        input bounds of stage 0 only (set after "lbu"/"ubu"): the node's FIXED_U0
 *                                   branch pins u_0 this way, acados_mpc.cpp:604-608
 *   "bounds_stage" double [N][8]    input box per STAGE, row k = lbu(4) | ubu(4) of stage k: the reference sets bounds one
 *                                   stage at a time (ocp_nlp_constraints_model_set(.., k, "lbu"|"ubu", ..),
 *                                   ocp_nlp_constraints_bgh.c:653-674).  Once given it takes precedence over "lbu","ubu",
 *                                   "lbu0","ubu0" and the per-instance arrays until cfnmpc_batch_clear("bounds_stage").
 *   "W_stage" double [N+1][17]      weight diagonals per STAGE (row k < N: W_k in cost order y = [x;u]; row N: W_e, 13 used): the
 *                                   reference sets the weight of one stage at a time (ocp_nlp_cost_model_set(.., k, "W", ..),
 *                                   ocp_nlp_cost_ls.c:301-331).  Takes precedence over "W", "W_e" and the per-instance arrays
 *                                   until cfnmpc_batch_clear("W_stage"); runs the general kernels.
 *   "W_dense_table" double [N+1][17][17]  FULL weight matrix per stage (row-major, symmetric positive definite, cost order
 *                                   y = [x;u]; row N: W_e in its leading 13 x 13 block): the reference accepts any SPD W
 *                                   (ocp_nlp_cost_ls.c:301-331; Hessian scaling (Cyt W_chol)(Cyt W_chol)', :743-772, gradient
 *                                   scaling Cyt W res, :883-912).  Takes precedence over every other weight field until
 *                                   cfnmpc_batch_clear("W_dense_table").  The step then runs as preparation kernel + the
 *                                   condensed feedback program with block size 1 (the program with a dense stage Hessian,
 *                                   csrc/cf_pcond_warp.h); not available together with "qp_cond_N" < N or "multipliers".
 *   "time_steps" double [N]         lengths of the shooting intervals (host or device pointer); each is also the
 *                                   scaling of its stage cost, as crazyflie_acados_create_with_discretization /
 *                                   crazyflie_acados_update_time_steps set them (c_templates_tera/acados_solver.in.c:
 *                                   133-153).  A uniform grid runs the specialised kernel, anything else the general one.
 * Per-instance parameter arrays (one value set per vehicle; the reference expresses this as one solver object per
 * vehicle, each with its own ocp_nlp_cost_model_set(..,"W",..) / ocp_nlp_constraints_model_set(..,"lbu"|"ubu",..)
 * calls, acados_mpc.cpp:596-608).  Once given they take precedence over the solver-wide value until cleared:
 *   "W_batch"    double [B][17]     "W_e_batch"  double [B][13]
 *   "lbu_batch","ubu_batch"   double [B][4]   stages 0..N-1
 *   "lbu0_batch","ubu0_batch" double [B][4]   stage 0 only
 * Copies are asynchronous on the handle's stream. */
int cfnmpc_batch_set(cfnmpc_batch *h, const char *field, const void *src, int src_on_device);
/* Stop using a per-instance parameter array ("W_batch", ...) or the per-stage table "bounds_stage": back to the solver-wide value. */
int cfnmpc_batch_clear(cfnmpc_batch *h, const char *field);

/* Integer options:
 *   "lin_res_check" (default 0)  1: evaluate, after every Riccati solve, the linear-system residuals HPIPM evaluates for
 *                   its safety nets (x_ocp_qp_ipm.c:2029-2059 LQ re-factorisation, :2311-2318 iterative refinement) and
 *                   report in "flags" where the reference would have taken one of them.  1 is diagnostic only (results do
 *                   not depend on it).  2 also performs the second net, iterative refinement of the corrector step (at most
 *                   two rounds, itref_corr_max = 2; "flags" bit 5, bit 6 when two rounds were not enough).  The first net has
 *                   no counterpart: HPIPM's LQ re-factorisation exists for its square-root recursion only and falls through
 *                   to the plain factorisation for the classical recursion used here (x_ocp_qp_kkt.c:769-774).
 *   "max_ipm_iter"  (default 50 = qp_solver_iter_max of the reference configuration)
 *   "two_kernels"   (default 1) cfnmpc_batch_solve / _solve_from_host / _tick run the step as two launches -- preparation
 *                   of every instance, then feedback of every instance, the linearisations travelling through a
 *                   per-instance store of 8 * (252 N + 18) bytes -- instead of one fused kernel (0).  Same results bit for
 *                   bit; 2 % faster on B200 (profiles/README.md).  Falls back to the fused kernel when the store cannot
 *                   be allocated.  Also CFNMPC_TWO_KERNELS=0 in the environment.
 *   "qp_cond_N"     (default 0 = N: the reference's own configuration) partial condensing of the stage-wise QP to this many
 *                   stages before the interior-point solve = qp_cond_N of the reference (ocp_nlp_solver_opts_set /
 *                   crazyflie_acados_update_qp_solver_cond_N; acados/acados/ocp_qp/ocp_qp_partial_condensing.c:235-258,
 *                   457-576 -> external/hpipm/cond/x_part_cond.c:36-54,505-560,658-742).  Blocks of up to 3 stages are
 *                   implemented (qp_cond_N >= ceil(N/3); coarser values return CFNMPC_EINVAL).  Results equal the
 *                   reference's at the same qp_cond_N (~1e-13) and its qp_cond_N = N results to the interior-point
 *                   tolerances.  The step runs as preparation kernel + condensed feedback kernel; "lin_res_check" is not
 *                   available on this path.
 *   "multipliers"   (default 0) 1: keep the multipliers of every solved instance = what ocp_nlp_out_get "pi" / "lam" / "t"
 *                   hand out after a step in the reference (acados_c/ocp_nlp_interface.c:576-590; full-step duals,
 *                   ocp_nlp_common.c:2917-2925), 8 (29 N + 182) bytes per instance; read them with cfnmpc_batch_get
 *                   "pi", "lam", "t", "lam_x0".  Not available together with "qp_cond_N" < N. */
int cfnmpc_batch_set_option(cfnmpc_batch *h, const char *option, int value);

/* Enqueue n_rti consecutive RTI steps (preparation + feedback) for every instance,
 * inputs frozen between steps.  Asynchronous; pair with cfnmpc_batch_sync or a
 * getter to a host pointer. */
int cfnmpc_batch_solve(cfnmpc_batch *h, int n_rti);
/* The two halves of a real-time iteration as separate calls = rti_phase 1 (PREPARATION) and 2 (FEEDBACK) of the reference
 * (ocp_nlp_solver_opts_set(.., "rti_phase", ..), acados/acados/ocp_nlp/ocp_nlp_sqp_rti.c:189-198,495-683,1213-1237).
 * cfnmpc_batch_prepare linearises around the current iterate with the current "yref"/"yref_e"/weights (integration with
 * sensitivities, gradients) and keeps the result per instance (8 * (252 N + 18) bytes each, allocated on first use);
 * cfnmpc_batch_feedback then takes the "x0" (and bounds) of the moment, solves the QP and updates the iterate -- the
 * latency-critical half.  prepare + feedback with unchanged inputs gives bit-identical results to cfnmpc_batch_solve(h, 1);
 * feedback without a preparation that belongs to the current iterate returns CFNMPC_ESTATE. */
int cfnmpc_batch_prepare(cfnmpc_batch *h);
int cfnmpc_batch_feedback(cfnmpc_batch *h);
int cfnmpc_batch_sync(cfnmpc_batch *h);
/* One tick fed from HOST buffers (pinned memory for real overlap): equivalent to cfnmpc_batch_set of "x0", "yref",
 * "yref_e" followed by cfnmpc_batch_solve(h, 1), but the solve kernel starts at once and the inputs follow in n_chunks
 * (1..32) contiguous chunks on a second stream; a warp that reaches an instance whose inputs have not arrived yet waits
 * for the upload front.  This is the batched form of the node's per-tick ocp_nlp_constraints_model_set /
 * ocp_nlp_cost_model_set / acados_solve sequence (acados_mpc.cpp:581-611).  Getters on the handle's stream see the
 * results as after cfnmpc_batch_solve. */
int cfnmpc_batch_solve_from_host(cfnmpc_batch *h, const double *x0, const double *yref, const double *yref_e, int n_chunks);

/* Copy results out.  Fields:
 *   "u"        stage k in [0,N)   double [B][4]
 *   "x"        stage k in [0,N]   double [B][13]
 *   "u_all"    double [B][N][4]      "x_all"  double [B][N+1][13]      (stage ignored)
 *   "status"   int [B]   acados status of the last step (0 ok, 4 QP failure)
 *   "qp_iter"  int [B]   interior-point iterations of the last step
 *   "qp_status" int [B]  HPIPM status 0 ok / 1 max-iter / 2 min-step / 3 NaN
 *   "flags"    int [B]   bit 3: a non-positive Riccati pivot was replaced by 0; bit 4: step length or duality measure not
 *                        finite (both always on); bit 5/6: iterative refinement ran / left a residual (lin_res_check 2);
 *                        bit 0/1: the reference's LQ / iterative-refinement safety nets would have fired (needs the option
 *                        "lin_res_check"; 0 otherwise); bit 2: cfnmpc_batch_solve_from_host gave up waiting for this
 *                        instance's inputs (5 s) and solved it with whatever was in device memory
 *   "res"      double [B][4]  final QP residual inf-norms (stationarity, dynamics, bounds, complementarity)
 * and, with the option "multipliers" set before the solve (instances whose QP failed keep their previous values):
 *   "pi"       double [B][13]  multiplier of the dynamics x_{stage+1} = phi(x_stage, u_stage), stage 0..N-1
 *   "lam"      double [B][8]   multipliers of the input box of `stage`: lower(4) | upper(4), stage 0..N-1
 *   "t"        double [B][8]   the slacks of those bounds
 *   "pi_all" [B][N][13], "lam_all" [B][N][8], "t_all" [B][N][8]: every stage at once (stage ignored)
 *   "lam_x0"   double [B][13]  signed multiplier of the eliminated constraint x_0 = x0 (stage must be 0): a value v >= 0
 *                              is the lower-bound multiplier of the reference's lbx_0 (upper one 1e-16), v < 0 means the
 *                              upper-bound multiplier is -v (external/hpipm/ocp_qp/x_ocp_qp_red.c:820-840)
 * Copies to host pointers synchronise the stream before returning. */
int cfnmpc_batch_get(cfnmpc_batch *h, const char *field, int stage, void *dst, int dst_on_device);

/* ---- closed-loop driver: the rest of the node's control tick, on the device ----------------
 * Reference: NMPC::iteration, crazyflie_controller/src/acados_mpc.cpp:430-516 (reference window by policy),
 * :619-670 (what is published from the solution).  Additional cfnmpc_batch_set fields:
 *   "policy"    int    [B]      0 Regulation, 1 Tracking, 2 Position_Hold (enum order of acados_mpc.cpp:129-133)
 *   "traj_iter" int    [B]      row of the trajectory table the tracking window starts at (`iter` of the node)
 *   "setpoint"  double [B][3]   regulation point (xq_des, yq_des, zq_des)
 *   "uss"       double [1]      hover speed written into regulation / hold references; default = the node's float
 *                               computation sqrt(mq*g0/(4*Ct)) with g0 = 9.80665 (:107,189,253)
 * and cfnmpc_batch_get fields "policy", "traj_iter", "yref", "yref_e", "x0",
 *   "motors" int [B][4]      motor speeds truncated to int32 as PropellerSpeedsStamped carries them (:632-640)
 *   "euler"  double [B][3]   (phi, theta, psi) of the normalised quaternion of x_4 (quatern2euler :384-404)
 *   "twist"  double [B][4]   linear.x = pitch [deg], linear.y = -roll [deg], linear.z = krpm2pwm(mean(u_1)),
 *                            angular.z = yaw rate of x_4 [deg/s]  (:641-668) */
/* Trajectory table, n_rows x 17 row-major = the file format of crazyflie_controller/traj (acados_mpc.cpp:258-283). */
int cfnmpc_batch_set_trajectory(cfnmpc_batch *h, const double *table, int n_rows, int src_on_device);
/* yref / yref_e of every instance from its policy, then iter++ (tracking) or the switch to position hold. */
int cfnmpc_batch_update_reference(cfnmpc_batch *h);
/* "motors", "euler", "twist" from the current solution; motors_from_u1 != 0 selects u_1 (the FIXED_U0 variant). */
int cfnmpc_batch_commands(cfnmpc_batch *h, int motors_from_u1);
/* One control tick = update_reference + one RTI step + commands (three launches + the solve kernel). */
int cfnmpc_batch_tick(cfnmpc_batch *h, int motors_from_u1);
/* Simulated plant for closed-loop studies: x0 <- ERK4(x0, u_applied, dt) in n_steps steps with the model of the
 * OCP; u_applied = u_0 of the current solution, or the int32 "motors" of the last cfnmpc_batch_commands when
 * truncated_motors != 0 (what the real vehicle and the estimator see, acados_estimator.cpp:463-471). */
int cfnmpc_batch_plant_step(cfnmpc_batch *h, double dt, int n_steps, int truncated_motors);

/* ---- batched state predictor ---------------------------------------------------------------
 * Reference: the estimator node's delay compensation, acados_estimator.cpp:573-593:
 *   sim_in_set(.., "T", &delay); sim_in_set(.., "x", x0); sim_in_set(.., "u", u0);
 *   crazyflie_acados_sim_solve(); sim_out_get(.., "xn", xn);
 * against acados/interfaces/acados_c/sim_interface.h:96-127 and the generated acados_sim_solver_crazyflie.h
 * (c_templates_tera/acados_sim_solver.in.h:81-95).  Same integrator as the OCP (ERK, 4 stages). */
typedef struct cfnmpc_sim cfnmpc_sim;
int cfnmpc_sim_create(int batch, int device, cfnmpc_sim **out);
int cfnmpc_sim_destroy(cfnmpc_sim *s);
int cfnmpc_sim_set_stream(cfnmpc_sim *s, void *cuda_stream);
/* options: "num_steps" (default 1), "sens_forw" (0/1, default 0; the estimator only reads xn), "num_stages" (4) */
int cfnmpc_sim_opts_set(cfnmpc_sim *s, const char *field, int value);
/* fields: "x" double [B][13], "u" double [B][4], "T" double [1] (all instances) or "T_batch" double [B] */
int cfnmpc_sim_set(cfnmpc_sim *s, const char *field, const void *src, int src_on_device);
int cfnmpc_sim_solve(cfnmpc_sim *s);
/* fields: "xn" double [B][13]; "S_forw" double [B][13*17] column-major 13 x 17, columns [x | u] (sim_out "S_forw") */
int cfnmpc_sim_get(cfnmpc_sim *s, const char *field, void *dst, int dst_on_device);
int cfnmpc_sim_launches(cfnmpc_sim *s, long long *n);

/* Device pointer of a batch array ("x0","yref","yref_e","x","u","status","qp_iter"), for
 * callers that produce inputs / consume outputs on the GPU without staging copies. */
int cfnmpc_batch_device_ptr(cfnmpc_batch *h, const char *field, void **ptr);

/* Integer properties: "batch","N","n_slots","sm_count","warps_per_block","blocks_per_sm",
 * "regs_per_thread","smem_per_block" (fused kernel), "two_kernels", "feedback_regs_per_thread",
 * "feedback_blocks_per_sm", "feedback_grid", "preparation_grid", "prepared_bytes", "scratch_bytes", "launches" (kernels launched so far). */
int cfnmpc_batch_info(cfnmpc_batch *h, const char *what, long long *value);
/* Device time of the last cfnmpc_batch_solve in ms (CUDA events on the stream); syncs. */
int cfnmpc_batch_last_solve_ms(cfnmpc_batch *h, double *ms);
/* The same split by launch when the step ran as two kernels (the default on a uniform grid: preparation kernel, then
 * feedback kernel): ms2[0] = preparation, ms2[1] = feedback; a fused step reports {0, total}. */
int cfnmpc_batch_last_phase_ms(cfnmpc_batch *h, double *ms2);

/* Test hook: copy the scratch slot that solved instance 0 when batch == 1 (QP data,
 * factors, IPM vectors) and limit the IPM iteration count; see tests/test_gpu_parity.py. */
/* offsets12 = {slot size, stage-block stride, b_m, b_lu, b_px, r_ux, r_pi, r_rq, r_b, r_resg, r_dux, r_d} in doubles
 * (layout of a stage block: crazyflie_nmpc_b200/csrc/cf_rti_warp.h) */
int cfnmpc_debug_scratch(cfnmpc_batch *h, double *dst, size_t max_doubles, size_t *n_doubles, long long *offsets12);
int cfnmpc_debug_max_ipm_iter(cfnmpc_batch *h, int max_iter);
/* Profiling aid: the first call switches per-pass cycle counters on; later calls copy out and reset them:
 * cycles_calls12[2*p] = warp cycles summed over all warps, [2*p+1] = calls, p = linearisation, residual+factorisation
 * sweep, forward sweep, rhs-only backward sweep, mu_aff, primal update. */
int cfnmpc_debug_pass_cycles(cfnmpc_batch *h, unsigned long long *cycles_calls12);

const char *cfnmpc_last_error(void);
const char *cfnmpc_version(void);
/* Sizes this library was generated for (nx, nu) and the horizon / final time of its OCP description; any pointer may be
 * NULL.  libcfnmpc.so is the Crazyflie OCP (13, 4, 50, 0.75).  The same kernel sources compiled against another generated
 * description (tools/gen_spec.py --model <name>, crazyflie_nmpc_b200/build.py) give libcfnmpc_<name>.so, which exports the
 * core of this header -- cfnmpc_batch_create / destroy / set / set_option / solve / prepare / feedback / sync / get /
 * last_solve_ms / info, cfnmpc_last_error, cfnmpc_version, cfnmpc_model_dims -- with array shapes following its nx, nu
 * (csrc/cfnmpc_generic.cu). */
int cfnmpc_model_dims(int *nx, int *nu, int *N, double *Tf);

/* ------------------------------------------------------------------ several GPUs behind one handle
 * The batch dimension sharded over the GPUs of one node: shard i owns the contiguous instances
 * [first_i, first_i + count_i), count = B / n_dev (+1 for the first B mod n_dev), has its own cfnmpc_batch and stream on
 * devices[i], and every call below runs the shards concurrently, one host thread per device.  There is no exchange
 * between shards.  All pointers are HOST pointers covering the whole batch ([B][...], instance-major); solver-wide
 * fields ("W", "lbu", "bounds_stage", "time_steps", ...) are given once and applied to every shard.  A device may be
 * listed more than once (several shards on one GPU).  For device-resident data use the per-shard handle. */
typedef struct cfnmpc_multi cfnmpc_multi;
int cfnmpc_multi_create(int batch, int N, double Ts, int n_dev, const int *devices, cfnmpc_multi **out);
int cfnmpc_multi_destroy(cfnmpc_multi *m);
int cfnmpc_multi_num_shards(cfnmpc_multi *m);
/* the single-device handle, device, first instance and instance count of shard i (any output may be NULL) */
int cfnmpc_multi_shard(cfnmpc_multi *m, int i, cfnmpc_batch **h, int *device, int *first, int *count);
int cfnmpc_multi_set(cfnmpc_multi *m, const char *field, const void *host_src);   /* fields of cfnmpc_batch_set; returns when copied */
int cfnmpc_multi_set_option(cfnmpc_multi *m, const char *option, int value);
int cfnmpc_multi_set_trajectory(cfnmpc_multi *m, const double *table, int n_rows);
int cfnmpc_multi_solve(cfnmpc_multi *m, int n_rti);                                 /* asynchronous on every device */
int cfnmpc_multi_solve_from_host(cfnmpc_multi *m, const double *x0, const double *yref, const double *yref_e, int n_chunks);
int cfnmpc_multi_tick(cfnmpc_multi *m, int motors_from_u1);
int cfnmpc_multi_sync(cfnmpc_multi *m);
int cfnmpc_multi_get(cfnmpc_multi *m, const char *field, int stage, void *host_dst);   /* per-instance fields of cfnmpc_batch_get */
int cfnmpc_multi_last_solve_ms(cfnmpc_multi *m, double *ms);                        /* slowest shard */
const char *cfnmpc_multi_last_error(void);

/* Measured fp64 FMA throughput of the device in TFLOP/s (a dependency-free DFMA stream on every SM): the denominator of
 * the secondary, FLOP-based roofline figure of bench.py.  No reference counterpart. */
int cfnmpc_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* CFNMPC_H */
