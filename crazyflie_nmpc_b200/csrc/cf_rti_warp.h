// One real-time-iteration SQP step of the Crazyflie OCP executed by ONE WARP for ONE
// problem instance.  Everything the reference does inside acados_solve() for this OCP:
//
//   preparation   ERK4 + forward sensitivities, Gauss-Newton gradient, bound residuals, initial IPM variables
//                 (acados/acados/ocp_nlp/ocp_nlp_sqp_rti.c:495-542,
//                  acados/acados/sim/sim_erk_integrator.c:658-731,
//                  acados/acados/ocp_nlp/ocp_nlp_cost_ls.c:810-916,
//                  acados/acados/ocp_nlp/ocp_nlp_constraints_bgh.c:1613-1648)
//   feedback      x0 elimination, Mehrotra predictor-corrector IPM on the stage-wise QP with a
//                 Riccati factorisation, primal update
//                 (ocp_nlp_sqp_rti.c:545-683, external/hpipm/ocp_qp/x_ocp_qp_red.c:268-455,
//                  x_ocp_qp_ipm.c:1442-1774,1943-2759, x_ocp_qp_kkt.c:401-762,1108-1292,
//                  x_ocp_qp_res.c:334-637, ipm_core/x_core_qp_ipm_aux.c:36-457)
//
// Mapping.  Stage variables are [u(4); x(13)] -> lanes 0..16; lane 17 carries the extra
// "gradient / b" row of the (nv+1) x nv Riccati blocks.  Lane r owns ROW r of the stage matrices
// (the linearisation produces [B';A';b'] as 18 x 13; the feedback program keeps the 14 rows that are not unit vectors,
// see "Free states" below).  Per IPM iteration the warp runs four sweeps over the stages
// (residual_factorize, forward, backward_rhs, forward); each sweep stages the part of a stage block it
// needs in shared memory with ONE TMA bulk copy issued one stage ahead of the arithmetic (double
// buffered, mbarrier-tracked).  Inside a sweep the code is straight-line and branch-free (clamped
// indices, predicated stores); matrix products of the factorisation are fp64 tensor-core tiles,
// GEMV-shaped work reads shared-memory rows / columns with 128-bit loads.  Lanes 18..31 help only
// in element-wise passes and as tensor-core fragment holders.  DESIGN.md section 2 has the full story,
// profiles/README.md the measurements behind each choice.
//
// All stages use the uniform nv = 17 layout: stage 0 keeps 13 decoupled dummy x-variables
// (its A-rows are zeroed after the x0 elimination folded A0*xbar into b0) and stage N keeps 4
// decoupled dummy inputs; both stay exactly zero and cost 2/51 of the work.
//
// The per-instance working set (51 stage blocks of 556 doubles = 227 KB) does not fit on chip; it
// lives in a per-warp scratch slot in global memory (L2/HBM).
//
// Free states.  The dynamics do not depend on the first CF_NF states (CF_SPEC_NFREE, derived by tools/gen_spec.py from
// df/dx: the position of the Crazyflie), so the forward sensitivities of those states stay exactly [I;0] through every
// RK stage (sim_erk_integrator.c:658-731 applied to a zero Jacobian column) and their rows of [B';A'] are unit vectors
// in every stage.  The feedback program neither stores nor multiplies them: the scratch slot keeps the CF_CR = 14 real
// rows (inputs, then states CF_NF..12) with b_k as a vector of its own, the factorisation runs on two instead of three
// 8-row tensor-core tiles (34 instead of 54 DMMA per stage), and the unit rows enter as copies (W rows = P rows, columns
// of S = columns of W).  Stage 0, whose state rows the x0 elimination drops, treats them as zero where it matters.
#pragma once
#include "cf_model.h"

#ifndef CF_SPLIT_FWD
#define CF_SPLIT_FWD 0   // 1: separate copies of the forward sweep for predictor (no multiplier step) and corrector
#endif
#ifndef CF_CHK_IN_UNIFORM
#define CF_CHK_IN_UNIFORM 0   // 1: the uniform-grid (benchmarked) kernels also carry the lin_res_check code
#endif
#ifndef CF_KSPLIT_LXU
#define CF_KSPLIT_LXU 1   // Lxu' dx of the forward substitution split over the lane groups
#endif

// ------------------------------------------------------------------ sizes / layout
#define CF_MROWS (CF_NV + 1)              // rows of [B';A';res_b'] held per stage (18)
#define CF_MSZ (CF_MROWS * CF_NX)         // 234 doubles, element (r,c) at c*18 + r: the form the preparation produces
#define CF_LU (CF_MROWS * CF_NU)          // factor, input columns: 18 x 4, (r,j) at r*4 + j
#define CF_XP ((CF_NX + 1) & ~1)          // a state vector padded to an even number of doubles (14)
#define CF_UP ((CF_NU + 1) & ~1)          // an input vector padded likewise (4)
#define CF_TRI_NX ((CF_NX * (CF_NX + 1)) / 2)   // packed lower triangle of an nx x nx matrix (91)
// compact [B';A'] of the scratch slot (see "Free states" above)
#define CF_NF CF_SPEC_NFREE
#define CF_CR (CF_NV - CF_NF)             // stored rows (14): row m = input m (m < nu), state m - nu + CF_NF otherwise
#define CF_CST ((CF_CR + 1) & ~1)         // column stride (14): element (m,c) at c*14 + m
#define CF_CMSZ (CF_CST * CF_NX)          // 182 doubles
#define CF_PC0 (CF_NU - CF_NF)            // shared-memory array of P: stored state i sits at column i + 1 (its tile column),
#define CF_PCF 16                         //   free state i at column 16 + i (an aligned pair + one), column 19 stays zero
#define CF_PCOL(i) ((i) < CF_NF ? CF_PCF + (i) : (i) + CF_PC0)
#define CF_LUST 6                         // row stride of the 18 x 4 block while it is factorised in shared memory: 128-bit
                                          //   row loads without bank conflicts, column stores 2-way instead of 5-way
#define CF_PST CF_XP                      // row stride of the cost-to-go Hessian once expanded in shared memory (14)
#define CF_LX ((CF_TRI_NX + 1) & ~1)      // cost-to-go Hessian P of the state block in HBM: packed lower triangle (91) + pad
// One contiguous block per stage in the scratch slot.  The field order makes whatever a sweep needs of a stage ONE
// contiguous, 16-byte aligned range = one TMA bulk copy, issued one stage ahead of the arithmetic, with as little as
// possible that the sweep does not need (the residual sweep carries 40 unused doubles, the other two 4):
//   residual+factorisation sweep [0, B_RD)      rhs-only backward sweep [R_BKP, B_BWE)      forward sweep [R_LAM, B_PX | CF_SB)
// 17-vectors are padded to 18, 13-vectors to 14; bound fields are [lb(4) | ub(4)].
//   R_UX  ux_k            R_PI  pi_{k-1} (multiplier of the dynamics ENTERING stage k)   R_DPI  its step
//   R_RQ  gradient        R_D   bound data [lb - u ; u - ub]        R_B  b_k (linearisation, never changes inside the IPM)
//   R_DUX  x-part in: p_k left by a backward sweep; out: the step dux_k
//   R_BKP lam*t of the iterate (res_m backup)          R_PB   P_{k+1} res_b (cached for the rhs-only sweeps)
//   R_RESG stationarity residual (the right-hand side of the rhs-only sweeps)
//   R_DLAM, R_DT steps    R_LAM, R_T multipliers / slacks       R_LU4  l_u left by a backward sweep for the forward sweep
//   B_M    [B';A'] in compact form: the CF_CR = 14 rows that are not unit vectors, element (m,c) at c*14 + m
//   R_RESD bound residual
//   B_LU   factor of the 4 input columns (18 x 4), INVERSE pivots on the diagonal (like BLASFEO's dA)
//   R_RESM complementarity rhs of the next solve       R_RESB dynamics residual
//   B_PX   packed lower triangle of P_{k+1} (what the forward sweep of stage k multiplies with, after expanding it to
//          full symmetric rows in shared memory; written by the factorisation of stage k+1)
enum { R_UX = 0, R_PI = R_UX + CF_MROWS, R_DPI = R_PI + CF_XP, R_RQ = R_DPI + CF_XP, R_D = R_RQ + CF_MROWS, R_B = R_D + 2 * CF_NU,
       R_DUX = R_B + CF_XP, R_BKP = R_DUX + CF_MROWS, R_PB = R_BKP + 2 * CF_NU, R_RESG = R_PB + CF_XP, R_DLAM = R_RESG + CF_MROWS,
       R_DT = R_DLAM + 2 * CF_NU, R_LAM = R_DT + 2 * CF_NU, R_T = R_LAM + 2 * CF_NU, R_LU4 = R_T + 2 * CF_NU, B_M = R_LU4 + CF_UP,
       B_RD = B_M + CF_CMSZ, R_RESD = B_RD, B_LU = R_RESD + 2 * CF_NU, B_BWE = B_LU + CF_LU, R_RESM = B_BWE,
       R_RESB = R_RESM + 2 * CF_NU, B_PX = R_RESB + CF_XP, CF_SB = B_PX + CF_LX };   // 556 doubles per stage (nx = 13, nu = 4)
static_assert(B_M % 2 == 0 && B_RD % 2 == 0 && B_LU % 2 == 0 && B_BWE % 2 == 0 && B_PX % 2 == 0 && CF_SB % 2 == 0 && R_BKP % 2 == 0 &&
              R_LAM % 2 == 0, "16-byte alignment of TMA ranges");
#if CF_CRAZYFLIE
static_assert(R_PI == 18 && R_DPI == 32 && R_RQ == 46 && R_D == 64 && R_B == 72 && R_DUX == 86 && R_BKP == 104 && R_PB == 112 &&
              R_RESG == 126 && R_DLAM == 144 && R_DT == 152 && R_LAM == 160 && R_T == 168 && R_LU4 == 176 && B_M == 180 && B_RD == 362 &&
              B_LU == 370 && B_BWE == 442 && R_RESB == 450 && B_PX == 464 && CF_SB == 556, "stage block layout of the tuned program");
static_assert(CF_NF == 3 && CF_CR == 14 && CF_CST == 14 && CF_PC0 == 1, "compact [B';A'] of the tuned program");
#endif

// HPIPM arguments in effect for the reference configuration (BALANCE mode + acados
// overrides): acados/acados/ocp_qp/ocp_qp_hpipm.c:96-108, x_ocp_qp_ipm.c:133-161
#define CF_RES_G_MAX 1e-6
#define CF_RES_B_MAX 1e-8
#define CF_RES_D_MAX 1e-8
#define CF_RES_M_MAX 1e-8
#define CF_ALPHA_MIN 1e-8
#define CF_MU0 1.0
#define CF_REG_PRIM 1e-15
#define CF_LAM_MIN 1e-16
#define CF_T_MIN 1e-16
#define CF_TAU_MIN 1e-16
#define CF_ITER_MAX 50
#define CF_THR0 0.1

// status codes: acados/acados/utils/types.h:75-83
#define CF_ACADOS_SUCCESS 0
#define CF_ACADOS_QP_FAILURE 4
// flag bits OR-ed into the per-instance `flags` output (not part of the reference API)
#define CF_FLAG_LIN_RES_FACT 1   // reference would have switched to the LQ factorisation (x_ocp_qp_ipm.c:2029-2059)
#define CF_FLAG_LIN_RES_CORR 2   // reference would have run iterative refinement (:2311-2318)
#define CF_FLAG_INPUT_LATE 4     // host-fed tick: the inputs of this instance had not arrived after 5 s (solved with stale data)
// always on (no option needed, a handful of instructions per stage):
#define CF_FLAG_BAD_PIVOT 8      // a non-positive pivot was replaced by 0 in the Riccati factorisation (BLASFEO's rule, silent there)
#define CF_FLAG_NONFINITE 16     // the step length or the duality measure left the finite range (HPIPM status 3 follows)
#define CF_FLAG_ITREF 32         // lin_res_check = 2: the corrector step was iteratively refined (x_ocp_qp_ipm.c:2275-2366)
#define CF_FLAG_ITREF_LEFT 64    // ... and two rounds left a residual above the reference's tolerances (it continues, as HPIPM does)

struct CfParams
{
    double Wdiag[CF_NY];   // stage weights, cost order y = [x;u] (generate_c_code.py:61-84)
    double WNdiag[CF_NX];  // terminal weights (:113)
    double lbu[CF_NU], ubu[CF_NU];    // input box, stages 1..N-1
    double lbu0[CF_NU], ubu0[CF_NU];  // input box of stage 0 (the node's FIXED_U0 branch pins it, acados_mpc.cpp:604-608)
    double Ts;
    int N;
    int max_ipm_iter;      // CF_ITER_MAX unless a test truncates the loop
    int lin_res_check;     // != 0: evaluate the linear-system residuals of every solve (sets the CF_FLAG_LIN_RES_* bits);
                           // 2: and refine the corrector step iteratively where the reference would (itref_corr_max = 2);
                           // 4 (tests): 2 with every corrector solve deliberately 10 % off in du, so that refinement runs
    int pad_;
};
#define CF_PAR_DOUBLES (CF_NY + CF_NX + 4 * CF_NU + 3)  // sizeof(CfParams) / 8 (49): the per-warp copy in shared memory
static_assert(sizeof(CfParams) == CF_PAR_DOUBLES * 8, "CfParams layout");

struct CfBatchView
{
    int B;                 // instances this launch solves: [first, first + B)
    int first;
    const int *ready;      // null, or a device counter: only instances < *ready have their inputs in place (host-fed ticks
                           // whose upload overlaps the solve, cfnmpc_batch_solve_from_host)
    const double *x0;      // [B][13]
    const double *yref;    // [B][N][17]
    const double *yref_e;  // [B][13]
    double *x;             // [B][N+1][13]  iterate, in/out
    double *u;             // [B][N][4]
    int *status;           // [B] acados status
    int *qp_iter;          // [B]
    int *qp_status;        // [B] HPIPM status 0/1/2/3
    int *flags;            // [B]
    double *res;           // [B][4] final residual inf-norms (may be null)
    double *scratch;       // n_slots * scratch_stride doubles
    long scratch_stride;
    int *counter;          // work queue
    // optional per-instance overrides of the solver-wide CfParams (null = not given), SET_WEIGHTS / FIXED_U0 of the
    // node with one value set per vehicle: W [B][17], W_e [B][13], lbu/ubu [B][4] (stages 1..N-1), lbu0/ubu0 [B][4]
    const double *W_b, *WN_b, *lbu_b, *ubu_b, *lbu0_b, *ubu0_b;
    // optional profiling counters (null = off): per-pass warp cycles and call counts, see cfnmpc_debug_pass_cycles
    unsigned long long *prof;
    // non-uniform shooting grid (crazyflie_acados_create_with_discretization / _update_time_steps,
    // c_templates_tera/acados_solver.in.c:133-153): dts[N] = length of every interval = its cost scaling; only read by
    // the kernel variants compiled with VDT (null otherwise: CfParams::Ts applies to every interval)
    const double *dts;
    // split real-time iteration (rti_phase 1 / 2, ocp_nlp_sqp_rti.c:495-542,545-683): what the preparation phase leaves
    // for the feedback phase, per INSTANCE: N stage records [ [B';A';b'] (234) | gradient (18) ] + the terminal gradient
    double *prep;
    long prep_stride;
    // input box per STAGE, [N][8] = lbu(4) | ubu(4) of stage k (ocp_nlp_constraints_model_set addresses one stage at a
    // time, ocp_nlp_constraints_bgh.c:653-674); null = the stage-0 / path boxes of CfParams.  Takes precedence over them
    // and over the per-instance arrays.
    const double *bnd_stage;
    // cost weights per STAGE, [N+1][17]: row k < N = diagonal of W_k in cost order y = [x;u], row N = diagonal of W_e (13
    // used): ocp_nlp_cost_model_set addresses one stage at a time (ocp_nlp_cost_ls.c:301-331); null = the weights of
    // CfParams / the per-instance arrays.  Only read by the general kernel variants (VDT) and the condensed feedback.
    const double *W_stage;
    // optional per-instance multiplier output (null = off; option "multipliers"): what ocp_nlp_out_get "pi" / "lam" / "t"
    // hand out after a step (acados_c/ocp_nlp_interface.c:576-590; full-step duals, ocp_nlp_common.c:2917-2925).  Layout
    // per instance (mult_stride doubles): pi [N][13] | lam [N][8] = lower(4) | upper(4) of the input box | t [N][8] |
    // lam_x0 [13]: SIGNED multiplier of the eliminated constraint x_0 = x0 (>= 0: lower-bound multiplier, < 0: minus the
    // upper-bound one, x_ocp_qp_red.c:820-840) | A_0 [13][13] scratch (the stage-0 state rows, saved before the
    // elimination drops them; element (c, r) = d x1_c / d x0_r at c*13 + r)
    double *mult;
    long mult_stride;
    // full (non-diagonal) weight matrices per stage (ocp_nlp_cost_ls.c:301-331 accepts any SPD W), null = diagonal weights:
    // [(N+1)][2][17*17] in stage-variable order [u;x] -- first (Cyt W_chol)(Cyt W_chol)' (what the reference's Hessian is,
    // :743-772), then W itself (gradient, :883-912); row N carries W_e in its state block.  Read by the condensed feedback
    // program with block size 1 (cf_pcond_warp.h), which has the dense stage Hessian the uncondensed program lacks.
    const double *W_dense;
};
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    long cf_mult_stride(int N) { return ((long) N * 29 + 13 + 169 + 1) & ~1L; }
// layout of the prepared linearisation of one instance (doubles)
#define CF_PREP_STAGE (CF_MSZ + CF_MROWS)
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    long cf_prep_stride(int N) { return ((long) N * CF_PREP_STAGE + CF_MROWS + 15) & ~15L; }
enum { CF_PROF_LIN = 0, CF_PROF_RF, CF_PROF_FWD, CF_PROF_BWD, CF_PROF_MUAFF, CF_PROF_UPDATE, CF_PROF_N };

// offsets (in doubles) of the arrays inside one scratch slot; every block that the TMA engine
// touches (M, LF) starts on a 16-byte boundary
struct CfScratchLayout
{
    long blk, total;
};
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    CfScratchLayout
    cf_scratch_layout(int N)
{
    CfScratchLayout s;
    s.blk = 0;
    s.total = ((long) (N + 1) * CF_SB + 15) & ~15L;  // keep every slot 128-byte aligned
    return s;
}

// per-warp shared memory (doubles); every region starts on a 16-byte boundary
#define CF_MAX2(a, b) ((a) > (b) ? (a) : (b))
// sweeps: staged range of a stage block, double buffered (468 doubles for nx = 13, nu = 4: the two [B';A';b'] of the linearisation)
#define CF_WST CF_ALST                                 // row stride of W (a swizzled stride-24 layout without bank conflicts in
                                                       //   stores and loads was measured: fewer wavefronts, not faster)
#define CF_SM_WLU (16 * CF_WST + CF_MROWS * CF_LUST)   // factorisation: W (16 rows) and the 18 x 4 block behind it
#define CF_SM_BUFSZ CF_MAX2(CF_MAX2(B_RD, CF_SB - R_LAM), CF_MAX2(CF_MAX2(B_BWE - R_BKP, CF_SM_WLU), 2 * CF_MSZ))
#define CF_SM_BUF0 0
#define CF_SM_BUF1 CF_SM_BUFSZ
#define CF_SM_MS0 0                            // linearisation: [B';A';b'] staging, double buffered
#define CF_SM_MS1 CF_MSZ
#define CF_ALST 20                             // row stride 20: conflict-free fp64 tensor-core fragment loads
#define CF_SM_P (2 * CF_SM_BUFSZ)              // factorisation: P_{k+1}, 13 x 20 (the W / input-column block, 18 x 20,
                                               //   overlays the staged block of the stage being factorised);
                                               //   forward sweep: P_{k+1} expanded to full rows, 13 x 14
#define CF_SM_V0 (CF_SM_P + CF_NX * CF_ALST)   // four 20-double broadcast vectors
#define CF_SM_V1 (CF_SM_V0 + 20)
#define CF_SM_V2 (CF_SM_V1 + 20)
#define CF_SM_V3 (CF_SM_V2 + 20)
#define CF_SM_BAR (CF_SM_V3 + 20)              // two mbarriers
#define CF_SM_PAR (CF_SM_BAR + 4)              // this instance's CfParams (solver-wide values + per-instance overrides)
#define CF_SM_DOUBLES (CF_SM_PAR + ((CF_PAR_DOUBLES + 1) & ~1))  // 1330 doubles = 10640 bytes per warp
static_assert(CF_SM_DOUBLES % 2 == 0, "every warp's shared-memory slice must start on a 16-byte boundary");
static_assert(B_RD <= CF_SM_BUFSZ && CF_SB - R_LAM <= CF_SM_BUFSZ && B_BWE - R_BKP <= CF_SM_BUFSZ && CF_SM_WLU <= CF_SM_BUFSZ,
              "staging buffers");
static_assert(!CF_CRAZYFLIE || CF_SM_BUFSZ == 468, "shared-memory layout of the tuned program");

CF_DEV int cf_tri(int i) { return (i * (i + 1)) >> 1; }
// PH: 0 = preparation + feedback in one go (what acados_solve() does with rti_phase 0), 1 = preparation only,
//     2 = feedback only (ocp_nlp_sqp_rti.c:1213-1237).  VDT: per-interval time steps CfBatchView::dts instead of the
//     uniform CfParams::Ts.  The benchmarked kernel is <0, false>; the other variants cost it nothing.
enum { CF_PH_BOTH = 0, CF_PH_PREPARATION = 1, CF_PH_FEEDBACK = 2 };
template <int PH, bool VDT>
struct CfWarpT
{
    // ---- per-warp context
    const CfParams *P;    // this instance's weights and bounds (per-warp copy in shared memory)
    const CfParams *PG;   // solver-wide scalars Ts, N, max_ipm_iter (kernel parameter, constant bank)
    int lane, N;
    double *sm;  // per-warp shared memory
    uint64_t *bar;
    unsigned par;  // phase parity of the two mbarriers
    // scratch arrays
    double *SLOT;
    double *PREP;      // this instance's prepared linearisation (split phases only)
    const double *DT;  // per-interval time steps (VDT only)
    const double *BST; // per-stage input boxes (null: the boxes of P)
    const double *WST; // per-stage weights (general variants only; null: the weights of P)
    double *A0S;       // where the stage-0 state rows go before the x0 elimination drops them (null: not wanted)
    // lane constants
    double Hs, HN;     // Hessian diagonal for this lane's variable (stage / terminal)
    double W2;         // (sqrt(w))^2 of this lane's stage weight: the stage Hessian is dt_k * W2
    // IPM scalars (warp-uniform)
    double mu, alpha, mu_aff, sigma, pm_max;
    double nrm[4];     // inf-norms of res_g, res_b, res_d, res_m
    double lin[4];     // inf-norms of the linear-system residual of the last solve
    int flags;

    CF_MEM void bind(const CfParams *P_, const CfParams *PG_, double *slot, double *sm_, double *prep_, const double *dts_)
    {
        P = P_; PG = PG_; N = PG_->N; sm = sm_; lane = cf_lane();
        PREP = prep_; DT = dts_; BST = nullptr; WST = nullptr; A0S = nullptr;
        bar = reinterpret_cast<uint64_t *>(sm_ + CF_SM_BAR);
        par = 0;
        CfScratchLayout s = cf_scratch_layout(N);
        SLOT = slot + s.blk;
        // hess = scaling * (sqrt(W))^2 : ocp_nlp_cost_ls.c:739-772 (terminal scaling stays 1.0, :265)
        double w = 1.0, wN = 1.0;
        if (lane < CF_NU) w = P->Wdiag[CF_NX + lane];
        else if (lane < CF_NV) { w = P->Wdiag[lane - CF_NU]; wN = P->WNdiag[lane - CF_NU]; }
        double r = sqrt(w), rN = sqrt(wN);
        W2 = r * r;
        Hs = PG->Ts * W2;
        HN = (lane < CF_NU) ? Hs : (rN * rN);
    }
    // weights that differ from stage to stage (general variants): the terminal Hessian comes from row N of the table
    CF_MEM void set_stage_weights(const double *wst)
    {
        if constexpr (VDT) {
            WST = wst;
            if (WST && lane >= CF_NU && lane < CF_NV) { const double rN = sqrt(WST[N * CF_NY + lane - CF_NU]); HN = rN * rN; }
        }
    }
    // weight of cost component idx (cost order y = [x;u]) at stage k <= N
    CF_MEM double wgt(int k, int idx) const
    {
        if constexpr (VDT) { if (WST) return WST[k * CF_NY + idx]; }
        return k < N ? P->Wdiag[idx] : P->WNdiag[idx];
    }
    // length of shooting interval k = scaling of its cost term (ocp_nlp_in "Ts" / cost "scaling")
    CF_MEM double dt(int k) const
    {
        if constexpr (VDT) return DT[k];
        else return PG->Ts;
    }
    // Hessian diagonal of this lane's variable at stage k < N
    CF_MEM double hess(int k) const
    {
        if constexpr (VDT) {
            if (WST) {
                const double r = sqrt(WST[k * CF_NY + (lane < CF_NU ? CF_NX + lane : (lane < CF_NV ? lane - CF_NU : 0))]);
                return DT[k] * (r * r);
            }
            return DT[k] * W2;
        } else return Hs;
    }

    // ---- TMA staging: one mbarrier per buffer; lane 0 issues, every lane waits
    CF_MEM void pass_begin()
    {
        cf_syncwarp();                            // all generic stores of the previous pass are ordered ...
        if (lane == 0) cf_fence_proxy_async();    // ... before the bulk (async-proxy) reads of this pass
    }
    CF_MEM double *blk(int k) const { return SLOT + (long) k * CF_SB; }
    CF_MEM double *rec(int k) const { return blk(k); }   // record fields are addressed from the block start
    CF_MEM double *buf(int bf) const { return sm + (bf ? CF_SM_BUF1 : CF_SM_BUF0); }
    // stage doubles [start, start+len) of stage block k into buffer `bf` (one bulk copy, lane 0 issues)
    CF_MEM void fetch_to(double *dst, int bf, int k, int start, int len)
    {
        if (lane == 0) {
            cf_bulk_expect(bar + bf, len * 8);
            cf_bulk_g2s_raw(dst, blk(k) + start, len * 8, bar + bf);
        }
    }
    CF_MEM void fetch(int bf, int k, int start, int len) { fetch_to(buf(bf), bf, k, start, len); }
    CF_MEM void wait(int bf)
    {
        cf_bulk_wait(bar + bf, (par >> bf) & 1u);
        par ^= 1u << bf;
    }

    // =============================================================== preparation
    // Multiple shooting makes the N intervals independent, so the NOMINAL integration is done first, one interval per
    // lane (two rounds of 32 lanes at N = 50) instead of redundantly on every lane of every stage: classic RK4
    // (tableau sim_collocation_utils.c:611-640, operation order of sim_erk_integrator.c:658-731).  Leaves, per interval,
    // the four RK stage states, b_k = phi(x_k,u_k) - x_{k+1} (ocp_nlp_dynamics_cont.c:822-823) and u_k in the (still
    // unused) factor area of the stage block, from where the sensitivity pass stages them by TMA:
    //   NOM = [ xs_0 (13+1) | xs_1 | xs_2 | xs_3 | b (13+1) | u (4) ]  at B_LU
#define CF_NOM (5 * CF_XP + CF_UP)    // 74
    // where the NOM record sits in the stage block: the factor area, or -- preparation-only builds of small models whose
    // factor area is too short -- the start of the block (nothing else of the block is touched by a preparation phase)
#define CF_NOM_OFF ((CF_NOM <= CF_SB - B_LU) ? B_LU : 0)
    static_assert(CF_NOM <= CF_SB - CF_NOM_OFF && (CF_NOM_OFF == B_LU || PH == CF_PH_PREPARATION || !CF_CRAZYFLIE), "NOM record");
    CF_MEM void nominal_pass(const double *xg, const double *ug)
    {
        CF_NOUNROLL
        for (int k = lane; k < N; k += 32) {
            const double h = dt(k);
            double *nom = blk(k) + CF_NOM_OFF;
            double x[CF_NX], xs[CF_NX], acc[CF_NX], uu[CF_NU];
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) { x[i] = xg[k * CF_NX + i]; xs[i] = x[i]; acc[i] = x[i]; }
            CF_UNROLL
            for (int i = 0; i < CF_NU; i++) uu[i] = ug[k * CF_NU + i];
            CF_UNROLL
            for (int i = 0; i + 1 < CF_NU; i += 2) cf_st2(nom + 5 * CF_XP + i, uu[i], uu[i + 1]);
            if (CF_NU & 1) nom[5 * CF_XP + CF_NU - 1] = uu[CF_NU - 1];
            CF_UNROLL
            for (int s = 0; s < 4; s++) {
                const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
                const double a_next = (s == 2) ? 1.0 : 0.5;
                const double bh = h * bw, ah = a_next * h;
                CF_UNROLL
                for (int i = 0; i + 1 < CF_NX; i += 2) cf_st2(nom + s * CF_XP + i, xs[i], xs[i + 1]);
                if (CF_NX & 1) nom[s * CF_XP + CF_NX - 1] = xs[CF_NX - 1];
                double f[CF_NX];
                cf_ode(xs, uu, f);
                CF_UNROLL
                for (int i = 0; i < CF_NX; i++) {
                    acc[i] += bh * f[i];
                    xs[i] = x[i] + ah * f[i];
                }
            }
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) acc[i] -= xg[(k + 1) * CF_NX + i];
            CF_UNROLL
            for (int i = 0; i + 1 < CF_NX; i += 2) cf_st2(nom + 4 * CF_XP + i, acc[i], acc[i + 1]);
            if (CF_NX & 1) nom[4 * CF_XP + CF_NX - 1] = acc[CF_NX - 1];
        }
        pass_begin();   // the generic stores above are read back by bulk copies
        if (N > 0) fetch_nom(0, 0);
    }
    CF_MEM double *nom_buf(int bf) const { return sm + CF_SM_BUFSZ + bf * ((CF_NOM + 7) & ~7); }   // (sm + 468 + bf * 80)
    CF_MEM void fetch_nom(int bf, int k)
    {
        if (lane == 0) {
            cf_bulk_expect(bar + bf, CF_NOM * 8);
            cf_bulk_g2s_raw(nom_buf(bf), blk(k) + CF_NOM_OFF, CF_NOM * 8, bar + bf);
        }
    }

    // Forward sensitivities of stage k along the stored nominal RK stages; lane c pushes sensitivity column c
    // ([Su(4) | Sx(13)] -> rows of [B';A']).  Writes M_k (rows 0..16 + b row) by bulk store, rq_k, d_k and the initial
    // IPM variables of the stage.
    CF_MEM void linearize_stage(int k, const double *xg, const double *x0g, const double *yrefg)
    {
        double *MS = sm + ((k & 1) ? CF_SM_MS1 : CF_SM_MS0);
        const double h = dt(k);
        const int bf = k & 1;
        // preparation-only: the linearisation goes to the instance's prepared record instead of the warp's scratch slot
        double *mdst = (PH == CF_PH_PREPARATION) ? PREP + (long) k * CF_PREP_STAGE : blk(k) + B_M;
        double *rqdst = (PH == CF_PH_PREPARATION) ? PREP + (long) k * CF_PREP_STAGE + CF_MSZ : rec(k) + R_RQ;
        // this stage's reference (consumed at the end of the stage)
        const double yr_pre = (lane < CF_NU) ? yrefg[k * CF_NY + CF_NX + lane] : ((lane < CF_NV) ? yrefg[k * CF_NY + lane - CF_NU] : 0.0);
        wait(bf);
        cf_syncwarp();   // every lane is done with the other nominal buffer and with MS of stage k-2 ...
        if (k + 1 < N) fetch_nom(bf ^ 1, k + 1);
        if (lane == 0) cf_bulk_s2g_wait_read1();  // ... and the bulk store that last read this MS buffer (stage k-2) is done
        cf_syncwarp();
        const double *NOMS = nom_buf(bf);
        double uu[CF_NU];
        CF_UNROLL
        for (int i = 0; i + 1 < CF_NU; i += 2) { const cf_d2 v = cf_ld2(NOMS + 5 * CF_XP + i); uu[i] = v.x; uu[i + 1] = v.y; }
        if (CF_NU & 1) uu[CF_NU - 1] = NOMS[5 * CF_XP + CF_NU - 1];
        // the accumulated sensitivity column of this lane lives in its row of the staging block (MS[c*18+lane]),
        // not in registers: the RK stage below is register-hungry enough
        const bool col = lane < CF_NV;
        double *Mrow = MS + (col ? lane : CF_NV);
        double Ss[CF_NX];
        CF_UNROLL
        for (int i = 0; i < CF_NX; i++) {
            Ss[i] = (lane - CF_NU == i) ? 1.0 : 0.0;
            if (col) Mrow[i * CF_MROWS] = Ss[i];
        }
        CF_NOUNROLL
        for (int s = 0; s < 4; s++) {
            const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
            const double a_next = (s == 2) ? 1.0 : 0.5;
            const double bh = h * bw, ah = a_next * h;
            double xs[CF_NX];
            CF_UNROLL
            for (int i = 0; i + 1 < CF_NX; i += 2) { const cf_d2 v = cf_ld2(NOMS + s * CF_XP + i); xs[i] = v.x; xs[i + 1] = v.y; }
            if (CF_NX & 1) xs[CF_NX - 1] = NOMS[s * CF_XP + CF_NX - 1];
            double ks[CF_NX];
            cf_jvp_x(xs, uu, Ss, ks);
            if (lane < CF_NU) cf_add_ju_col(xs, uu, lane, ks);
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) {
                if (col) Mrow[i * CF_MROWS] += bh * ks[i];
                Ss[i] = ((lane - CF_NU == i) ? 1.0 : 0.0) + ah * ks[i];
            }
        }
        // row nv (17): b_k
        if (lane < CF_NX) MS[lane * CF_MROWS + CF_NV] = NOMS[4 * CF_XP + lane];
        cf_syncwarp();
        if (PH != CF_PH_PREPARATION && k == 0) eliminate_x0(MS, xg, x0g);   // (the feedback phase does it otherwise)
        // gradient: scaling * W * (y - yref), [u;x] order (ocp_nlp_cost_ls.c:883-912)
        const double uk = (lane < CF_NU) ? NOMS[5 * CF_XP + lane] : 0.0;
        if (lane < CF_NV) {
            double g;
            if (lane < CF_NU) g = (wgt(k, CF_NX + lane) * (uk - yr_pre)) * h;
            else g = (k == 0) ? 0.0 : (wgt(k, lane - CF_NU) * (NOMS[lane - CF_NU] - yr_pre)) * h;
            rqdst[lane] = g;
        }
        if (PH == CF_PH_PREPARATION && lane == CF_NV) rqdst[lane] = 0.0;   // pad of the 18-double gradient record
        if (PH != CF_PH_PREPARATION) stage_bounds_init(k, uk);
        cf_syncwarp();
        if constexpr (PH == CF_PH_PREPARATION) {
            if (lane == 0) cf_bulk_s2g(mdst, MS, CF_MSZ * 8);
        } else {
            // one-launch step: the compact form goes straight to the scratch slot (generic stores; the pass_begin() of the
            // first sweep orders them before its bulk reads)
            CF_NOUNROLL
            for (int e = lane; e < CF_CMSZ; e += 32) {
                const int c = e / CF_CST, m = e - c * CF_CST;
                mdst[e] = (m < CF_CR) ? MS[c * CF_MROWS + (m < CF_NU ? m : m + CF_NF)] : 0.0;
            }
            if (lane < CF_XP) rec(k)[R_B + lane] = lane < CF_NX ? MS[lane * CF_MROWS + CF_NV] : 0.0;   // (pad: a staged zero)
        }
    }

    // x0 elimination on the staged block of stage 0 (x_ocp_qp_red.c:310-330): xbar = lbx - x_0 ; b_0 += A_0 xbar ; drop
    // the A rows
    CF_MEM void eliminate_x0(double *MS, const double *xg, const double *x0g)
    {
        const bool xl = lane >= CF_NU && lane < CF_NV;
        double *Mrow = MS + (lane < CF_NV ? lane : CF_NV);
        const double xbar = xl ? (x0g[lane - CF_NU] - xg[lane - CF_NU]) : 0.0;
        CF_NOUNROLL
        for (int i = 0; i < CF_NX; i++) {
            if (A0S && xl) A0S[i * CF_NX + lane - CF_NU] = Mrow[i * CF_MROWS];   // kept for the x0 multipliers
            const double tot = cf_warp_sum(xl ? Mrow[i * CF_MROWS] * xbar : 0.0);
            if (lane == CF_NV) MS[i * CF_MROWS + CF_NV] = tot + MS[i * CF_MROWS + CF_NV];
            else if (xl) Mrow[i * CF_MROWS] = 0.0;
        }
        cf_syncwarp();
    }

    // bound data and the initial interior-point variables of stage k < N; uk = u_k[lane] on lanes 0..3
    CF_MEM void stage_bounds_init(int k, double uk)
    {
        // bounds (ocp_nlp_constraints_bgh.c:1634-1636): d = [lb - u ; u - ub]
        double v0 = 0.0;
        if (lane < CF_NU) {
            double lb = (k == 0) ? P->lbu0[lane] : P->lbu[lane], ub = (k == 0) ? P->ubu0[lane] : P->ubu[lane];
            if (BST) { lb = BST[k * 2 * CF_NU + lane]; ub = BST[k * 2 * CF_NU + CF_NU + lane]; }
            const double dl = lb - uk;
            const double du = uk - ub;
            rec(k)[R_D + lane] = dl;
            rec(k)[R_D + CF_NU + lane] = du;
            // OCP_QP_INIT_VAR scheme 1 (x_ocp_qp_ipm.c:1491-1530,1636-1769): slacks at ux = 0, pushed 0.1 inside
            double tl = -dl, tu = -du;
            if (tl < CF_THR0) {
                if (tu < CF_THR0) { v0 = 0.5 * (dl - du); tl = CF_THR0; tu = CF_THR0; }
                else { tl = CF_THR0; v0 = dl + CF_THR0; }
            } else if (tu < CF_THR0) { tu = CF_THR0; v0 = -du - CF_THR0; }
            rec(k)[R_T + lane] = tl; rec(k)[R_T + CF_NU + lane] = tu;
            rec(k)[R_LAM + lane] = CF_MU0 / tl; rec(k)[R_LAM + CF_NU + lane] = CF_MU0 / tu;
        }
        init_stage_vectors(k, v0);
    }

    // ux = v (0 unless a bound had to be respected), pi = 0 and zero steps, so that the first residual pass can be an
    // "update with step 0" (one code copy of that sweep)
    CF_MEM void init_stage_vectors(int k, double v)
    {
        double *rk = rec(k);
        if (lane < CF_NV) { rk[R_UX + lane] = v; rk[R_DUX + lane] = 0.0; }
        if (lane < CF_NX) { rk[R_PI + lane] = 0.0; rk[R_DPI + lane] = 0.0; }
        if (lane < 2 * CF_NU) { rk[R_DLAM + lane] = 0.0; rk[R_DT + lane] = 0.0; }
    }

    CF_MEM void terminal_gradient(const double *xg, const double *yref_eg)
    {
        if (PH != CF_PH_PREPARATION) init_stage_vectors(N, 0.0);
        if (lane < CF_NV) {
            double g = 0.0;
            if (lane >= CF_NU) g = wgt(N, lane - CF_NU) * (xg[N * CF_NX + lane - CF_NU] - yref_eg[lane - CF_NU]);
            if (PH == CF_PH_PREPARATION) PREP[(long) N * CF_PREP_STAGE + lane] = g;
            else rec(N)[R_RQ + lane] = g;
        }
        if (lane == 0) cf_bulk_s2g_wait_all();  // every M_k has landed in global memory
    }

#if CF_CRAZYFLIE   // ---- everything below, up to the end of the class, is the tuned program for nx = 13, nu = 4
    // Feedback phase of a split real-time iteration (ocp_nlp_sqp_rti.c:545-683): the linearisation comes from the
    // instance's prepared record; what depends on data that may have changed since the preparation -- the measured
    // state (stage-0 elimination), the bound vectors (ocp_nlp_approximate_qp_vectors_sqp, ocp_nlp_common.c:2258-2292) --
    // is evaluated now, together with the initial interior-point variables.
    CF_MEM void load_prepared(const double *xg, const double *ug, const double *x0g)
    {
        // Four stage records per trip travel prepared store -> shared memory -> scratch slot on the TMA engine (one bulk
        // load of the contiguous records, two bulk stores per stage: [B';A'] and the gradient); the lanes meanwhile
        // evaluate the bound vectors and the initial interior-point variables of those stages, and then bring [B';A';b']
        // into the compact form of the slot (CF_CR): b_k leaves as a vector, the rows that are not unit vectors move up.
        double *ST = sm;
        static_assert(4 * CF_PREP_STAGE <= CF_SM_V0, "staging area of load_prepared");
        pass_begin();   // generic accesses of the previous instance to this shared memory precede the bulk writes
        CF_NOUNROLL
        for (int k0 = 0; k0 < N; k0 += 4) {
            const int n = (N - k0 < 4) ? N - k0 : 4;
            if (lane == 0) {
                cf_bulk_expect(bar, n * CF_PREP_STAGE * 8);
                cf_bulk_g2s_raw(ST, PREP + (long) k0 * CF_PREP_STAGE, n * CF_PREP_STAGE * 8, bar);
            }
            CF_NOUNROLL
            for (int q = 0; q < n; q++) stage_bounds_init(k0 + q, lane < CF_NU ? ug[(k0 + q) * CF_NU + lane] : 0.0);
            wait(0);
            if (k0 == 0) eliminate_x0(ST, xg, x0g);
            cf_syncwarp();
            CF_NOUNROLL
            for (int q = 0; q < n; q++) {
                double *Sq = ST + q * CF_PREP_STAGE;
                if (lane < CF_XP) rec(k0 + q)[R_B + lane] = lane < CF_NX ? Sq[lane * CF_MROWS + CF_NV] : 0.0;   // (pad: a staged zero)
                // in place, 32 elements per trip in increasing order: a destination never lies above a source still to be read
                CF_NOUNROLL
                for (int e0 = 0; e0 < CF_CMSZ; e0 += 32) {
                    const int e = e0 + lane, c = e / CF_CST, m = e - c * CF_CST;
                    const bool on = e < CF_CMSZ;
                    const double v = (on && m < CF_CR) ? Sq[c * CF_MROWS + (m < CF_NU ? m : m + CF_NF)] : 0.0;
                    cf_syncwarp();
                    if (on) Sq[e] = v;
                }
            }
            cf_syncwarp();
            if (lane == 0) {
                CF_NOUNROLL
                for (int q = 0; q < n; q++) {
                    cf_bulk_s2g(blk(k0 + q) + B_M, ST + q * CF_PREP_STAGE, CF_CMSZ * 8);
                    cf_bulk_s2g(rec(k0 + q) + R_RQ, ST + q * CF_PREP_STAGE + CF_MSZ, 18 * 8);
                }
                cf_bulk_s2g_wait_read0();   // the staging area may be refilled
            }
            cf_syncwarp();
        }
        init_stage_vectors(N, 0.0);
        if (lane < CF_NV) rec(N)[R_RQ + lane] = PREP[(long) N * CF_PREP_STAGE + lane];
        if (lane == 0) cf_bulk_s2g_wait_all();   // every block has landed in the slot before the sweeps fetch it
    }

    // =============================================================== IPM pieces
    // lane -> stored row of the compact [B';A'] (inputs 0..3, states CF_NF.. behind them); free states and idle lanes get a
    // harmless row of their own half-warp (the same address as an active lane: a broadcast, no bank conflict)
    CF_MEM bool stored_lane() const { return lane < CF_NU || (lane >= CF_NU + CF_NF && lane < CF_NV); }
    CF_MEM bool free_lane() const { return lane >= CF_NU && lane < CF_NU + CF_NF; }
    CF_MEM int mrow() const { return stored_lane() ? (lane < CF_NU ? lane : lane - CF_NF) : (lane < 16 ? 0 : CF_CR - 1); }
    // lane -> position of its variable in a broadcast vector for the column products: stored rows first, free states behind
    CF_MEM int vidx() const { return lane < CF_NU ? lane : (lane < CF_NU + CF_NF ? CF_CR + lane - CF_NU : lane - CF_NF); }

    // y[c] = sum_{m < 14} M[m][c] v[m] + v[14 + c] (c < CF_NF: the unit row of free state c) + bv[c] (when given) for the
    // state column c of lanes 4..16, the 14 stored rows of a column shared by a lane pair: the state lane takes rows 0..5
    // and the two extra terms, its partner (lanes 20..31 and lane 17) rows 6..13; one exchange combines them.  (The split at
    // row 6 keeps the 128-bit loads of every quarter warp on distinct banks at column stride 14.)  Mk: staged compact
    // [B';A'] (element (m,c) at c*14 + m), v: 20-double vector in the order of vidx().
    // (the lane pairing is a per-lane constant: col_consts() computes it once per sweep and pins it in registers -- ptxas
    // otherwise re-derives it in every stage, 3 % of all executed instructions)
    CF_MEM void col_consts(int &cgo, int &clp) const
    {
        // partner of state lane L: L ^ 16 (lanes 20..31), except lane 16 whose partner is lane 17; lanes 0..3, 18, 19 compute
        // a throw-away duplicate of column 0
        const int lp = lane == 16 ? 17 : (lane == 17 ? 16 : lane ^ 16);
        const bool lo = lane >= CF_NU && lane < CF_NV, hi = lane == 17 || lane >= 20;
        const int cc = lo ? lane - CF_NU : (hi ? lp - CF_NU : 0), r0 = lo ? 0 : 6;
        cgo = cc * CF_CST + r0;   // this lane's part of its column: rows r0.. of column cc
        clp = lp;
        cf_keep(cgo); cf_keep(clp);
    }
    CF_MEM double col_gemv(const double *Mk, const double *v, const double *bv, const int cgo, const int clp) const
    {
        const bool lo = lane >= CF_NU && lane < CF_NV;
        const double *Mc = Mk + cgo, *vc = v + (lo ? 0 : 6);
        double s0 = 0.0, s1 = 0.0;
        CF_UNROLL
        for (int rp = 0; rp < 3; rp++) {   // rows 0..5 | 6..11
            const cf_d2 m2 = cf_ld2(Mc + 2 * rp), v2 = cf_ld2(vc + 2 * rp);
            s0 += m2.x * v2.x;
            s1 += m2.y * v2.y;
        }
        if (lo) {
            const int cc = lane - CF_NU;
            s0 += (cc < CF_NF) ? v[CF_CR + cc] : 0.0;
            s1 += bv ? bv[cc] : 0.0;
        } else {                           // rows 12, 13
            const cf_d2 m3 = cf_ld2(Mc + 6), v3 = cf_ld2(vc + 6);
            s0 += m3.x * v3.x;
            s1 += m3.y * v3.y;
        }
        const double part = s0 + s1;
        return part + cf_shfl(part, clp);
    }

    // One BACKWARD sweep that does, per stage, everything the reference spreads over three passes:
    //   UPDATE_VAR_QP            x_core_qp_ipm_aux.c:220-325   (variables += step length * direction, clipping)
    //   OCP_QP_RES_COMPUTE + _INF_NORM   x_ocp_qp_res.c:334-470,602-637   (residuals of the new iterate, norms, mu)
    //   OCP_QP_FACT_SOLVE_KKT_STEP, backward part   x_ocp_qp_kkt.c:401-762   (Riccati factorisation for the next direction)
    // The residuals of a stage are local (they couple stage k only to ux_{k+1} and pi_{k-1}, pi_k), so they can be
    // evaluated in the order the factorisation runs; the factorisation of the final iterate is wasted (1 of ~7 sweeps)
    // in exchange for one sweep less per interior-point iteration.  `a_raw` is the step length of the direction to
    // apply; the first pass applies 0 to the zeroed directions of the initialisation.
    //
    // Factorisation: HPIPM's classical Riccati recursion (square_root_alg = 0, x_ocp_qp_kkt.c:573-740) instead of the
    // square-root variant the reference selects (:445-528): only the 4 input columns of each stage block are
    // Cholesky-factorised, the state block is kept as the symmetric cost-to-go Hessian P_k.  Same arithmetic cost, but
    // 4 instead of 17 strictly sequential pivot steps per stage and every matrix product is a tensor-core tile product;
    // on this OCP the two recursions agree to 1e-14 (oracle/cfnmpc_oracle.c: cfo_set_classical_riccati,
    // tests/test_oracle_golden.py).  Per stage k < N:
    //   W   = [B';A';res_b']_k P_{k+1}                      GEMM_NT :621   (gradient row: Pb_k = P res_b, then += p_{k+1}')
    //   S   = [H_k + Gamma ; (res_g + gamma)'] + W [B';A']'  SYRK_LN_MN :652
    //   L   = chol of the 4 input columns of S               POTRF_L_MN(nv+1, nu) :653
    //   P_k = S_xx - Ls Ls',  p_k = s_x - Ls l_u             SYRK -1 :655
    // Stored per stage: LU (18 x 4 factor columns, inverse pivots on the diagonal), the packed lower triangle of P_k, and
    // the gradient parts l_u (R_LU4) and p_k (x-part of R_DUX) of the stage record.
    CF_MEM double step_adjust(double a) const { return (a < 1.0) ? a * ((1.0 - a) * 0.99 + a * 0.9999999) : a; }
    CF_MEM void residual_factorize(const double a_raw, const bool do_factor)
    {
        const double a = step_adjust(a_raw);
        double ng = 0, nb = 0, nd = 0, nm = 0, mus = 0;
        double *PS = sm + CF_SM_P;                 // P_{k+1}, lower triangle of a 13 x 20 array
        double *PV = sm + CF_SM_V0;                // p_{k+1}
        double *UXS = sm + CF_SM_V1, *PIS = sm + CF_SM_V2;   // residual part: ux_k (order of vidx()), pi_k broadcast
        double *G = sm + CF_SM_V1, *HD = sm + CF_SM_V2;      // factor part: gradient row / Hessian diagonal (order of vidx())
        pass_begin();
        fetch(N & 1, N, 0, B_M);
        const bool ul = lane < CF_NU, vl = lane < CF_NV;
        const bool xl = lane >= CF_NU && vl;
        const bool sl = stored_lane(), fl = free_lane();
        // (idle lanes read element 8, not 0: in the second half-warp lane 16 sits on the banks of element 0)
        const int ci = xl ? lane - CF_NU : 0, l4 = lane & 3, lv = vl ? lane : 8;
        const int mr = mrow(), vi = vidx();
        int cgo, clp;
        col_consts(cgo, clp);
        // tensor-core fragment coordinates (mma.sync.m8n8k4.f64): group row / k / n index and column pair
        const int fg = lane >> 2, fq = lane & 3;
        const int rl = lane < CF_MROWS ? lane : 17;
        // Row order of the tiles: the 14 stored rows, then (row 14) the gradient / res_b row; row 15 is padding.  Row r of the
        // tiles is row lr(r) of the 18 x 4 input-column block, whose rows stay in the order of the stage variables.
        const int r1 = 8 + fg;   // this lane's row in the second row tile
        const int lr0 = fg < CF_NU ? fg : fg + CF_NF, lr1 = r1 < CF_CR ? r1 + CF_NF : 17;
        // A fragments of the first product, as offsets into the staged block (loop-invariant): am[t][kk] = M[8t+fg][4kk+fq];
        // tile 1 takes row 14 from the res_b vector, row 15 and the columns past nx - 1 from a staged zero (the pad of b_k)
        const int zo = R_B + CF_NX;
        const int ao0 = B_M + fq * CF_CST + fg;   // tile 0, kk < 3: + kk * 4 * CF_CST
        int ao03 = (fq == 0) ? B_M + 12 * CF_CST + fg : zo, ao1[4];
        CF_UNROLL
        for (int kk = 0; kk < 4; kk++) {
            const int kc = 4 * kk + fq;
            ao1[kk] = (kc < CF_NX && r1 <= CF_CR) ? (r1 < CF_CR ? B_M + kc * CF_CST + r1 : R_B + kc) : zo;
            cf_keep(ao1[kk]);
        }
        cf_keep(ao03);
        // W tile stores / fragment loads (row stride 20) and the rows of this lane in the 18 x 4 input-column block, which
        // sits behind the 16 rows of W; row 15 of W is all zero and stands in where a lane has no row (fg == 7)
        int wso = fg * CF_WST + 2 * fq, wfo = fg * CF_WST + fq;
        int lo0 = 16 * CF_WST + lr0 * CF_LUST + fq, lo1 = (r1 <= CF_CR) ? 16 * CF_WST + lr1 * CF_LUST + fq : 15 * CF_WST + fq;
        // Schur-complement tile rows land in the array of P: tile row r = state r - 1, and the gradient row (r = 14) in the
        // row behind the 13 x 20 array, which is the vector p (indexed like the columns of P: state i at i + CF_PC0)
        int pso = (fg - 1) * CF_ALST + 2 * fq;
        cf_keep(wso); cf_keep(wfo); cf_keep(lo0); cf_keep(lo1); cf_keep(pso);
        static_assert(CF_SM_V0 == CF_SM_P + CF_NX * CF_ALST, "p_{k+1} is row 13 of the shared-memory array of P");
        // free states: their rows of the input-column block as A / B fragment (Ls[p_fg][fq], zero for fg >= CF_NF), the
        // free x free block of P_{k+1} as C fragment, and this lane's own column in the arrays of P / p
        int lfo = (fg < CF_NF) ? 16 * CF_WST + (CF_NU + fg) * CF_LUST + fq : 15 * CF_WST + fq;
        int pfo = (fg < CF_NF ? fg : 0) * CF_ALST + CF_PCF + 2 * (fq & 1);
        const int pcl = CF_PCOL(ci);
        cf_keep(lfo); cf_keep(pfo);
        // P_{k+1} is kept in shared memory as the lower triangle of a 13 x 20 array, row = state index, column = the
        // state's column CF_PCOL (stored states: their tile column, exactly where the Schur-complement tiles fall, so they
        // are stored as whole tiles; free states: columns 16..18); the fragment loads of the next stage's first product swap indices
        // instead (addresses are loop-invariant): pa[kk][h] = address of P[4kk+fq][8h+fg], or of an element that is always
        // zero (column 19 of row 0: filled at the terminal stage, never stored to)
        int pa[4][2];
        CF_UNROLL
        for (int kk = 0; kk < 4; kk++)
            CF_UNROLL
            for (int hh = 0; hh < 2; hh++) {
                const int i = 4 * kk + fq, j = 8 * hh + fg;
                const bool ok = i < CF_NX && j < CF_NX;
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                pa[kk][hh] = ok ? hi * CF_ALST + CF_PCOL(lo) : CF_ALST - 1;
                cf_keep(pa[kk][hh]);
            }
        // The packed lower triangle of P_{k+1} travels to the block of stage k from this shared-memory array, element
        // e = lane + 32 t of the triangle per lane and trip: three coalesced stores per stage (scattering it from the
        // tensor-core fragments cost 87 mostly predicate / index instructions per stage)
        int pk[3];
        CF_UNROLL
        for (int t = 0; t < 3; t++) {
            const int e = lane + 32 * t;
            int i = 0;
            CF_UNROLL
            for (int q = 1; q < CF_NX; q++) i += (e >= cf_tri(q)) ? 1 : 0;
            pk[t] = (e < 91) ? i * CF_ALST + CF_PCOL(e - cf_tri(i)) : -1;
            cf_keep(pk[t]);
        }
        double ux_next = 0.0;   // lanes 4..16: x-part of ux_{k+1} (new iterate)
        double pi_k = 0.0;      // lanes 4..16: pi_k (new iterate), read from the record of stage k+1
        CF_NOUNROLL
        for (int k = N; k >= 0; k--) {
            const int bf = k & 1;
            const bool kl = k < N;
            wait(bf);
            cf_syncwarp();  // every lane is done with buffer bf^1 (incl. the W block of stage k+1), PS/PV are complete
            if (k > 0) fetch(bf ^ 1, k - 1, 0, B_RD);
            if (do_factor && kl) {   // P_{k+1}, p_{k+1} are complete in shared memory since the barrier above
                double *LFk = blk(k) + B_PX;
                CF_UNROLL
                for (int t = 0; t < 3; t++)
                    if (pk[t] >= 0) LFk[lane + 32 * t] = PS[pk[t]];
                if (xl) rec(k + 1)[R_DUX + lane] = PV[pcl];   // p_{k+1} for the forward sweep
            }
            double *VS = buf(bf);   // block k from offset 0
            double *rk = rec(k);
            // ---------------- update + residuals of stage k
            const double uxc = vl ? VS[R_UX + lv] + a * VS[R_DUX + lv] : 0.0;
            if (vl) rk[R_UX + lane] = uxc;
            const double pim = (k > 0 && xl) ? VS[R_PI + ci] + a * VS[R_DPI + ci] : 0.0;   // pi_{k-1}
            if (k > 0 && xl) rk[R_PI + ci] = pim;
            const double Hk = hess(k < N ? k : 0);
            double rg = ((k == N) ? HN : Hk) * uxc + VS[R_RQ + lv] - pim;
            double Gam = 0.0, gam = 0.0;
            {   // bounds of input l4 (lanes >= 4 compute duplicates that are never stored nor counted)
                double ll = VS[R_LAM + l4] + a * VS[R_DLAM + l4], lu = VS[R_LAM + 4 + l4] + a * VS[R_DLAM + 4 + l4];
                double tl = VS[R_T + l4] + a * VS[R_DT + l4], tu = VS[R_T + 4 + l4] + a * VS[R_DT + 4 + l4];
                ll = ll <= CF_LAM_MIN ? CF_LAM_MIN : ll; lu = lu <= CF_LAM_MIN ? CF_LAM_MIN : lu;
                tl = tl <= CF_T_MIN ? CF_T_MIN : tl; tu = tu <= CF_T_MIN ? CF_T_MIN : tu;
                const double rdl = VS[R_D + l4] + tl - uxc, rdu = VS[R_D + 4 + l4] + tu + uxc;
                const double rml = ll * tl, rmu = lu * tu;
                if (ul && kl) {
                    rk[R_LAM + lane] = ll; rk[R_LAM + 4 + lane] = lu;
                    rk[R_T + lane] = tl; rk[R_T + 4 + lane] = tu;
                    rk[R_RESD + lane] = rdl; rk[R_RESD + 4 + lane] = rdu;
                    rk[R_BKP + lane] = rml; rk[R_BKP + 4 + lane] = rmu;
                    rk[R_RESM + lane] = rml - CF_TAU_MIN; rk[R_RESM + 4 + lane] = rmu - CF_TAU_MIN;  // predictor rhs
                    rg += lu - ll;
                    mus += rml + rmu;
                    cf_amax(nd, rdl); cf_amax(nd, rdu);
                    cf_amax(nm, rml); cf_amax(nm, rmu);
                    // Gamma, gamma of the predictor system (x_core_qp_ipm_aux.c:38-111)
                    const double til = cf_rcp(tl), tiu = cf_rcp(tu);
                    Gam = til * ll + tiu * lu;
                    gam = til * ((rml - CF_TAU_MIN) - ll * rdl) - tiu * ((rmu - CF_TAU_MIN) - lu * rdu);
                }
            }
            if (kl) {   // warp-uniform
                if (vl) UXS[vi] = uxc;
                if (xl) PIS[ci] = pi_k;
                cf_syncwarp();
                double *Mk = VS + B_M;
                {   // res_g += [B';A'] pi_k   (row layout: own elements stride 14; a free state's row is a unit vector, and
                    // dropped like the other state rows at stage 0)
                    double s0 = 0.0, s1 = 0.0;
                    CF_UNROLL
                    for (int cp = 0; cp < 6; cp++) {
                        const cf_d2 p2 = cf_ld2(PIS + 2 * cp);
                        s0 += Mk[(2 * cp) * CF_CST + mr] * p2.x;
                        s1 += Mk[(2 * cp + 1) * CF_CST + mr] * p2.y;
                    }
                    s0 += Mk[12 * CF_CST + mr] * PIS[12];
                    rg += sl ? s0 + s1 : ((fl && k > 0) ? pi_k : 0.0);
                }
                {   // res_b_k = (b_k - x_{k+1}) + [A B] ux_k   (column layout: contiguous)
                    const double rb = col_gemv(Mk, UXS, VS + R_B, cgo, clp) - ux_next;
                    cf_syncwarp();   // every lane has read b_k: the staged vector becomes res_b (ROWIN :490)
                    if (xl) {
                        cf_amax(nb, rb);
                        rk[R_RESB + ci] = rb;
                        VS[R_B + ci] = rb;
                    }
                }
            }
            if (vl) { rk[R_RESG + lane] = rg; cf_amax(ng, rg); }
            ux_next = uxc;
            pi_k = pim;
            // ---------------- factorisation of stage k (skipped when the caller predicts that this iterate is final)
            if (!do_factor) continue;
            if (!kl) {
                // terminal stage: no dynamics. P_N = diag(H_N) + reg, p_N = res_g_N; dummy inputs decoupled.
                for (int i = lane; i < 13 * CF_ALST; i += 32) PS[i] = 0.0;
                cf_syncwarp();
                const double hN = HN + CF_REG_PRIM;
                if (xl) {
                    PS[ci * CF_ALST + pcl] = hN;
                    PV[pcl] = rg;    // p_N
                }
                continue;
            }
            const double g = vl ? rg + gam : 0.0, hd = Hk + CF_REG_PRIM + Gam;
            cf_syncwarp();  // res_b is complete
            double *WS = VS;                      // W rows 16 x 16 (CF_WST) overlay the staged block once M is in registers,
            double *LUs = VS + 16 * CF_WST;       // the 18 x 4 input-column block sits behind them
            // ---- W(15x13) = [M;res_b'](15x13) * P(13x13): row tiles t (rows 8t+fg), column tiles 0..1, K padded to 16
            double am[2][4];     // am[t][kk] = M[8t+fg][4kk+fq]: A fragment here, B fragment (M') of the second product
            double wt[2][2][2];
            CF_UNROLL
            for (int t = 0; t < 2; t++) { wt[t][0][0] = wt[t][0][1] = wt[t][1][0] = wt[t][1][1] = 0.0; }
            CF_UNROLL
            for (int kk = 0; kk < 4; kk++) {
                const double b0 = PS[pa[kk][0]];   // P[kc][fg]
                const double b1 = PS[pa[kk][1]];   // P[kc][8+fg]
                am[0][kk] = VS[kk < 3 ? ao0 + kk * 4 * CF_CST : ao03];
                am[1][kk] = VS[ao1[kk]];
                CF_UNROLL
                for (int t = 0; t < 2; t++) {
                    cf_dmma(wt[t][0][0], wt[t][0][1], am[t][kk], b0);
                    cf_dmma(wt[t][1][0], wt[t][1][1], am[t][kk], b1);
                }
            }
            // row 14 (tile 1, fg == 6): Pb_k = P res_b (ROWEX :622), then + p_{k+1}' (GEAD :623)
            if (r1 == CF_CR) {
                double *pb = rk + R_PB;
                CF_UNROLL
                for (int tp = 0; tp < 2; tp++)
                    CF_UNROLL
                    for (int e = 0; e < 2; e++) {
                        const int c = 8 * tp + 2 * fq + e;
                        if (c < CF_NX) { pb[c] = wt[1][tp][e]; wt[1][tp][e] += PV[CF_PCOL(c)]; }
                    }
            }
            cf_syncwarp();  // every lane holds its M fragments: the staged block may be overwritten by W
            CF_UNROLL
            for (int t = 0; t < 2; t++) {
                cf_st2(WS + wso + 8 * t * CF_WST, wt[t][0][0], wt[t][0][1]);
                cf_st2(WS + wso + 8 * t * CF_WST + 8, wt[t][1][0], wt[t][1][1]);
            }
            // rows of the free states in the input-column block: S[p_i][u_j] = W[u_j][i] (their rows of [B';A'] are unit
            // vectors), held by the lanes of tile row j < nu
            if (fg < CF_NU && fq < 2) {
                CF_UNROLL
                for (int e = 0; e < 2; e++)
                    if (2 * fq + e < CF_NF) LUs[(CF_NU + 2 * fq + e) * CF_LUST + fg] = wt[0][0][e];
            }
            if (vl) { G[vi] = g; HD[vi] = hd; }
            if (fl) PS[ci * CF_ALST + pcl] += hd;   // S[p_i][p_i] = P_{k+1}[i][i] + the diagonal (the first product has read P_{k+1})
            cf_syncwarp();
            // ---- S = D + W * M': A fragments from W, B fragments are the M fragments already in registers
            double wf[2][4];
            CF_UNROLL
            for (int t = 0; t < 2; t++)
                CF_UNROLL
                for (int kk = 0; kk < 4; kk++) wf[t][kk] = WS[wfo + 8 * t * CF_WST + 4 * kk];
            double sx[2][2][2];  // lower tiles (t, tp <= t): S[8t+fg][8tp+2fq+{0,1}]
            const double *GG = (r1 == CF_CR) ? G : WS + 15 * CF_WST;   // row 15 of W: sixteen zeros
            CF_UNROLL
            for (int t = 0; t < 2; t++) {
                CF_UNROLL
                for (int tp = 0; tp <= t; tp++) {
                    double s0 = 0.0, s1 = 0.0;
                    CF_UNROLL
                    for (int kk = 0; kk < 4; kk++) cf_dmma(s0, s1, wf[t][kk], am[tp][kk]);
                    const int r = 8 * t + fg, c0 = 8 * tp + 2 * fq;
                    if (t == 1) { const cf_d2 g2 = cf_ld2(GG + c0); s0 += g2.x; s1 += g2.y; }   // gradient row (zeros elsewhere)
                    if (r == c0) s0 += HD[r];
                    if (r == c0 + 1) s1 += HD[r];
                    sx[t][tp][0] = s0; sx[t][tp][1] = s1;
                }
            }
            if (fq < 2) {   // the 4 input columns of the stored rows and of the gradient row
                cf_st2(VS + lo0 + fq, sx[0][0][0], sx[0][0][1]);
                if (r1 <= CF_CR) cf_st2(VS + lo1 + fq, sx[1][0][0], sx[1][0][1]);
            }
            cf_syncwarp();
            // ---- POTRF_L_MN(nv+1, nu): the 4 input columns, one per step, lane = row; non-positive pivot -> 0
            // (BLASFEO kernel_dgemm_4x4_lib4.c:5701-5714)
            double o[CF_NU];
            {
                // lanes 18..31 carry no row: they must not read row 17 while lane 17 rewrites it below (racecheck)
                cf_d2 o01 = {0.0, 0.0}, o23 = {0.0, 0.0};
                if (lane < CF_MROWS) { o01 = cf_ld2(LUs + rl * CF_LUST); o23 = cf_ld2(LUs + rl * CF_LUST + 2); }
                double og[CF_NU];
                o[0] = o01.x; o[1] = o01.y; o[2] = o23.x; o[3] = o23.y;
                CF_UNROLL
                for (int j = 0; j < CF_NU; j++) {
                    double v = o[j];
                    CF_UNROLL
                    for (int c = 0; c < j; c++) v -= o[c] * LUs[j * CF_LUST + c];
                    const double piv = cf_shfl(v, j);
                    double dj, inv;
                    cf_sqrt_rsqrt(piv, dj, inv);
                    if (!(piv > 0.0)) { dj = 0.0; inv = 0.0; flags |= CF_FLAG_BAD_PIVOT; }   // piv is warp-uniform
                    // (rows above the pivot keep a finite throw-away value: the strictly upper part of the 4 x 4 block is
                    // never used -- no select for it)
                    const double vi_ = v * inv;
                    o[j] = (rl == j) ? dj : vi_;
                    if (lane < CF_MROWS) LUs[rl * CF_LUST + j] = o[j];   // lanes 18..31 mirror row 17 in registers only (racecheck)
                    og[j] = (rl == j) ? inv : vi_;
                    cf_syncwarp();
                }
                // factor columns to global memory (LU block of stage k), inverse pivots on the diagonal; the gradient row
                // l_u goes where the forward sweep picks it up
                double *LFk = blk(k) + B_LU;
                if (lane < CF_MROWS) {
                    cf_st2(LFk + lane * 4, og[0], og[1]);
                    cf_st2(LFk + lane * 4 + 2, og[2], og[3]);
                }
                if (lane == 17) {
                    cf_st2(rk + R_LU4, o[0], o[1]);
                    cf_st2(rk + R_LU4 + 2, o[2], o[3]);
                }
            }
            // ---- Schur complement on the tensor cores (K = 4 = the input columns): S -= Ls Ls'
            const double la[2] = {VS[lo0], VS[lo1]};
            const double lf = VS[lfo];             // rows of the free states
            const cf_d2 pf = cf_ld2(PS + pfo);     // S[p_i][p_j] = P_{k+1}[i][j] (+ the diagonal, added above)
            cf_syncwarp();  // all reads of PS/PV are complete; they are rewritten below
            CF_UNROLL
            for (int t = 0; t < 2; t++) {
                const bool row_on = (t == 0) ? fg >= CF_NU : fg < 7;   // tile row = a state (or, row 14, the gradient row)
                CF_UNROLL
                for (int tp = 0; tp <= t; tp++) {
                    cf_dmma(sx[t][tp][0], sx[t][tp][1], -la[t], la[tp]);
                    // P_k / p_k: whole tile rows into the shared-memory array (positions above the diagonal are never read;
                    // row 13 of the array is p_k); the input columns of the first column tile are skipped
                    if (row_on && (tp > 0 || fq >= 2)) cf_st2(PS + pso + 8 * t * CF_ALST + 8 * tp, sx[t][tp][0], sx[t][tp][1]);
                }
                // columns of the free states: their rows of [B';A'] are unit vectors, so S[r][p_c] = W[r][c] -- still in the
                // accumulators of the first product (+ the gradient of p_c in the gradient row); same update
                double s0 = wt[t][0][0], s1 = wt[t][0][1];
                if (t == 1 && r1 == CF_CR) { const cf_d2 g2 = cf_ld2(G + CF_CR + 2 * (fq & 1)); s0 += g2.x; s1 += g2.y; }
                cf_dmma(s0, s1, -la[t], lf);
                if (row_on && fq == 0) cf_st2(PS + pso + 8 * t * CF_ALST + CF_PCF, s0, s1);   // columns 16, 17
                if (row_on && fq == 1) PS[pso + 8 * t * CF_ALST + CF_PCF] = s0;                // column 18 (pso holds 2 fq)
            }
            {   // free x free block, rows fg < CF_NF
                double s0 = pf.x, s1 = pf.y;
                cf_dmma(s0, s1, -lf, lf);
                if (fg < CF_NF && fq == 0) cf_st2(PS + pfo, s0, s1);   // (above the diagonal: never read)
                if (fg < CF_NF && fq == 1) PS[pfo] = s0;
            }
        }
        nrm[0] = cf_warp_max(ng); nrm[1] = cf_warp_max(nb); nrm[2] = cf_warp_max(nd); nrm[3] = cf_warp_max(nm);
        mu = cf_warp_sum(mus) * (1.0 / (double) (2 * CF_NU * N));
        cf_syncwarp();
    }

    // gamma for the condensed right-hand side (x_core_qp_ipm_aux.c:38-111); lanes 0..3 of stage k<N.
    // rm_mode: 0 predictor (bkp - tau_min), 1 corrector (bkp + dt*dlam - sigma_mu, stored),
    //          2 re-centering (bkp - sigma_mu, stored), 3 use the stored RESM as is.
    // `q` points at field R_DLAM of stage k's record, either in global memory or in its staged copy.
    CF_MEM void bound_terms(int k, const double *q, int rm_mode, double sigma_mu, double &Gam, double &gam)
    {
        const double *f = q - R_DLAM;
        double ll = f[R_LAM + lane], lu = f[R_LAM + 4 + lane];
        const double til = cf_rcp(f[R_T + lane]), tiu = cf_rcp(f[R_T + 4 + lane]);
        double rml, rmu;
        if (rm_mode == 3) { rml = rec(k)[R_RESM + lane]; rmu = rec(k)[R_RESM + 4 + lane]; }   // (not staged by the backward sweep)
        else {
            rml = f[R_BKP + lane]; rmu = f[R_BKP + 4 + lane];
            if (rm_mode == 0) { rml -= CF_TAU_MIN; rmu -= CF_TAU_MIN; }
            else if (rm_mode == 1) {
                rml = rml + f[R_DT + lane] * f[R_DLAM + lane] - sigma_mu;
                rmu = rmu + f[R_DT + 4 + lane] * f[R_DLAM + 4 + lane] - sigma_mu;
            } else { rml -= sigma_mu; rmu -= sigma_mu; }
            if (rm_mode != 0) { rec(k)[R_RESM + lane] = rml; rec(k)[R_RESM + 4 + lane] = rmu; }
        }
        double gl = til * (rml - ll * f[R_RESD + lane]);
        double gu = tiu * (rmu - lu * f[R_RESD + 4 + lane]);
        Gam = til * ll + tiu * lu;
        gam = gl - gu;
    }

    // Forward substitution shared by the factorise-and-solve (:536-570) and the rhs-only solve (:1250-1290): the
    // backward sweep before it (factorize or backward_rhs) left l_u of stage k in R_LU4, p_k in R_DUX of stage k's record and
    // the complementarity rhs in R_RESM.  Computes dux, dpi, then dlam, dt (:741-758, x_core_qp_ipm_aux.c:117-142), the
    // step length ingredients (:146-216) and the inf-norms of the linear-system residual (OCP_QP_RES_COMPUTE_LIN,
    // x_ocp_qp_res.c:474-598) on the fly.  Branch-free: every lane computes with clamped indices (lanes with equal
    // lane & 3 hold identical input/bound quantities), stores and norm contributions are predicated.
    // `need_pi`: the multiplier step dpi is only consumed by the variable update (and by the optional linear-residual
    // check), never by the corrector -- the affine (predictor) solve skips it together with the P_{k+1} it would read.
    // MODE 1 / 2 are the two cold variants of iterative refinement (OCP_QP_IPM_DELTA_STEP, x_ocp_qp_ipm.c:2275-2366; only
    // with lin_res_check = 2).  1: no substitution -- the STORED corrector step (dux, dpi, dlam, dt of the records) is put
    // through the residual formulas and the residual VECTORS of the linear system are left in the right-hand-side fields
    // (res_g, res_b, res_d, res_m of every stage; they are dead until the next residual sweep rewrites them), a copy of dux
    // in the R_BKP | R_PB area.  2: the right-hand side is such a residual: solve for the correction, ADD it to the stored
    // step, leave the new residual in place, step length from the sums.
    CF_MEM void forward(const bool need_pi_) { forward_t<0>(need_pi_); }
    // Compiled into the GENERAL kernel variants only (VDT), which the API selects for lin_res_check >= 2: three copies of
    // the forward sweep in the benchmarked kernels cost 1.8 % there (profiles/README.md) although they are never executed.
    static constexpr bool HAS_REFINE = VDT;
    CF_MEM bool refine_enabled() const { return HAS_REFINE && PG->lin_res_check >= 2; }
    // NPI: -1 = `need_pi_` decides at run time (one copy of the sweep serves predictor and corrector); 0 / 1 = compile-time
    // (CF_SPLIT_FWD: one copy each, without the warp-uniform branches around the multiplier step in the stage loop)
    template <int MODE, int NPI = -1>
    CF_MEM void forward_t(const bool need_pi_)
    {
        double *XS = sm + CF_SM_V0, *DS = sm + CF_SM_V1, *PS = sm + CF_SM_V3;
        // running step lengths to the boundary kept as ratios num/den (den < 0): alpha = min(1, min -lam/dlam, -t/dt)
        double dn = 1.0, dd = -1.0, pn_ = 1.0, pd_ = -1.0;
        double lg = 0, lb = 0, ld = 0, lm = 0;    // linear residual norms
        double dxk = 0.0;        // lanes 4..16: dx_k ; stage 0 has none
        double dpi_prev = 0.0;   // lanes 4..16: dpi_{k-1}
        // the reference's linear-system residual checks: like the refinement, compiled into the general variants only (the
        // API selects them for lin_res_check != 0), so that the benchmarked kernels do not carry the code
        const bool chk = MODE ? true : (CF_CHK_IN_UNIFORM || VDT) && PG->lin_res_check != 0;
        const bool need_pi = (NPI >= 0 ? NPI != 0 : need_pi_) || chk;
        const int VO = R_LAM, VN = (need_pi ? CF_SB : B_PX) - R_LAM;   // staged part of the stage block: [R_LAM, end | B_PX)
        pass_begin();
        if (N > 0) fetch(0, 0, VO, VN);
        if (lane < 20) { XS[lane] = 0.0; PS[lane] = 0.0; }
        const bool ul = lane < CF_NU, vl = lane < CF_NV;
        const bool xl = lane >= CF_NU && vl;
        // (idle lanes read element 8, not 0: in the second half-warp lane 16 sits on the banks of element 0)
        const int ci = xl ? lane - CF_NU : 0, l4 = lane & 3, lv = vl ? lane : 8;
        const int vi = vidx();
        int cgo, clp;
        col_consts(cgo, clp);
        // P_{k+1} travels packed (lower triangle) and is expanded to full symmetric rows in shared memory: element
        // e = lane + 32 t of the packed triangle goes to (i,j) and (j,i)
        double *PE = sm + CF_SM_P;
        int pe_a[3], pe_b[3];
        CF_UNROLL
        for (int t = 0; t < 3; t++) {
            const int e = lane + 32 * t;
            int i = 0;
            CF_UNROLL
            for (int q = 1; q < CF_NX; q++) i += (e >= cf_tri(q)) ? 1 : 0;
            const int j = e - cf_tri(i);
            pe_a[t] = (e < 91) ? i * CF_PST + j : -1;
            pe_b[t] = j * CF_PST + i;
            cf_keep(pe_a[t]); cf_keep(pe_b[t]);   // (kept: ptxas re-derived them in every stage, 2.7 % of all instructions)
        }
        CF_NOUNROLL
        for (int k = 0; k < N; k++) {
            const int bf = k & 1;
            double *rk = rec(k);
            const double pnext = need_pi ? rec(k + 1)[R_DUX + lv] : 0.0;  // p_{k+1} left by the backward sweep (x lanes)
            wait(bf);
            cf_syncwarp();  // every lane is done with buffer bf^1 and with the expanded P of the previous stage
            if (k + 1 < N) fetch(bf ^ 1, k + 1, VO, VN);
            const double *VS = buf(bf) - VO;   // VS[offset within the stage block]
            const double *Mk = VS + B_M, *LU = VS + B_LU, *LX = PE;
            if (need_pi) {   // warp-uniform
                CF_UNROLL
                for (int t = 0; t < 3; t++) {
                    if (pe_a[t] >= 0) {
                        const double v = VS[B_PX + lane + 32 * t];
                        PE[pe_a[t]] = v;
                        PE[pe_b[t]] = v;
                    }
                }
            }
            // ---- u-part: du = Luu^-T ( -l_u - Lxu' dx )      TRSV_LTN_MN(nv, nu); input l4 on every lane
            double v;
#if CF_KSPLIT_LXU
            {   // the 13 terms of input l4 are spread over the 8 lane groups (lane >> 2 takes states g and g + 8: the factor
                // rows 4.. are then two contiguous 32-lane loads) and summed by three exchanges, instead of every lane
                // running through all 13 (20 loads per lane)
                const int g = lane >> 2;
                double t = LU[16 + lane] * XS[g];
                const double t2 = LU[lane < 20 ? 48 + lane : 48] * XS[8 + g];
                t += (g < 5) ? t2 : 0.0;
                t += cf_shfl(t, lane ^ 4);
                t += cf_shfl(t, lane ^ 8);
                t += cf_shfl(t, lane ^ 16);
                v = -VS[R_LU4 + l4] - t;
            }
#else
            {
                double v0 = -VS[R_LU4 + l4], v1 = 0.0;
                CF_UNROLL
                for (int ip = 0; ip < 6; ip++) {
                    const cf_d2 x2 = cf_ld2(XS + 2 * ip);
                    v0 -= LU[(CF_NU + 2 * ip) * 4 + l4] * x2.x;
                    v1 -= LU[(CF_NU + 2 * ip + 1) * 4 + l4] * x2.y;
                }
                v0 -= LU[16 * 4 + l4] * XS[12];
                v = v0 + v1;
            }
#endif
            const double invd = LU[l4 * 4 + l4];   // inverse pivot
            double du = 0.0;
            CF_UNROLL
            for (int j = CF_NU - 1; j >= 0; j--) {
                const double duj = cf_shfl(v * invd, j);
                const double vn = v - LU[j * 4 + l4] * duj;
                du = (l4 == j) ? duj : du;
                v = (l4 < j) ? vn : v;
            }
            if constexpr (HAS_REFINE && MODE == 0) {
                if (chk && need_pi_ && PG->lin_res_check == 4) du *= 0.9;   // tests: a deliberately inaccurate corrector
            }
            if constexpr (MODE == 1) du = rk[R_DUX + l4];   // the stored step
            const double duxk = (MODE == 1) ? (vl ? rk[R_DUX + lv] : 0.0) : (ul ? du : dxk);  // lane r: dux_k[r]
            if constexpr (MODE == 2) {
                if (vl) { const double tot = rk[R_BKP + lane] + duxk; rk[R_DUX + lane] = tot; rk[R_BKP + lane] = tot; }
            } else if constexpr (MODE == 1) {
                if (vl) rk[R_BKP + lane] = duxk;
            } else {
                if (vl && need_pi) rk[R_DUX + lane] = duxk;   // the affine primal step itself is never read again
            }
            // ---- dlam, dt, alpha for the bounds of input l4
            double dlam_l, dlam_u;
            {
                const double ll = VS[R_LAM + l4], lu = VS[R_LAM + 4 + l4];
                const double tl = VS[R_T + l4], tu = VS[R_T + 4 + l4];
                const double rdl = VS[R_RESD + l4], rdu = VS[R_RESD + 4 + l4];
                const double rml = VS[R_RESM + l4], rmu = VS[R_RESM + 4 + l4];
                const double til = cf_rcp(tl), tiu = cf_rcp(tu);
                double dtl = du, dtu = -du;
                dlam_l = -til * (rml + (ll * dtl) - (ll * rdl));
                dlam_u = -tiu * (rmu + (lu * dtu) - (lu * rdu));
                dtl -= rdl; dtu -= rdu;
                if constexpr (MODE == 1) {
                    dlam_l = rk[R_DLAM + l4]; dlam_u = rk[R_DLAM + 4 + l4];
                    dtl = rk[R_DT + l4]; dtu = rk[R_DT + 4 + l4];
                }
                // the step the length test sees and the records keep: this solve's, or (refinement) the sum so far
                double al = dlam_l, au = dlam_u, bl = dtl, bu = dtu;
                if constexpr (MODE == 2) {
                    al += rk[R_DLAM + l4]; au += rk[R_DLAM + 4 + l4];
                    bl += rk[R_DT + l4]; bu += rk[R_DT + 4 + l4];
                }
                if (MODE != 1 && ul) {
                    rk[R_DLAM + lane] = al; rk[R_DLAM + 4 + lane] = au;
                    rk[R_DT + lane] = bl; rk[R_DT + 4 + lane] = bu;
                }
                // a*d > v with a = n/dd, dd < 0  <=>  n*d < v*dd ; then the new ratio is v/d (d < 0)
                bool c;
                c = dn * al < ll * dd; dn = c ? ll : dn; dd = c ? al : dd;
                c = pn_ * bl < tl * pd_; pn_ = c ? tl : pn_; pd_ = c ? bl : pd_;
                c = dn * au < lu * dd; dn = c ? lu : dn; dd = c ? au : dd;
                c = pn_ * bu < tu * pd_; pn_ = c ? tu : pn_; pd_ = c ? bu : pd_;
                if (chk) {   // linear residuals of the complementarity / bound rows
                    const double qdl = rdl + dtl - du, qdu = rdu + dtu + du;
                    const double qml = rml + ll * dtl + dlam_l * tl, qmu = rmu + lu * dtu + dlam_u * tu;
                    cf_amax(ld, qdl); cf_amax(ld, qdu);
                    cf_amax(lm, qml); cf_amax(lm, qmu);
                    if (MODE != 0 && ul) {
                        rk[R_RESD + lane] = qdl; rk[R_RESD + 4 + lane] = qdu;
                        rk[R_RESM + lane] = qml; rk[R_RESM + 4 + lane] = qmu;
                    }
                }
            }
            // stationarity residual, part 1: H dux + rhs_g - dpi_{k-1} + dlam_ub - dlam_lb
            double rgl = 0.0;
            if (chk) {
                rgl = hess(k) * duxk + rk[R_RESG + lv] - dpi_prev;   // (not part of the staged range: cold path)
                rgl += ul ? dlam_u - dlam_l : 0.0;
            }
            // ---- dx+ = [A B] dux + res_b        GEMV_T, column layout of M_k (contiguous)
            if (vl) DS[vi] = duxk;
            cf_syncwarp();
            double dxn;
            {
                const double sacc = col_gemv(Mk, DS, nullptr, cgo, clp), rbk = VS[R_RESB + ci];
                dxn = xl ? sacc + rbk : 0.0;
                if constexpr (MODE == 1) dxn = xl ? rec(k + 1)[R_DUX + lane] : 0.0;
                if (chk) cf_amax(lb, xl ? (rbk - dxn) + sacc : 0.0);
                if (MODE != 0 && xl) rk[R_RESB + ci] = (rbk - dxn) + sacc;
                if (xl) XS[ci] = dxn;
            }
            double dpik = 0.0;
            if (need_pi) {   // warp-uniform
            cf_syncwarp();
            // ---- dpi = P_{k+1} dx+ + p_{k+1}   (GEMV_N :712,729)
                const double *Li = LX + ci * CF_PST;   // row ci of the symmetric P_{k+1}
                double z0 = pnext, z1 = 0.0;
                CF_UNROLL
                for (int cp = 0; cp < 6; cp++) {
                    const cf_d2 p2 = cf_ld2(Li + 2 * cp), x2 = cf_ld2(XS + 2 * cp);
                    z0 += p2.x * x2.x;
                    z1 += p2.y * x2.y;
                }
                z0 += Li[12] * XS[12];
                dpik = xl ? z0 + z1 : 0.0;
                if constexpr (MODE == 1) dpik = xl ? rec(k + 1)[R_DPI + ci] : 0.0;
                if (xl) {   // the record of stage k+1 holds pi_k
                    if (MODE != 1) rec(k + 1)[R_DPI + ci] = (MODE == 2) ? rec(k + 1)[R_DPI + ci] + dpik : dpik;
                    PS[ci] = dpik;
                }
            }
            if (chk) {   // warp-uniform
            cf_syncwarp();
            // stationarity residual, part 2: + [B';A'] dpi_k   (row layout; unit rows of the free states, none at stage 0)
                const int mr = mrow();
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int cp = 0; cp < 6; cp++) {
                    const cf_d2 p2 = cf_ld2(PS + 2 * cp);
                    s0 += Mk[(2 * cp) * CF_CST + mr] * p2.x;
                    s1 += Mk[(2 * cp + 1) * CF_CST + mr] * p2.y;
                }
                s0 += Mk[12 * CF_CST + mr] * PS[12];
                const double sp = stored_lane() ? s0 + s1 : ((free_lane() && k > 0) ? dpik : 0.0);
                cf_amax(lg, vl ? rgl + sp : 0.0);
                if (MODE != 0 && vl) rk[R_RESG + lane] = rgl + sp;
            }
            dpi_prev = dpik;
            dxk = dxn;
        }
        // terminal stage: no inputs, no bounds, no dynamics
        if (vl) {
            const double duxN = ul ? 0.0 : dxk;
            double *rN = rec(N);
            const double qg = chk ? HN * duxN + rN[R_RESG + lane] - dpi_prev : 0.0;
            if constexpr (MODE == 2) {
                const double tot = rN[R_BKP + lane] + duxN;
                rN[R_DUX + lane] = tot; rN[R_BKP + lane] = tot;
            } else if constexpr (MODE == 1) {
                rN[R_BKP + lane] = duxN;
            } else {
                if (need_pi) rN[R_DUX + lane] = duxN;
            }
            if (chk) cf_amax(lg, qg);
            if (MODE != 0) rN[R_RESG + lane] = qg;
        }
        if (chk) { lin[0] = cf_warp_max(lg); lin[1] = cf_warp_max(lb); lin[2] = cf_warp_max(ld); lin[3] = cf_warp_max(lm); }
        else { lin[0] = lin[1] = lin[2] = lin[3] = 0.0; }
        // alpha = min(1, prim, dual) as in x_core_qp_ipm_aux.c:146-216 (running values are negative)
        const double a_p = cf_warp_max(pn_ / pd_), a_d = cf_warp_max(dn / dd);
        alpha = -(a_p > a_d ? a_p : a_d);
        cf_syncwarp();
    }

    // OCP_QP_SOLVE_KKT_STEP, backward vector recursion with cached Pb (x_ocp_qp_kkt.c:1147-1245):
    // leaves l_k = [L^-1 rhs]_u ; p_k in dux_k for the forward sweep.
    // FRESH (iterative refinement only): the right-hand side is new, so P_{k+1} res_b is computed instead of taken from
    // the cache R_PB of the factorisation (use_Pb = 0, x_ocp_qp_kkt.c:1187-1190); the packed P_{k+1} is read from the
    // stage block in global memory.
    CF_MEM void backward_rhs(int rm_mode, double sigma_mu) { backward_rhs_t<false>(rm_mode, sigma_mu); }
    template <bool FRESH>
    CF_MEM void backward_rhs_t(int rm_mode, double sigma_mu)
    {
        double *TS = sm + CF_SM_V0;
        const int VO = R_BKP, VN = B_BWE - R_BKP;   // staged part of the stage block: [R_BKP, B_BWE)
        pass_begin();
        if (N > 0) fetch(0, N - 1, VO, VN);
        // terminal stage: rhs = res_g_N, nothing to eliminate (dummy inputs are zero)
        double pn = 0.0;  // lanes 4..16: p_{k+1}
        if (lane < CF_NV) {
            pn = rec(N)[R_RESG + lane];
            rec(N)[R_DUX + lane] = pn;
        }
        CF_NOUNROLL
        for (int k = N - 1; k >= 0; k--) {
            const int bf = (N - 1 - k) & 1;
            cf_syncwarp();  // previous stage's reads of TS and of buffer bf^1 are complete
            if (k > 0) fetch(bf ^ 1, k - 1, VO, VN);
            wait(bf);
            const double *VS = buf(bf) - VO;   // VS[offset within the stage block]
            const double *Mk = VS + B_M, *LU = VS + B_LU;
            double Gam = 0.0, gam = 0.0;
            if (lane < CF_NU) bound_terms(k, VS + R_DLAM, rm_mode, sigma_mu, Gam, gam);
            double rhs = 0.0;
            if (lane < CF_NV) rhs = VS[R_RESG + lane] + gam;
            if constexpr (FRESH) {
                double *RB = sm + CF_SM_V1;
                const bool xq = lane >= CF_NU && lane < CF_NV;
                const int iq = xq ? lane - CF_NU : 0;
                if (xq) RB[iq] = rec(k)[R_RESB + iq];   // (not staged by the backward sweep)
                cf_syncwarp();
                const double *Pp = blk(k) + B_PX;
                double pb = 0.0;
                CF_NOUNROLL
                for (int j = 0; j < CF_NX; j++) pb += Pp[iq >= j ? cf_tri(iq) + j : cf_tri(j) + iq] * RB[j];
                if (xq) TS[iq] = pn + pb;
            } else {
                if (lane >= CF_NU && lane < CF_NV) TS[lane - CF_NU] = pn + VS[R_PB + lane - CF_NU];
            }
            cf_syncwarp();
            if (stored_lane()) {
                const int mr = mrow();
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int cp = 0; cp < 6; cp++) {
                    const cf_d2 t2 = cf_ld2(TS + 2 * cp);
                    s0 += Mk[(2 * cp) * CF_CST + mr] * t2.x;
                    s1 += Mk[(2 * cp + 1) * CF_CST + mr] * t2.y;
                }
                s0 += Mk[12 * CF_CST + mr] * TS[12];
                rhs += s0 + s1;
            } else if (free_lane() && k > 0) rhs += TS[lane - CF_NU];   // unit row of a free state (none at stage 0)
            // TRSV_LNN_MN(nv, nu): rows of the 4 input columns of L_k
            const int rl = lane < CF_NV ? lane : 0;
            const cf_d2 l01 = cf_ld2(LU + rl * 4), l23 = cf_ld2(LU + rl * 4 + 2);
            const double Lr[CF_NU] = {l01.x, l01.y, l23.x, l23.y};
            const double invd = LU[(lane & 3) * 5];   // inverse pivot (only lanes 0..3 feed the shuffles)
            CF_UNROLL
            for (int j = 0; j < CF_NU; j++) {
                const double zj = cf_shfl(rhs * invd, j);
                if (lane == j) rhs = zj;
                else if (lane > j && lane < CF_NV) rhs -= Lr[j] * zj;
            }
            if (lane < CF_NV) rec(k)[lane < CF_NU ? R_LU4 + lane : R_DUX + lane] = rhs;
            pn = rhs;
        }
        cf_syncwarp();
    }

    // COMPUTE_MU_AFF_QP (x_core_qp_ipm_aux.c:329-352): all 32 lanes sweep the bound records.  The same sweep predicts
    // the complementarity residual norm of the iterate the adjusted step would produce (pm_max), used to decide whether
    // the next residual sweep needs to factorise at all.
    CF_MEM void compute_mu_aff()
    {
        double s0 = 0.0, s1 = 0.0, pm = 0.0;
        const double aa = step_adjust(alpha);
        const int e = lane & 7;
        int k = lane >> 3;
        CF_NOUNROLL
        for (; k + 12 < N; k += 16) {   // four stages per trip: sixteen independent loads in flight per lane
            double l[4], d[4], t[4], u[4];
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                const double *r = rec(k + 4 * q);
                l[q] = r[R_LAM + e]; d[q] = r[R_DLAM + e]; t[q] = r[R_T + e]; u[q] = r[R_DT + e];
            }
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                if (q & 1) s1 += (l[q] + alpha * d[q]) * (t[q] + alpha * u[q]);
                else s0 += (l[q] + alpha * d[q]) * (t[q] + alpha * u[q]);
                cf_amax(pm, (l[q] + aa * d[q]) * (t[q] + aa * u[q]));
            }
        }
        CF_NOUNROLL
        for (; k < N; k += 4) {
            const double *ra = rec(k);
            const double la = ra[R_LAM + e], da = ra[R_DLAM + e], ta = ra[R_T + e], ua = ra[R_DT + e];
            s0 += (la + alpha * da) * (ta + alpha * ua);
            cf_amax(pm, (la + aa * da) * (ta + aa * ua));
        }
        mu_aff = cf_warp_sum(s0 + s1) * (1.0 / (double) (2 * CF_NU * N));
        pm_max = cf_warp_max(pm);
    }

    CF_MEM bool lin_res_ok_fact() const
    {   // x_ocp_qp_ipm.c:2029-2040: switch to LQ when any norm > 1e-5 (or NaN)
        return !(lin[0] > 1e-5 || lin[1] > 1e-5 || lin[2] > 1e-5 || lin[3] > 1e-5 || lin[0] != lin[0]);
    }
    CF_MEM bool lin_res_ok_corr() const
    {   // x_ocp_qp_ipm.c:2311-2318
        return (lin[0] < CF_RES_G_MAX || lin[0] < 1e-3 * nrm[0]) && (lin[1] < CF_RES_B_MAX || lin[1] < 1e-3 * nrm[1]) &&
               (lin[2] < CF_RES_D_MAX || lin[2] < 1e-3 * nrm[2]) && (lin[3] < CF_RES_M_MAX || lin[3] < 1e-3 * nrm[3]);
    }
#endif  // CF_CRAZYFLIE
};

// OCP_QP_IPM_SOLVE, delta formulation (x_ocp_qp_ipm.c:2409-2759), written as a small state
// machine so that every pass has exactly ONE call site: each pass is inlined once into the
// kernel (address spaces of all pointers known, warp context in registers) and the code of
// the whole warp program stays small enough for the instruction caches.  Returns HPIPM status.
#if defined(CF_SIMT_EMU)
#define CF_PROF_BEGIN()
#define CF_PROF_END(id)
#else
#define CF_PROF_BEGIN() const long long cf_t0_ = prof ? clock64() : 0
#define CF_PROF_END(id) do { if (prof && cf_lane() == 0) { atomicAdd(prof + 2 * (id), (unsigned long long) (clock64() - cf_t0_)); atomicAdd(prof + 2 * (id) + 1, 1ull); } } while (0)
#endif
template <class CfWarp>
CF_DEV int cf_ipm_solve(CfWarp &w, int &iters, unsigned long long *prof)
{
    w.alpha = 1.0;
    w.flags = 0;
    cf_syncwarp();
    const int itmax = w.PG->max_ipm_iter < CF_ITER_MAX ? w.PG->max_ipm_iter : CF_ITER_MAX;
    enum { ST_RF, ST_FWD, ST_BWD };
    int st = ST_RF, kk = 0, brm = 1;
    bool first = true, predictor = true, do_factor = true, redo = false;
    double sigma_mu = 0.0, mu_aff0 = 0.0;
    for (;;) {
        if (st == ST_RF) {
            // variables += alpha * direction, residuals of the new iterate, and the factorisation of the next affine
            // system, OCP_QP_IPM_DELTA_STEP (:1943-2405).  The factorisation is skipped when the iterate is predicted to
            // be the last one; a wrong prediction costs a second pass with step 0 (same residuals, now factorising).
            CF_PROF_BEGIN();
            w.residual_factorize((first || redo) ? 0.0 : w.alpha, do_factor);
            CF_PROF_END(CF_PROF_RF);
            if (first) { first = false; w.alpha = 1.0; }
            else if (!redo) kk++;
            const bool go = kk < itmax && w.alpha > CF_ALPHA_MIN &&
                            (w.nrm[0] > CF_RES_G_MAX || w.nrm[1] > CF_RES_B_MAX || w.nrm[2] > CF_RES_D_MAX ||
                             fabs(w.nrm[3] - CF_TAU_MIN) > CF_RES_M_MAX);
            if (!go) break;
            redo = !do_factor;
            do_factor = true;
            if (redo) continue;
            predictor = true;
            st = ST_FWD;
        } else if (st == ST_FWD) {
            CF_PROF_BEGIN();
#if CF_SPLIT_FWD
            if (predictor) w.template forward_t<0, 0>(false);
            else w.template forward_t<0, 1>(true);
#else
            w.forward(!predictor);
#endif
            CF_PROF_END(CF_PROF_FWD);
            if (predictor) {
                if (!w.lin_res_ok_fact()) w.flags |= CF_FLAG_LIN_RES_FACT;
                CF_PROF_BEGIN();
                w.compute_mu_aff();
                CF_PROF_END(CF_PROF_MUAFF);
                const double tmp = w.mu_aff / w.mu;
                w.sigma = tmp * tmp * tmp;
                sigma_mu = w.sigma * w.mu;
                sigma_mu = sigma_mu > CF_TAU_MIN ? sigma_mu : CF_TAU_MIN;
                brm = 1;            // centering-corrector rhs
                st = ST_BWD;
            } else {
                mu_aff0 = w.mu_aff;
                CF_PROF_BEGIN();
                w.compute_mu_aff();
                CF_PROF_END(CF_PROF_MUAFF);
                // conditional predictor-corrector (:2230-2273)
                const bool recenter = brm == 1 && w.mu_aff > 2.0 * mu_aff0;
                if (recenter) { brm = 2; st = ST_BWD; }
                else {
                    if (!w.lin_res_ok_corr()) {
                        w.flags |= CF_FLAG_LIN_RES_CORR;
                        if constexpr (CfWarp::HAS_REFINE) if (w.refine_enabled()) {
                            // iterative refinement of the corrector step, at most itref_corr_max = 2 rounds
                            // (x_ocp_qp_ipm.c:2275-2366): residual of the linear system -> right-hand side -> correction
                            w.flags |= CF_FLAG_ITREF;
                            w.template forward_t<1>(true);
                            bool ok = false;
                            for (int r = 0; r < 2 && !ok; r++) {
                                w.template backward_rhs_t<true>(3, 0.0);
                                w.template forward_t<2>(true);
                                ok = w.lin_res_ok_corr();
                            }
                            if (!ok) w.flags |= CF_FLAG_ITREF_LEFT;
                            w.compute_mu_aff();   // the complementarity prediction of the refined step
                        }
                    }
                    // will the updated iterate pass the exit test?  The linear residuals shrink by (1 - step), the
                    // complementarity products were just evaluated; half the tolerances as margin.
                    const double r = 1.0 - w.step_adjust(w.alpha);
                    const bool conv = r * w.nrm[0] < 0.5 * CF_RES_G_MAX && r * w.nrm[1] < 0.5 * CF_RES_B_MAX &&
                                      r * w.nrm[2] < 0.5 * CF_RES_D_MAX && fabs(w.pm_max - CF_TAU_MIN) < 0.5 * CF_RES_M_MAX;
                    do_factor = !(conv || kk + 1 >= itmax);
                    redo = false;
                    st = ST_RF;
                }
            }
        } else {
            CF_PROF_BEGIN();
            w.backward_rhs(brm, sigma_mu);
            CF_PROF_END(CF_PROF_BWD);
            predictor = false;
            st = ST_FWD;
        }
    }
    iters = kk;
    if (!(fabs(w.alpha) <= 1.0) || !(fabs(w.mu) < 1e300)) w.flags |= CF_FLAG_NONFINITE;
    if (kk == itmax) return 1;
    if (w.alpha <= CF_ALPHA_MIN) return 2;
    if (w.mu != w.mu) return 3;
    return 0;
}

// The whole RTI step for instance `inst` (what acados_solve() does, ocp_nlp_sqp_rti.c:1232-1237).
// `sm` must be 16-byte aligned and its two mbarriers initialised (cf_warp_init_smem).
CF_DEV void cf_warp_init_smem(double *sm)
{
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + CF_SM_BAR);
    if (cf_lane() == 0) { cf_mbar_init(bar); cf_mbar_init(bar + 1); }
    cf_syncwarp();
}

template <int PH = CF_PH_BOTH, bool VDT = false>
CF_DEV void cf_rti_instance(const CfParams *Pg, const CfBatchView &bv, int inst, double *slot, double *sm, unsigned &par)
{
    typedef CfWarpT<PH, VDT> CfWarp;
    // this instance's parameters: the solver-wide set, overridden by whatever per-instance arrays the caller gave
    CfParams *P = reinterpret_cast<CfParams *>(sm + CF_SM_PAR);
    {
        const int lane = cf_lane();
        const double *src = reinterpret_cast<const double *>(Pg);
        double *dst = sm + CF_SM_PAR;
        if (lane < CF_PAR_DOUBLES) dst[lane] = src[lane];
        if (lane + 32 < CF_PAR_DOUBLES) dst[lane + 32] = src[lane + 32];
        cf_syncwarp();
        if (bv.W_b && lane < CF_NY) P->Wdiag[lane] = bv.W_b[(long) inst * CF_NY + lane];
        if (bv.WN_b && lane < CF_NX) P->WNdiag[lane] = bv.WN_b[(long) inst * CF_NX + lane];
        if (lane < CF_NU) {
            if (bv.lbu_b) P->lbu[lane] = P->lbu0[lane] = bv.lbu_b[(long) inst * CF_NU + lane];
            if (bv.ubu_b) P->ubu[lane] = P->ubu0[lane] = bv.ubu_b[(long) inst * CF_NU + lane];
            if (bv.lbu0_b) P->lbu0[lane] = bv.lbu0_b[(long) inst * CF_NU + lane];
            if (bv.ubu0_b) P->ubu0[lane] = bv.ubu0_b[(long) inst * CF_NU + lane];
        }
        cf_syncwarp();
    }
    CfWarp w;
    w.bind(P, Pg, slot, sm, (PH != CF_PH_BOTH) ? bv.prep + (long) inst * bv.prep_stride : nullptr, VDT ? bv.dts : nullptr);
    w.par = par;
    w.BST = bv.bnd_stage;
    w.set_stage_weights(bv.W_stage);
    if (PH != CF_PH_PREPARATION && bv.mult) w.A0S = bv.mult + (long) inst * bv.mult_stride + (long) Pg->N * 29 + CF_NX;
    const int N = Pg->N;
    double *xg = bv.x + (long) inst * (N + 1) * CF_NX;
    double *ug = bv.u + (long) inst * N * CF_NU;
    const double *x0g = bv.x0 + (long) inst * CF_NX;
    const double *yrefg = bv.yref + (long) inst * N * CF_NY;
    const double *yref_eg = bv.yref_e + (long) inst * CF_NX;
    unsigned long long *prof = bv.prof;
    if constexpr (PH != CF_PH_FEEDBACK) {
        CF_PROF_BEGIN();
        w.nominal_pass(xg, ug);
        CF_NOUNROLL
        for (int k = 0; k < N; k++) w.linearize_stage(k, xg, x0g, yrefg);
        w.terminal_gradient(xg, yref_eg);
        cf_syncwarp();
        CF_PROF_END(CF_PROF_LIN);
    } else {
        w.load_prepared(xg, ug, x0g);
        cf_syncwarp();
    }
    if constexpr (PH == CF_PH_PREPARATION) {   // the solution, status and statistics of the instance are left as they are
        par = w.par;
        return;
    } else {
    int iters = 0;
    const int qp_status = cf_ipm_solve(w, iters, prof);
    CF_PROF_BEGIN();
    // ocp_nlp_sqp_rti.c:651-674: QP max-iter is not fatal; anything else leaves the iterate untouched
    int status = CF_ACADOS_SUCCESS;
    if (qp_status == 0 || qp_status == 1) {
        // primal update, full step (ocp_nlp_common.c:2900-2952); x_0 takes the eliminated step xbar.  Four stages per
        // trip so that the (independent) step loads of several stages are in flight together.
        const int lane = w.lane;
        const bool ul = lane < CF_NU, xl = lane >= CF_NU && lane < CF_NV;
        const int i = xl ? lane - CF_NU : 0;
        if (bv.mult) {
            // full-step duals (ocp_nlp_common.c:2917-2925): the multipliers of the QP solution, and for the eliminated
            // x_0 = x0 what OCP_QP_RESTORE_EQ_DOF recovers from the stationarity condition of the unreduced stage 0
            // (x_ocp_qp_red.c:820-840): q_0 + Q_0 xbar + A_0' pi_0
            double *mo = bv.mult + (long) inst * bv.mult_stride;
            double *pi_o = mo, *lam_o = mo + (long) N * CF_NX, *t_o = lam_o + (long) N * 8, *l0 = t_o + (long) N * 8;
            CF_NOUNROLL
            for (int k = 0; k < N; k++) {
                if (lane < CF_NX) pi_o[k * CF_NX + lane] = w.rec(k + 1)[R_PI + lane];
                if (lane < 8) { lam_o[k * 8 + lane] = w.rec(k)[R_LAM + lane]; t_o[k * 8 + lane] = w.rec(k)[R_T + lane]; }
            }
            if (xl && N > 0) {
                const double xbar = x0g[i] - xg[i];
                const double q0 = (w.wgt(0, i) * (xg[i] - yrefg[i])) * w.dt(0);
                const double *a0 = l0 + CF_NX;
                double s0 = 0.0;
                CF_NOUNROLL
                for (int c = 0; c < CF_NX; c++) s0 += a0[c * CF_NX + i] * w.rec(1)[R_PI + c];
                l0[i] = (q0 + w.hess(0) * xbar) + s0;
            }
        }
        if (xl) xg[i] += x0g[i] - xg[i];
        if (ul && N > 0) ug[lane] += w.rec(0)[R_UX + lane];
        CF_NOUNROLL
        for (int k = 1; k <= N; k += 4) {
            double d[4], v[4];
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                const bool on = k + q <= N && ((ul && k + q < N) || xl);
                d[q] = on ? w.rec(k + q)[R_UX + lane] : 0.0;
                v[q] = on ? (ul ? ug[(k + q) * CF_NU + lane] : xg[(k + q) * CF_NX + i]) : 0.0;
            }
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                if (k + q <= N) {
                    if (ul && k + q < N) ug[(k + q) * CF_NU + lane] = v[q] + d[q];
                    if (xl) xg[(k + q) * CF_NX + i] = v[q] + d[q];
                }
            }
        }
    } else {
        status = CF_ACADOS_QP_FAILURE;
    }
    if (w.lane == 0) {
        bv.status[inst] = status;
        bv.qp_iter[inst] = iters;
        bv.qp_status[inst] = qp_status;
        bv.flags[inst] = w.flags;
        if (bv.res) { for (int i = 0; i < 4; i++) bv.res[inst * 4 + i] = w.nrm[i]; }
    }
    par = w.par;
    cf_syncwarp();
    CF_PROF_END(CF_PROF_UPDATE);
    }   // PH != CF_PH_PREPARATION
}
