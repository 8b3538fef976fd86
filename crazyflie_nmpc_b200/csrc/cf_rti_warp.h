// One real-time-iteration SQP step of the Crazyflie OCP executed by ONE WARP for ONE
// problem instance.  Everything the reference does inside acados_solve() for this OCP:
//
//   preparation   ERK4 + forward sensitivities, Gauss-Newton gradient, bound residuals
//                 (acados/acados/ocp_nlp/ocp_nlp_sqp_rti.c:495-542,
//                  acados/acados/sim/sim_erk_integrator.c:658-731,
//                  acados/acados/ocp_nlp/ocp_nlp_cost_ls.c:810-916,
//                  acados/acados/ocp_nlp/ocp_nlp_constraints_bgh.c:1613-1648)
//   feedback      x0 elimination, Mehrotra predictor-corrector IPM on the stage-wise QP with a
//                 square-root Riccati factorisation, primal update
//                 (ocp_nlp_sqp_rti.c:545-683, external/hpipm/ocp_qp/x_ocp_qp_red.c:268-455,
//                  x_ocp_qp_ipm.c:1442-1774,1943-2759, x_ocp_qp_kkt.c:401-762,1108-1292,
//                  x_ocp_qp_res.c:334-637, ipm_core/x_core_qp_ipm_aux.c:36-457)
//
// Mapping.  Stage variables are [u(4); x(13)] -> lanes 0..16; lane 17 carries the extra
// "gradient / b" row of the (nv+1) x nv square-root Riccati blocks.  Lane r owns ROW r of
// the stage matrices it works on ([B';A';b'] is 18 x 13, the factor L is 18 x 17), so the
// dense per-stage kernels (TRMM, SYRK, Cholesky) are rank-1 register updates with the
// broadcast operand read from shared memory, and triangular/GEMV sweeps use either the row
// layout or a column-per-lane layout loaded from the same packed global block.  Lanes 18..31
// help only in the element-wise passes.
//
// All stages use the uniform nv = 17 layout: stage 0 keeps 13 decoupled dummy x-variables
// (its A-rows are zeroed after the x0 elimination folded A0*xbar into b0) and stage N keeps 4
// decoupled dummy inputs; both stay exactly zero and cost 2/51 of the work.
//
// The per-instance working set (linearisation [B';A';b'] 94 KB, factors 69 KB, IPM vectors)
// does not fit on chip; it lives in a per-warp scratch slot in global memory (L2/HBM).
#pragma once
#include "cf_model.h"

// ------------------------------------------------------------------ sizes / layout
#define CF_MROWS 18                       // rows of [B';A';res_b'] held per stage
#define CF_MSZ (CF_MROWS * CF_NX)         // 234 doubles, element (r,c) at c*18 + r
#define CF_LSZ 170                        // packed lower-trapezoid 18x17, column-major
#define CF_BND 64                         // doubles per stage of bound data: 8 fields x [lb4 | ub4]
enum { CF_F_D = 0, CF_F_LAM, CF_F_T, CF_F_RESD, CF_F_RESM, CF_F_BKP, CF_F_DLAM, CF_F_DT };

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

struct CfParams
{
    double Wdiag[CF_NY];   // stage weights, cost order y = [x;u] (generate_c_code.py:61-84)
    double WNdiag[CF_NX];  // terminal weights (:113)
    double lbu[CF_NU], ubu[CF_NU];
    double Ts;
    int N;
    int max_ipm_iter;      // CF_ITER_MAX unless a test truncates the loop
};

struct CfBatchView
{
    int B;
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
};

// offsets (in doubles) of the arrays inside one scratch slot
struct CfScratchLayout
{
    long M, L, b, rq, ux, pi, res_g, dux, dpi, Pb, bnd, total;
};
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    CfScratchLayout
    cf_scratch_layout(int N)
{
    CfScratchLayout s;
    long o = 0;
    s.M = o;     o += (long) N * CF_MSZ;
    s.L = o;     o += (long) (N + 1) * CF_LSZ;
    s.b = o;     o += (long) N * CF_NX + 1;
    s.rq = o;    o += (long) (N + 1) * CF_NV + 1;
    s.ux = o;    o += (long) (N + 1) * CF_NV + 1;
    s.pi = o;    o += (long) (N + 1) * CF_NX + 1;
    s.res_g = o; o += (long) (N + 1) * CF_NV + 1;
    s.dux = o;   o += (long) (N + 1) * CF_NV + 1;
    s.dpi = o;   o += (long) (N + 1) * CF_NX + 1;
    s.Pb = o;    o += (long) N * CF_NX + 1;
    s.bnd = o;   o += (long) (N + 1) * CF_BND;
    s.total = (o + 15) & ~15L;  // keep every slot 128-byte aligned
    return s;
}

// per-warp shared memory (doubles)
#define CF_SM_LROWS 0                          // 17 x 17 finished factor rows (stride 17)
#define CF_SM_AL (CF_SM_LROWS + 17 * 17 + 1)   // 18 x 14 AL rows (stride 14 -> 16-byte aligned rows)
#define CF_SM_V0 (CF_SM_AL + 18 * 14)          // small broadcast vectors, 32 each
#define CF_SM_V1 (CF_SM_V0 + 32)
#define CF_SM_V2 (CF_SM_V1 + 32)
#define CF_SM_DOUBLES (CF_SM_V2 + 32)          // 638 doubles = 5104 bytes per warp

CF_DEV int cf_loff(int c) { return 18 * c - (c * (c - 1)) / 2; }  // start of column c in a packed factor

struct CfWarp
{
    // ---- per-warp context
    const CfParams *P;
    int lane, N;
    double *sm;  // per-warp shared memory
    // scratch arrays
    double *M, *L, *b, *rq, *ux, *pi, *res_g, *dux, *dpi, *Pb, *bnd;
    // lane constants
    double Hs, HN;     // Hessian diagonal for this lane's variable (stage / terminal)
    // IPM scalars (warp-uniform)
    double mu, alpha, mu_aff, sigma;
    double nrm[4];     // inf-norms of res_g, res_b, res_d, res_m
    double lin[4];     // inf-norms of the linear-system residual of the last solve
    int flags;

    CF_MEM void bind(const CfParams *P_, double *slot, double *sm_)
    {
        P = P_; N = P_->N; sm = sm_; lane = cf_lane();
        CfScratchLayout s = cf_scratch_layout(N);
        M = slot + s.M; L = slot + s.L; b = slot + s.b; rq = slot + s.rq; ux = slot + s.ux; pi = slot + s.pi;
        res_g = slot + s.res_g; dux = slot + s.dux; dpi = slot + s.dpi; Pb = slot + s.Pb; bnd = slot + s.bnd;
        // hess = scaling * (sqrt(W))^2 : ocp_nlp_cost_ls.c:739-772 (terminal scaling stays 1.0, :265)
        double w = 1.0, wN = 1.0;
        if (lane < CF_NU) w = P->Wdiag[CF_NX + lane];
        else if (lane < CF_NV) { w = P->Wdiag[lane - CF_NU]; wN = P->WNdiag[lane - CF_NU]; }
        double r = sqrt(w), rN = sqrt(wN);
        Hs = P->Ts * (r * r);
        HN = (lane < CF_NU) ? Hs : (rN * rN);
    }

    // =============================================================== preparation
    // ERK4 with forward sensitivities for stage k; lane c pushes sensitivity column c
    // ([Su(4) | Sx(13)] -> rows of [B';A']), the nominal state is advanced once per warp in
    // shared memory.  Writes M_k (rows 0..16), b_k, rq_k, d_k.
    CF_MEM void linearize_stage(int k, const double *xg, const double *ug, const double *x0g,
                                const double *yrefg, const double *yref_eg)
    {
        double *X0 = sm + CF_SM_V0, *ACC = sm + CF_SM_V1, *XS = sm + CF_SM_V2, *UU = sm + CF_SM_V2 + 16;
        const double h = P->Ts;
        cf_syncwarp();
        if (lane < CF_NX) { double v = xg[k * CF_NX + lane]; X0[lane] = v; ACC[lane] = v; XS[lane] = v; }
        if (lane < CF_NU) UU[lane] = ug[k * CF_NU + lane];
        cf_syncwarp();
        double uu[CF_NU];
        CF_UNROLL
        for (int i = 0; i < CF_NU; i++) uu[i] = UU[i];
        double Ss[CF_NX], acc[CF_NX];
        CF_UNROLL
        for (int i = 0; i < CF_NX; i++) { Ss[i] = (lane - CF_NU == i) ? 1.0 : 0.0; acc[i] = Ss[i]; }
        CF_UNROLL
        for (int s = 0; s < 4; s++) {
            // tableau: sim_collocation_utils.c:611-640 (classic RK4)
            const double bw = (s == 0 || s == 3) ? (1.0 / 6.0) : (1.0 / 3.0);
            const double a_next = (s == 2) ? 1.0 : 0.5;
            double xs[CF_NX], f[CF_NX], ks[CF_NX];
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) xs[i] = XS[i];
            cf_ode(xs, uu, f);
            cf_jvp_x(xs, Ss, ks);
            if (lane < CF_NU) cf_add_ju_col(uu, lane, ks);
            cf_syncwarp();  // everyone has read XS
            const double bh = h * bw, ah = a_next * h;
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) {
                acc[i] += bh * ks[i];
                if (s < 3) Ss[i] = ((lane - CF_NU == i) ? 1.0 : 0.0) + ah * ks[i];
            }
            if (lane == 17) {
                CF_UNROLL
                for (int i = 0; i < CF_NX; i++) {
                    ACC[i] += bh * f[i];
                    if (s < 3) XS[i] = X0[i] + ah * f[i];
                }
            }
            cf_syncwarp();
        }
        // lane 17: b_k = phi(x_k,u_k) - x_{k+1}   (ocp_nlp_dynamics_cont.c:822-823)
        if (lane == 17) {
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) acc[i] = ACC[i] - xg[(k + 1) * CF_NX + i];
        }
        if (k == 0) {
            // x0 elimination (x_ocp_qp_red.c:310-330): xbar = lbx - x_0 ; b_0 += A_0 xbar ; drop the A rows
            double xbar = (lane >= CF_NU && lane < CF_NV) ? (x0g[lane - CF_NU] - xg[lane - CF_NU]) : 0.0;
            CF_UNROLL
            for (int i = 0; i < CF_NX; i++) {
                double contrib = (lane >= CF_NU && lane < CF_NV) ? acc[i] * xbar : 0.0;
                double tot = cf_warp_sum(contrib);
                if (lane == 17) acc[i] = tot + acc[i];
                else if (lane >= CF_NU) acc[i] = 0.0;
            }
        }
        double *Mk = M + (long) k * CF_MSZ;
        if (lane < CF_MROWS) {
            CF_UNROLL
            for (int c = 0; c < CF_NX; c++) Mk[c * CF_MROWS + lane] = acc[c];
        }
        if (lane == 17) {
            CF_UNROLL
            for (int c = 0; c < CF_NX; c++) b[k * CF_NX + c] = acc[c];
        }
        // gradient: scaling * W * (y - yref), [u;x] order (ocp_nlp_cost_ls.c:883-912)
        if (lane < CF_NV) {
            double g;
            if (lane < CF_NU) g = (P->Wdiag[CF_NX + lane] * (UU[lane] - yrefg[k * CF_NY + CF_NX + lane])) * h;
            else g = (k == 0) ? 0.0 : (P->Wdiag[lane - CF_NU] * (X0[lane - CF_NU] - yrefg[k * CF_NY + lane - CF_NU])) * h;
            rq[k * CF_NV + lane] = g;
        }
        // bounds (ocp_nlp_constraints_bgh.c:1634-1636): d = [lb - u ; u - ub]
        if (lane < CF_NU) {
            double *bk = bnd + (long) k * CF_BND;
            bk[CF_F_D * 8 + lane] = P->lbu[lane] - UU[lane];
            bk[CF_F_D * 8 + 4 + lane] = UU[lane] - P->ubu[lane];
        }
        (void) yref_eg;
    }

    CF_MEM void terminal_gradient(const double *xg, const double *yref_eg)
    {
        if (lane < CF_NV) {
            double g = 0.0;
            if (lane >= CF_NU) g = P->WNdiag[lane - CF_NU] * (xg[N * CF_NX + lane - CF_NU] - yref_eg[lane - CF_NU]);
            rq[N * CF_NV + lane] = g;
        }
    }

    // =============================================================== IPM pieces
    // OCP_QP_INIT_VAR scheme 1 (x_ocp_qp_ipm.c:1491-1530,1636-1769)
    CF_MEM void init_var()
    {
        for (int k = 0; k <= N; k++) {
            double v = 0.0;
            if (lane < CF_NU && k < N) {
                double *bk = bnd + (long) k * CF_BND;
                double dl = bk[CF_F_D * 8 + lane], du = bk[CF_F_D * 8 + 4 + lane];
                double tl = -dl, tu = -du;
                if (tl < CF_THR0) {
                    if (tu < CF_THR0) { v = 0.5 * (dl - du); tl = CF_THR0; tu = CF_THR0; }
                    else { tl = CF_THR0; v = dl + CF_THR0; }
                } else if (tu < CF_THR0) { tu = CF_THR0; v = -du - CF_THR0; }
                bk[CF_F_T * 8 + lane] = tl; bk[CF_F_T * 8 + 4 + lane] = tu;
                bk[CF_F_LAM * 8 + lane] = CF_MU0 / tl; bk[CF_F_LAM * 8 + 4 + lane] = CF_MU0 / tu;
            }
            if (lane < CF_NV) ux[k * CF_NV + lane] = v;
            if (lane < CF_NX) pi[k * CF_NX + lane] = 0.0;
        }
    }

    // UPDATE_VAR_QP (x_core_qp_ipm_aux.c:220-325) fused with OCP_QP_RES_COMPUTE +
    // _INF_NORM (x_ocp_qp_res.c:334-470,602-637): one forward pass over the stages.
    CF_MEM void update_and_residuals(bool do_update)
    {
        double a = alpha;
        if (do_update && a < 1.0) a = a * ((1.0 - a) * 0.99 + a * 0.9999999);
        double ng = 0, nb = 0, nd = 0, nm = 0, mus = 0;
        double *UXS = sm + CF_SM_V0, *PIS = sm + CF_SM_V1;
        // prologue: ux_0
        double uxc = 0.0;
        if (lane < CF_NV) {
            uxc = ux[lane];
            if (do_update) { uxc += a * dux[lane]; ux[lane] = uxc; }
        }
        double pi_prev = 0.0;
        for (int k = 0; k <= N; k++) {
            double uxn = 0.0, pik = 0.0;
            if (k < N) {
                if (lane < CF_NV) {
                    uxn = ux[(k + 1) * CF_NV + lane];
                    if (do_update) { uxn += a * dux[(k + 1) * CF_NV + lane]; ux[(k + 1) * CF_NV + lane] = uxn; }
                }
                if (lane >= CF_NU && lane < CF_NV) {
                    pik = pi[k * CF_NX + lane - CF_NU];
                    if (do_update) { pik += a * dpi[k * CF_NX + lane - CF_NU]; pi[k * CF_NX + lane - CF_NU] = pik; }
                }
            }
            double rg = 0.0;
            if (lane < CF_NV) {
                rg = ((k == N) ? HN : Hs) * uxc + rq[k * CF_NV + lane];
                if (k > 0 && lane >= CF_NU) rg -= pi_prev;
            }
            if (lane < CF_NU && k < N) {
                double *bk = bnd + (long) k * CF_BND;
                double ll = bk[CF_F_LAM * 8 + lane], lu = bk[CF_F_LAM * 8 + 4 + lane];
                double tl = bk[CF_F_T * 8 + lane], tu = bk[CF_F_T * 8 + 4 + lane];
                if (do_update) {
                    ll += a * bk[CF_F_DLAM * 8 + lane]; lu += a * bk[CF_F_DLAM * 8 + 4 + lane];
                    tl += a * bk[CF_F_DT * 8 + lane]; tu += a * bk[CF_F_DT * 8 + 4 + lane];
                    ll = ll <= CF_LAM_MIN ? CF_LAM_MIN : ll; lu = lu <= CF_LAM_MIN ? CF_LAM_MIN : lu;
                    tl = tl <= CF_T_MIN ? CF_T_MIN : tl; tu = tu <= CF_T_MIN ? CF_T_MIN : tu;
                    bk[CF_F_LAM * 8 + lane] = ll; bk[CF_F_LAM * 8 + 4 + lane] = lu;
                    bk[CF_F_T * 8 + lane] = tl; bk[CF_F_T * 8 + 4 + lane] = tu;
                }
                rg += lu - ll;
                double rdl = bk[CF_F_D * 8 + lane] + tl - uxc, rdu = bk[CF_F_D * 8 + 4 + lane] + tu + uxc;
                double rml = ll * tl, rmu = lu * tu;
                bk[CF_F_RESD * 8 + lane] = rdl; bk[CF_F_RESD * 8 + 4 + lane] = rdu;
                bk[CF_F_BKP * 8 + lane] = rml; bk[CF_F_BKP * 8 + 4 + lane] = rmu;
                mus += rml + rmu;
                nd = fmax(nd, fmax(fabs(rdl), fabs(rdu)));
                nm = fmax(nm, fmax(fabs(rml), fabs(rmu)));
            }
            if (k < N) {
                cf_syncwarp();
                if (lane < CF_NV) UXS[lane] = uxc;
                if (lane >= CF_NU && lane < CF_NV) PIS[lane - CF_NU] = pik;
                cf_syncwarp();
                double *Mk = M + (long) k * CF_MSZ;
                if (lane < CF_NV) {  // res_g += [B';A'] pi_k   (row layout)
                    double s = 0.0;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) s += Mk[c * CF_MROWS + lane] * PIS[c];
                    rg += s;
                }
                if (lane >= CF_NU && lane < CF_NV) {  // res_b = b - x+ + [A B] ux   (column layout)
                    const int c = lane - CF_NU;
                    double s = 0.0;
                    CF_UNROLL
                    for (int r = 0; r < CF_NV; r++) s += Mk[c * CF_MROWS + r] * UXS[r];
                    double rb = (b[k * CF_NX + c] - uxn) + s;
                    nb = fmax(nb, fabs(rb));
                    Mk[c * CF_MROWS + 17] = rb;  // ROWIN(res_b) of x_ocp_qp_kkt.c:490
                }
            }
            if (lane < CF_NV) { res_g[k * CF_NV + lane] = rg; ng = fmax(ng, fabs(rg)); }
            pi_prev = pik;
            uxc = uxn;
        }
        nrm[0] = cf_warp_max(ng); nrm[1] = cf_warp_max(nb); nrm[2] = cf_warp_max(nd); nrm[3] = cf_warp_max(nm);
        mu = cf_warp_sum(mus) * (1.0 / (double) (2 * CF_NU * N));
        cf_syncwarp();
    }

    // gamma for the condensed right-hand side (x_core_qp_ipm_aux.c:38-111); lanes 0..3 of stage k<N.
    // rm_mode: 0 predictor (bkp - tau_min), 1 corrector (bkp + dt*dlam - sigma_mu, stored),
    //          2 re-centering (bkp - sigma_mu, stored), 3 use the stored RESM as is.
    CF_MEM void bound_terms(int k, int rm_mode, double sigma_mu, double &Gam, double &gam)
    {
        double *bk = bnd + (long) k * CF_BND;
        double ll = bk[CF_F_LAM * 8 + lane], lu = bk[CF_F_LAM * 8 + 4 + lane];
        double til = 1.0 / bk[CF_F_T * 8 + lane], tiu = 1.0 / bk[CF_F_T * 8 + 4 + lane];
        double rml, rmu;
        if (rm_mode == 3) { rml = bk[CF_F_RESM * 8 + lane]; rmu = bk[CF_F_RESM * 8 + 4 + lane]; }
        else {
            rml = bk[CF_F_BKP * 8 + lane]; rmu = bk[CF_F_BKP * 8 + 4 + lane];
            if (rm_mode == 0) { rml -= CF_TAU_MIN; rmu -= CF_TAU_MIN; }
            else if (rm_mode == 1) {
                rml = rml + bk[CF_F_DT * 8 + lane] * bk[CF_F_DLAM * 8 + lane] - sigma_mu;
                rmu = rmu + bk[CF_F_DT * 8 + 4 + lane] * bk[CF_F_DLAM * 8 + 4 + lane] - sigma_mu;
            } else { rml -= sigma_mu; rmu -= sigma_mu; }
            bk[CF_F_RESM * 8 + lane] = rml; bk[CF_F_RESM * 8 + 4 + lane] = rmu;
        }
        double gl = til * (rml - ll * bk[CF_F_RESD * 8 + lane]);
        double gu = tiu * (rmu - lu * bk[CF_F_RESD * 8 + 4 + lane]);
        Gam = til * ll + tiu * lu;
        gam = gl - gu;
    }

    // OCP_QP_FACT_SOLVE_KKT_STEP, backward factorisation (x_ocp_qp_kkt.c:445-528):
    //   L_k = chol( [H_k + Gamma ; (res_g + gamma)'] + AL AL' ),  AL = [B';A';res_b']_k Lxx_{k+1}
    CF_MEM void factorize()
    {
        double *LS = sm + CF_SM_LROWS, *ALS = sm + CF_SM_AL, *V = sm + CF_SM_V0;
        double Lp[CF_NV];  // this lane's row of L_{k+1}
        CF_UNROLL
        for (int c = 0; c < CF_NV; c++) Lp[c] = 0.0;
        for (int k = N; k >= 0; k--) {
            double s[CF_NV];
            CF_UNROLL
            for (int j = 0; j < CF_NV; j++) s[j] = 0.0;
            if (k < N) {
                const double *Mk = M + (long) k * CF_MSZ;
                double m[CF_NX];
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) m[c] = (lane < CF_MROWS) ? Mk[c * CF_MROWS + lane] : 0.0;
                // TRMM_RLNN in place: m[c] <- sum_{j>=c} m[j] * Lxx[j][c]
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) {
                    double v = 0.0;
                    CF_UNROLL
                    for (int j = c; j < CF_NX; j++) v += m[j] * LS[(CF_NU + j) * 17 + CF_NU + c];
                    m[c] = v;
                }
                cf_syncwarp();
                if (lane == 17) {
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) V[c] = m[c];
                }
                cf_syncwarp();
                // Pb = Lxx * (Lxx' res_b)  (TRMV_LNN, :492-493): lane 4+r uses its own factor row
                if (lane >= CF_NU && lane < CF_NV) {
                    double v = 0.0;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++)
                        if (c <= lane - CF_NU) v += Lp[CF_NU + c] * V[c];
                    Pb[k * CF_NX + lane - CF_NU] = v;
                }
                if (lane == 17) {
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) m[c] += Lp[CF_NU + c];
                }
                if (lane < CF_MROWS) {
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) ALS[lane * 14 + c] = m[c];
                }
                cf_syncwarp();
                // SYRK: s[j] = sum_c AL[r][c] * AL[j][c]
                CF_UNROLL
                for (int j = 0; j < CF_NV; j++) {
                    double v = 0.0;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) v += m[c] * ALS[j * 14 + c];
                    s[j] = v;
                }
            }
            // diagonal / gradient row
            cf_syncwarp();
            {
                double Gam = 0.0, gam = 0.0;
                if (lane < CF_NU && k < N) bound_terms(k, 0, 0.0, Gam, gam);
                double g = 0.0;
                if (lane < CF_NV) g = res_g[k * CF_NV + lane] + gam;
                V[lane] = g;
                const double hd = ((k == N) ? HN : Hs) + CF_REG_PRIM + Gam;
                CF_UNROLL
                for (int j = 0; j < CF_NV; j++)
                    if (lane == j) s[j] += hd;
            }
            cf_syncwarp();
            if (lane == 17) {
                CF_UNROLL
                for (int j = 0; j < CF_NV; j++) s[j] += V[j];
            }
            // right-looking Cholesky of the 18 x 17 block, one column per step
            // (POTRF_L_MN; non-positive pivot -> 0, BLASFEO kernel_dgemm_4x4_lib4.c:5701-5714)
            CF_UNROLL
            for (int j = 0; j < CF_NV; j++) {
                const double piv = cf_shfl(s[j], j);
                double dj = 0.0, inv = 0.0;
                if (piv > 0.0) { dj = sqrt(piv); inv = 1.0 / dj; }
                s[j] = (lane == j) ? dj : ((lane > j) ? s[j] * inv : 0.0);
                if (lane < CF_NV) LS[lane * 17 + j] = s[j];
                cf_syncwarp();
                CF_UNROLL
                for (int jj = j + 1; jj < CF_NV; jj++) s[jj] -= s[j] * LS[jj * 17 + j];
            }
            // store packed factor (column-major lower trapezoid)
            double *Lk = L + (long) k * CF_LSZ;
            if (lane < CF_MROWS) {
                CF_UNROLL
                for (int c = 0; c < CF_NV; c++)
                    if (lane >= c) Lk[cf_loff(c) + lane - c] = s[c];
            }
            CF_UNROLL
            for (int c = 0; c < CF_NV; c++) Lp[c] = s[c];
        }
        cf_syncwarp();
    }

    // Forward substitution shared by the factorise-and-solve (mode 0, :536-570) and the
    // rhs-only solve (mode 1, :1250-1290); computes dux, dpi, then dlam, dt (:741-758,
    // x_core_qp_ipm_aux.c:117-142), the step length ingredients (:146-216) and the inf-norms of
    // the linear-system residual (OCP_QP_RES_COMPUTE_LIN, x_ocp_qp_res.c:474-598) on the fly.
    // rm_mode selects which complementarity rhs the step was computed for (see bound_terms).
    CF_MEM void forward(int mode, int rm_mode)
    {
        double *XS = sm + CF_SM_V0, *DS = sm + CF_SM_V1, *YS = sm + CF_SM_V2;
        double a_p = -1.0, a_d = -1.0;            // running alpha_prim / alpha_dual (negated)
        double lg = 0, lb = 0, ld = 0, lm = 0;    // linear residual norms
        // column layout of L_0
        double col[CF_MROWS];
        load_factor_cols(0, col);
        double dxk = 0.0;        // lanes 4..16: dx_k ; stage 0 has none
        double dpi_prev = 0.0;   // lanes 4..16: dpi_{k-1}
        cf_syncwarp();
        if (lane < 32) XS[lane] = 0.0;
        cf_syncwarp();
        for (int k = 0; k <= N; k++) {
            // ---- u-part: du = Luu^-T ( -l_u - Lxu' dx )      TRSV_LTN_MN(nv, nu)
            double v = 0.0;
            if (lane < CF_NU) {
                if (mode == 1) v = -dux[k * CF_NV + lane];
                CF_UNROLL
                for (int t = 1; t < CF_MROWS; t++) {
                    const int i = lane + t;
                    if (i >= CF_NU && i < CF_NV) v -= col[t] * XS[i - CF_NU];
                    if (mode == 0 && i == 17) v -= col[t];  // - l~_u (last row of L_k)
                }
            }
            const double invd = 1.0 / col[0];
            double du = 0.0;
            CF_UNROLL
            for (int j = CF_NU - 1; j >= 0; j--) {
                const double duj = cf_shfl(v * invd, j);
                if (lane == j) du = duj;
                if (lane < j) {
                    const int t = j - lane;
                    const double lj = (t == 1) ? col[1] : ((t == 2) ? col[2] : col[3]);
                    v -= lj * duj;
                }
            }
            const double duxk = (lane < CF_NU) ? du : dxk;  // lane r: dux_k[r]
            if (lane < CF_NV) dux[k * CF_NV + lane] = duxk;
            // ---- dlam, dt, alpha (lanes 0..3)
            double dlam_l = 0, dlam_u = 0;
            if (lane < CF_NU && k < N) {
                double *bk = bnd + (long) k * CF_BND;
                double ll = bk[CF_F_LAM * 8 + lane], lu = bk[CF_F_LAM * 8 + 4 + lane];
                double tl = bk[CF_F_T * 8 + lane], tu = bk[CF_F_T * 8 + 4 + lane];
                double til = 1.0 / tl, tiu = 1.0 / tu;
                double rdl = bk[CF_F_RESD * 8 + lane], rdu = bk[CF_F_RESD * 8 + 4 + lane];
                double rml, rmu;
                if (rm_mode == 0) { rml = bk[CF_F_BKP * 8 + lane] - CF_TAU_MIN; rmu = bk[CF_F_BKP * 8 + 4 + lane] - CF_TAU_MIN; }
                else { rml = bk[CF_F_RESM * 8 + lane]; rmu = bk[CF_F_RESM * 8 + 4 + lane]; }
                double dtl = du, dtu = -du;
                dlam_l = -til * (rml + (ll * dtl) - (ll * rdl));
                dlam_u = -tiu * (rmu + (lu * dtu) - (lu * rdu));
                dtl -= rdl; dtu -= rdu;
                bk[CF_F_DLAM * 8 + lane] = dlam_l; bk[CF_F_DLAM * 8 + 4 + lane] = dlam_u;
                bk[CF_F_DT * 8 + lane] = dtl; bk[CF_F_DT * 8 + 4 + lane] = dtu;
                if (a_d * dlam_l > ll) a_d = ll / dlam_l;
                if (a_p * dtl > tl) a_p = tl / dtl;
                if (a_d * dlam_u > lu) a_d = lu / dlam_u;
                if (a_p * dtu > tu) a_p = tu / dtu;
                // linear residuals of the complementarity / bound rows
                ld = fmax(ld, fmax(fabs(rdl + dtl - du), fabs(rdu + dtu + du)));
                lm = fmax(lm, fmax(fabs(rml + ll * dtl + dlam_l * tl), fabs(rmu + lu * dtu + dlam_u * tu)));
            }
            // stationarity residual, part 1: H dux + rhs_g - dpi_{k-1} + dlam_ub - dlam_lb
            double rgl = 0.0;
            if (lane < CF_NV) {
                rgl = ((k == N) ? HN : Hs) * duxk + res_g[k * CF_NV + lane];
                if (k > 0 && lane >= CF_NU) rgl -= dpi_prev;
                rgl += dlam_u - dlam_l;
            }
            if (k == N) { if (lane < CF_NV) lg = fmax(lg, fabs(rgl)); break; }
            // ---- dx+ = [A B] dux + res_b        GEMV_T, column layout of M_k
            cf_syncwarp();
            if (lane < CF_NV) DS[lane] = duxk;
            cf_syncwarp();
            const double *Mk = M + (long) k * CF_MSZ;
            double dxn = 0.0, rbk = 0.0;
            if (lane >= CF_NU && lane < CF_NV) {
                const int c = lane - CF_NU;
                double sacc = 0.0;
                CF_UNROLL
                for (int r = 0; r < CF_NV; r++) sacc += Mk[c * CF_MROWS + r] * DS[r];
                rbk = Mk[c * CF_MROWS + 17];
                dxn = sacc + rbk;
                lb = fmax(lb, fabs((rbk - dxn) + sacc));
            }
            cf_syncwarp();
            if (lane >= CF_NU && lane < CF_NV) XS[lane - CF_NU] = dxn;
            cf_syncwarp();
            // ---- dpi: needs L_{k+1} in column layout (Lxx' dx+) and row layout (Lxx * .)
            double coln[CF_MROWS];
            load_factor_cols(k + 1, coln);
            double y = 0.0;
            if (lane >= CF_NU && lane < CF_NV) {
                CF_UNROLL
                for (int t = 0; t < CF_MROWS; t++) {
                    const int i = lane + t;
                    if (i < CF_NV) y += coln[t] * XS[i - CF_NU];
                }
                if (mode == 0) {  // + l~_x  (last row of L_{k+1})
                    CF_UNROLL
                    for (int t = 1; t < CF_MROWS; t++)
                        if (lane + t == 17) y += coln[t];
                }
                YS[lane - CF_NU] = y;
            }
            cf_syncwarp();
            double dpik = 0.0;
            if (lane >= CF_NU && lane < CF_NV) {
                const double *Ln = L + (long) (k + 1) * CF_LSZ;
                double z = 0.0;
                CF_UNROLL
                for (int c = CF_NU; c < CF_NV; c++)
                    if (lane >= c) z += Ln[cf_loff(c) + lane - c] * YS[c - CF_NU];
                if (mode == 1) z += dux[(k + 1) * CF_NV + lane];  // p_{k+1} from the backward sweep
                dpik = z;
                dpi[k * CF_NX + lane - CF_NU] = dpik;
            }
            // stationarity residual, part 2: + [B';A'] dpi_k   (row layout)
            cf_syncwarp();
            if (lane >= CF_NU && lane < CF_NV) YS[lane - CF_NU] = dpik;
            cf_syncwarp();
            if (lane < CF_NV) {
                double sacc = 0.0;
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) sacc += Mk[c * CF_MROWS + lane] * YS[c];
                lg = fmax(lg, fabs(rgl + sacc));
            }
            dpi_prev = dpik;
            dxk = dxn;
            CF_UNROLL
            for (int t = 0; t < CF_MROWS; t++) col[t] = coln[t];
        }
        lin[0] = cf_warp_max(lg); lin[1] = cf_warp_max(lb); lin[2] = cf_warp_max(ld); lin[3] = cf_warp_max(lm);
        a_p = cf_warp_max(a_p); a_d = cf_warp_max(a_d);
        alpha = -(a_p > a_d ? a_p : a_d);
        cf_syncwarp();
    }

    // lane c (0..16) loads column c of the packed factor of stage k: col[t] = L[c+t][c]
    CF_MEM void load_factor_cols(int k, double *col)
    {
        const double *Lk = L + (long) k * CF_LSZ;
        const int off = cf_loff(lane < CF_NV ? lane : 0);
        CF_UNROLL
        for (int t = 0; t < CF_MROWS; t++) col[t] = (lane < CF_NV && lane + t < CF_MROWS) ? Lk[off + t] : ((t == 0) ? 1.0 : 0.0);
    }

    // OCP_QP_SOLVE_KKT_STEP, backward vector recursion with cached Pb (x_ocp_qp_kkt.c:1147-1245):
    // leaves l_k = [L^-1 rhs]_u ; p_k in dux_k for the forward sweep.
    CF_MEM void backward_rhs(int rm_mode, double sigma_mu)
    {
        double *TS = sm + CF_SM_V0;
        double pn = 0.0;  // lanes 4..16: p_{k+1}
        for (int k = N; k >= 0; k--) {
            double Gam = 0.0, gam = 0.0;
            if (lane < CF_NU && k < N) bound_terms(k, rm_mode, sigma_mu, Gam, gam);
            double rhs = 0.0;
            if (lane < CF_NV) rhs = res_g[k * CF_NV + lane] + gam;
            if (k < N) {
                cf_syncwarp();
                if (lane >= CF_NU && lane < CF_NV) TS[lane - CF_NU] = pn + Pb[k * CF_NX + lane - CF_NU];
                cf_syncwarp();
                const double *Mk = M + (long) k * CF_MSZ;
                if (lane < CF_NV) {
                    double sacc = 0.0;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) sacc += Mk[c * CF_MROWS + lane] * TS[c];
                    rhs += sacc;
                }
            }
            // TRSV_LNN_MN(nv, nu): row layout of the 4 input columns of L_k
            const double *Lk = L + (long) k * CF_LSZ;
            double Lr[CF_NU];
            CF_UNROLL
            for (int j = 0; j < CF_NU; j++) Lr[j] = (lane < CF_NV && lane >= j) ? Lk[cf_loff(j) + lane - j] : 0.0;
            const double dg = (lane == 0) ? Lr[0] : ((lane == 1) ? Lr[1] : ((lane == 2) ? Lr[2] : ((lane == 3) ? Lr[3] : 1.0)));
            const double invd = 1.0 / dg;
            CF_UNROLL
            for (int j = 0; j < CF_NU; j++) {
                const double zj = cf_shfl(rhs * invd, j);
                if (lane == j) rhs = zj;
                else if (lane > j) rhs -= Lr[j] * zj;
            }
            if (lane < CF_NV) dux[k * CF_NV + lane] = rhs;
            pn = rhs;
        }
        cf_syncwarp();
    }

    // COMPUTE_MU_AFF_QP (x_core_qp_ipm_aux.c:329-352): all 32 lanes sweep the bound records
    CF_MEM void compute_mu_aff()
    {
        double s = 0.0;
        const int e = lane & 7;
        for (int k = lane >> 3; k < N; k += 4) {
            const double *bk = bnd + (long) k * CF_BND;
            s += (bk[CF_F_LAM * 8 + e] + alpha * bk[CF_F_DLAM * 8 + e]) * (bk[CF_F_T * 8 + e] + alpha * bk[CF_F_DT * 8 + e]);
        }
        mu_aff = cf_warp_sum(s) * (1.0 / (double) (2 * CF_NU * N));
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

    // OCP_QP_IPM_SOLVE, delta formulation (x_ocp_qp_ipm.c:2409-2759). Returns HPIPM status.
    CF_MEM int ipm_solve(int &iters)
    {
        init_var();
        alpha = 1.0;
        flags = 0;
        cf_syncwarp();
        update_and_residuals(false);
        int kk = 0;
        const int itmax = P->max_ipm_iter < CF_ITER_MAX ? P->max_ipm_iter : CF_ITER_MAX;
        for (; kk < itmax && alpha > CF_ALPHA_MIN &&
               (nrm[0] > CF_RES_G_MAX || nrm[1] > CF_RES_B_MAX || nrm[2] > CF_RES_D_MAX ||
                fabs(nrm[3] - CF_TAU_MIN) > CF_RES_M_MAX);
             kk++) {
            // ---- OCP_QP_IPM_DELTA_STEP (:1943-2405)
            factorize();
            forward(0, 0);
            if (!lin_res_ok_fact()) flags |= CF_FLAG_LIN_RES_FACT;
            compute_mu_aff();
            const double tmp = mu_aff / mu;
            sigma = tmp * tmp * tmp;
            double sigma_mu = sigma * mu;
            sigma_mu = sigma_mu > CF_TAU_MIN ? sigma_mu : CF_TAU_MIN;
            backward_rhs(1, sigma_mu);
            forward(1, 3);
            // conditional predictor-corrector (:2230-2273)
            const double mu_aff0 = mu_aff;
            compute_mu_aff();
            if (mu_aff > 2.0 * mu_aff0) {
                backward_rhs(2, sigma_mu);
                forward(1, 3);
            }
            if (!lin_res_ok_corr()) flags |= CF_FLAG_LIN_RES_CORR;
            update_and_residuals(true);
        }
        iters = kk;
        if (kk == itmax) return 1;
        if (alpha <= CF_ALPHA_MIN) return 2;
        if (mu != mu) return 3;
        return 0;
    }
};

// The whole RTI step for instance `inst` (what acados_solve() does, ocp_nlp_sqp_rti.c:1232-1237).
CF_DEV void cf_rti_instance(const CfParams *P, const CfBatchView &bv, int inst, double *slot, double *sm)
{
    CfWarp w;
    w.bind(P, slot, sm);
    const int N = P->N;
    double *xg = bv.x + (long) inst * (N + 1) * CF_NX;
    double *ug = bv.u + (long) inst * N * CF_NU;
    const double *x0g = bv.x0 + (long) inst * CF_NX;
    const double *yrefg = bv.yref + (long) inst * N * CF_NY;
    const double *yref_eg = bv.yref_e + (long) inst * CF_NX;
    for (int k = 0; k < N; k++) w.linearize_stage(k, xg, ug, x0g, yrefg, yref_eg);
    w.terminal_gradient(xg, yref_eg);
    cf_syncwarp();
    int iters = 0;
    const int qp_status = w.ipm_solve(iters);
    // ocp_nlp_sqp_rti.c:651-674: QP max-iter is not fatal; anything else leaves the iterate untouched
    int status = CF_ACADOS_SUCCESS;
    if (qp_status == 0 || qp_status == 1) {
        // primal update, full step (ocp_nlp_common.c:2900-2952); x_0 takes the eliminated step xbar
        const int lane = w.lane;
        for (int k = 0; k <= N; k++) {
            if (lane < CF_NU && k < N) ug[k * CF_NU + lane] += w.ux[k * CF_NV + lane];
            if (lane >= CF_NU && lane < CF_NV) {
                const int i = lane - CF_NU;
                if (k == 0) xg[i] += x0g[i] - xg[i];
                else xg[k * CF_NX + i] += w.ux[k * CF_NV + lane];
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
    cf_syncwarp();
}
