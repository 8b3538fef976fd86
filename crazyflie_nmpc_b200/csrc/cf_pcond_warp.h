// Feedback phase of one RTI step with PARTIAL CONDENSING, executed by one warp for one problem instance
// (the reference's qp_cond_N = N2 < N: acados/acados/ocp_qp/ocp_qp_partial_condensing.c:457-576 ->
//  external/hpipm/cond/x_part_cond.c:505-560 d_part_cond_qp_cond, :658-742 d_part_cond_qp_expand_sol, block sizes :36-54;
//  condensing routines x_cond_aux.c:36-100 COND_BABT, :208-470 COND_RSQRQ, :840 COND_DCTD, :1820 EXPAND_SOL).
//
// The N shooting intervals are grouped into N2 blocks of BS (or BS-1) consecutive stages.  A block [k0, k0+bs) becomes
// ONE stage of the condensed QP with state xbar = x_k0 and inputs ubar = [u_{k0+bs-1}; ...; u_k0] (last stage first, as
// HPIPM orders them); the states inside the block are eliminated through the dynamics:
//     x_{k0+j} = G_j' [ubar; xbar; 1],   G_0 = [0; I; 0],   G_{j+1} = G_j A_j' + [E_j B_j'; 0; b_j']
//     H2 = blkdiag(R_j.., Q_0) + sum_{j>=1} G_j Q_j G_j',   rq2 = [r_j..; q_0] + sum_{j>=1} G_j (q_j + Q_j c_j)
// (c_j = the constant row of G_j; only input boxes exist, so the inequality data are just re-ordered).  The interior-point
// method (the same state machine cf_ipm_solve as the uncondensed path, x_ocp_qp_ipm.c:2409-2759) then runs on N2 + 1 fat
// stages of nv = 4 BS + 13 variables with a DENSE Hessian; afterwards the inner states are recovered by the dynamics and
// the full step is applied.  Results agree with the reference run at the same qp_cond_N to ~1e-13
// (tests: oracle/cfnmpc_oracle.c cfo_rti_pcond pinned against the reference; tests/test_simt_emu.py, tests/test_gpu_pcond.py).
//
// Mapping.  Condensed stage variables [ubar(4 BS); xbar(13)] -> lanes 0..nv-1, lane nv carries the gradient / b row:
// 26 of 32 lanes for BS = 3 (18 of 32 in the uncondensed program).  Blocks shorter than BS keep decoupled dummy inputs
// (unit Hessian, zero rows, no bounds) at the end of ubar.  The sweeps are those of cf_rti_warp.h with three row tiles
// more per product: W = [B';A';res_b'] P (RT x 2 tiles), S = H2 + Gamma + W [B';A']' (lower tiles, accumulators
// initialised with H2), Cholesky of the 4 BS input columns, Schur complement with K = 4 BS.
//
// Needs the linearisation of the preparation kernel (CfBatchView::prep); per-stage input boxes, per-instance weights and
// per-interval time steps are honoured; the linear-residual diagnostics (lin_res_check) are not available here.
#pragma once
#include "cf_rti_warp.h"

// how the N intervals are split into N2 blocks (PART_COND_QP_COMPUTE_BLOCK_SIZE, x_part_cond.c:36-54)
struct CfPcBlocks
{
    int N2;      // condensed horizon
    int bs0;     // floor(N / N2)
    int n_big;   // the first n_big blocks hold bs0 + 1 stages
};
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    CfPcBlocks
    cf_pc_blocks(int N, int N2)
{
    CfPcBlocks b;
    b.N2 = N2; b.bs0 = N / N2; b.n_big = N - N2 * b.bs0;
    return b;
}

template <int BS>
struct CfPcWarpT
{
    enum : int {
        NUB = CF_NU * BS, NVB = NUB + CF_NX, MR = NVB + 1, NB2 = 2 * NUB,
        HSZ = ((NVB * (NVB + 1)) / 2 + 1) & ~1,       // packed lower triangle of the condensed Hessian
        // stage block: same field order as the uncondensed program (cf_rti_warp.h), the Hessian in front so that only the
        // residual+factorisation sweep stages it
        P_H = 0, R_UX = HSZ, R_PI = R_UX + MR, R_DPI = R_PI + CF_XP, R_RQ = R_DPI + CF_XP, R_D = R_RQ + MR, R_BKP = R_D + NB2,
        R_PB = R_BKP + NB2, R_DLAM = R_PB + CF_XP, R_DT = R_DLAM + NB2, R_LAM = R_DT + NB2, R_T = R_LAM + NB2, R_DUX = R_T + NB2,
        B_M = R_DUX + MR, MSZ = MR * CF_NX, B_RD = B_M + MSZ, R_RESD = B_RD, R_RESM = R_RESD + NB2, R_RESG = R_RESM + NB2,
        R_RESB = R_RESG + MR, B_LU = R_RESB + CF_XP, LUSZ = MR * NUB,   // factor columns, element (r,j) at j*MR + r
        B_PX = B_LU + LUSZ, SB = B_PX + CF_LX,
        RT = (MR + 7) / 8, CTV = (NVB + 7) / 8, T0 = NUB / 8, KSU = (NUB + 3) / 4, PKT = (CF_TRI_NX + 31) / 32,
        ALP = 20, PCO = 4 - (NUB & 1),               // P in shared memory as in the uncondensed program: element (i,j) at i*20 + 4 + j
                                                     //   (3 + j when the number of inputs is odd: the tile stores stay 16-byte aligned)
        WST = 20,                                    // row stride of W
        BUFSZ = (B_RD > HSZ + MR * WST) ? B_RD : HSZ + MR * WST,   // the W block (MR x 20) overlays the staged block behind the Hessian
        SM_BUF0 = 0, SM_BUF1 = BUFSZ, SM_P = 2 * BUFSZ, SM_V0 = SM_P + CF_NX * ALP, VST = 32, SM_V1 = SM_V0 + VST, SM_V2 = SM_V1 + VST,
        SM_BAR = SM_V2 + VST, SM_PAR = SM_BAR + 4, SM_DOUBLES = SM_PAR + ((CF_PAR_DOUBLES + 1) & ~1)
    };
    static_assert(MR <= 32 && (MR & 1) == 0, "one row per lane");
    static_assert(BUFSZ >= B_RD && BUFSZ >= SB - R_LAM && BUFSZ >= B_PX - R_BKP && MR * WST <= BUFSZ - HSZ, "staging buffers");
    static_assert((HSZ & 1) == 0 && (MSZ & 1) == 0 && (LUSZ & 1) == 0 && (SB & 1) == 0 && (ALP * CF_NX) % 2 == 0, "16-byte alignment");
    static_assert(CF_NX + PCO <= ALP && CF_NX <= 16, "two 8-column tiles hold the cost-to-go Hessian");
    static_assert(2 * BS * CF_PREP_STAGE + MSZ + HSZ <= SM_V2, "staging area of the condensing pass (buffers, P and two vectors; V2 holds the weights)");

    const CfParams *P, *PG;
    int lane, N, N2, bs0, n_big;
    double *sm;
    uint64_t *bar;
    unsigned par;
    double *SLOT, *PREP;
    const double *DT, *BST, *WTAB;
    const double *WD;    // full weight matrices per stage (block size 1 only): [(N+1)][2][17*17], see CfBatchView::W_dense
    double mu, alpha, mu_aff, sigma, pm_max;
    double nrm[4], lin[4];
    int flags;

    CF_MEM void bind(const CfParams *P_, const CfParams *PG_, double *slot, double *sm_, double *prep_, const double *dts_,
                     const CfPcBlocks &b)
    {
        P = P_; PG = PG_; N = PG_->N; sm = sm_; lane = cf_lane();
        N2 = b.N2; bs0 = b.bs0; n_big = b.n_big;
        PREP = prep_; DT = dts_; BST = nullptr; WTAB = nullptr; WD = nullptr; SLOT = slot;
        bar = reinterpret_cast<uint64_t *>(sm_ + SM_BAR);
        par = 0;
        lin[0] = lin[1] = lin[2] = lin[3] = 0.0;
    }
    CF_MEM double dt(int k) const { return DT ? DT[k] : PG->Ts; }
    CF_MEM double wgt(int k, int idx) const { return WTAB ? WTAB[k * CF_NY + idx] : (k < N ? P->Wdiag[idx] : P->WNdiag[idx]); }
    CF_MEM int bsz(int i) const { return i < N2 ? (i < n_big ? bs0 + 1 : bs0) : 0; }             // stages in block i
    CF_MEM int kfirst(int i) const { return i < n_big ? i * (bs0 + 1) : n_big * (bs0 + 1) + (i - n_big) * bs0; }
    CF_MEM double step_adjust(double a) const { return (a < 1.0) ? a * ((1.0 - a) * 0.99 + a * 0.9999999) : a; }

    // ---- TMA staging (as in CfWarpT)
    CF_MEM void pass_begin()
    {
        cf_syncwarp();
        if (lane == 0) cf_fence_proxy_async();
    }
    CF_MEM double *blk(int k) const { return SLOT + (long) k * SB; }
    CF_MEM double *rec(int k) const { return blk(k); }
    CF_MEM double *buf(int bf) const { return sm + (bf ? SM_BUF1 : SM_BUF0); }
    CF_MEM void fetch(int bf, int k, int start, int len)
    {
        if (lane == 0) {
            cf_bulk_expect(bar + bf, len * 8);
            cf_bulk_g2s_raw(buf(bf), blk(k) + start, len * 8, bar + bf);
        }
    }
    CF_MEM void wait(int bf)
    {
        cf_bulk_wait(bar + bf, (par >> bf) & 1u);
        par ^= 1u << bf;
    }

    // =============================================================== condensing
    // Builds the N2 + 1 condensed stage blocks in the scratch slot from the instance's prepared linearisation
    // (CF_PREP_STAGE records [ [B';A';b'] 18 x 13 | gradient 18 ]), evaluates the bound vectors and the initial
    // interior-point variables (OCP_QP_INIT_VAR scheme 1) of every real input, and eliminates x0 from stage 0
    // (x_ocp_qp_red.c:310-330).  Lane r < MR owns row r of G (registers) and of the condensed Hessian (shared memory).
    CF_MEM void condense(const double *xg, const double *ug, const double *x0g, const double *yrefg, const double *yref_eg)
    {
        double *GS = sm + 2 * BS * CF_PREP_STAGE;           // G (MR x 13, element (r,c) at c*MR + r) for the products
        double *HS = GS + MSZ;                              // packed lower triangle of H2
        const bool vl = lane < NVB, ul = lane < NUB, xl = lane >= NUB && vl;
        const int ci = xl ? lane - NUB : 0, lv = vl ? lane : 0, trl = cf_tri(lv);
        // (sqrt(w))^2 of the 13 state weights (ocp_nlp_cost_ls.c:739-772), once per instance unless they differ per stage
        double *QW = sm + SM_V2;   // (the staging area of this pass extends over P, V0 and V1)
        pass_begin();
        if (lane < CF_NX) { const double r = sqrt(P->Wdiag[lane]); QW[lane] = r * r; }
        // the prepared records of block i+1 travel while block i is condensed (two staging areas, two mbarriers)
        if (lane == 0) {
            cf_bulk_expect(bar, bsz(0) * CF_PREP_STAGE * 8);
            cf_bulk_g2s_raw(sm, PREP, bsz(0) * CF_PREP_STAGE * 8, bar);
        }
        CF_NOUNROLL
        for (int i = 0; i < N2; i++) {
            const int bs = bsz(i), k0 = kfirst(i), bf = i & 1;
            double *ST = sm + bf * (BS * CF_PREP_STAGE);
            double *bk = blk(i);
            if (i + 1 < N2 && lane == 0) {   // (the barrier that ended block i-1 ordered its reads of that staging area)
                cf_fence_proxy_async();
                cf_bulk_expect(bar + (bf ^ 1), bsz(i + 1) * CF_PREP_STAGE * 8);
                cf_bulk_g2s_raw(sm + (bf ^ 1) * (BS * CF_PREP_STAGE), PREP + (long) kfirst(i + 1) * CF_PREP_STAGE,
                                bsz(i + 1) * CF_PREP_STAGE * 8, bar + (bf ^ 1));
            }
            for (int e = lane; e < HSZ; e += 32) HS[e] = 0.0;
            double grow[CF_NX];
            double rq = 0.0, hdiag = 1.0, v0 = 0.0;        // dummy inputs: unit Hessian, nothing else
            wait(bf);
            if (i == 0) eliminate_x0(ST, xg, x0g);
            cf_syncwarp();
            if (xl) {
                double w2 = QW[ci];
                if (WTAB) { const double r = sqrt(WTAB[k0 * CF_NY + ci]); w2 = r * r; }
                hdiag = dt(k0) * w2;
                rq = ST[CF_MSZ + CF_NU + ci];               // zero for stage 0 (its x is eliminated)
            }
            CF_NOUNROLL
            for (int j = 0; j < bs; j++) {
                const double *Mj = ST + j * CF_PREP_STAGE, *gj = Mj + CF_MSZ;
                const int k = k0 + j, off = CF_NU * (bs - 1 - j);
                const double h = dt(k);
                const int e = lane - off;
                const bool mine = e >= 0 && e < CF_NU;     // this lane's input belongs to stage k
                if (mine) {
                    const double r = sqrt(wgt(k, CF_NX + e));
                    hdiag = h * (r * r);
                    rq = gj[e];
                    // bounds (ocp_nlp_constraints_bgh.c:1634-1636) and OCP_QP_INIT_VAR scheme 1 (x_ocp_qp_ipm.c:1636-1769)
                    const double uk = ug[k * CF_NU + e];
                    double lb = (k == 0) ? P->lbu0[e] : P->lbu[e], ub = (k == 0) ? P->ubu0[e] : P->ubu[e];
                    if (BST) { lb = BST[k * 2 * CF_NU + e]; ub = BST[k * 2 * CF_NU + CF_NU + e]; }
                    const double dl = lb - uk, du = uk - ub;
                    double tl = -dl, tu = -du;
                    if (tl < CF_THR0) {
                        if (tu < CF_THR0) { v0 = 0.5 * (dl - du); tl = CF_THR0; tu = CF_THR0; }
                        else { tl = CF_THR0; v0 = dl + CF_THR0; }
                    } else if (tu < CF_THR0) { tu = CF_THR0; v0 = -du - CF_THR0; }
                    bk[R_D + lane] = dl; bk[R_D + NUB + lane] = du;
                    bk[R_T + lane] = tl; bk[R_T + NUB + lane] = tu;
                    bk[R_LAM + lane] = CF_MU0 / tl; bk[R_LAM + NUB + lane] = CF_MU0 / tu;
                }
                const double *Mrow = Mj + (mine ? e : (lane == NVB ? CF_NV : (xl ? CF_NU + ci : 0)));   // this lane's row of the record
                if (j == 0) {
                    // G_1 = G_0 A' + [E B'; 0; b'] with G_0 = [0; I; 0]: the rows of the record itself, no product
                    const bool any = mine || xl || lane == NVB;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) grow[c] = any ? Mrow[c * CF_MROWS] : 0.0;
                    continue;
                }
                // cost of the inner state x_k = G' z:  H2 += G Q_k G',  rq2 += G (q_k + Q_k c)
                if (lane < MR) {
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) GS[c * MR + lane] = grow[c];
                }
                cf_syncwarp();
                {
                    double t[CF_NX], g0 = 0.0, g1 = 0.0;
                    CF_UNROLL
                    for (int c = 0; c < CF_NX; c++) {
                        double w2 = QW[c];
                        if (WTAB) { const double r = sqrt(WTAB[k * CF_NY + c]); w2 = r * r; }
                        const double q = h * w2;
                        t[c] = grow[c] * q;
                        const double v = grow[c] * (gj[CF_NU + c] + q * GS[c * MR + NVB]);
                        if (c & 1) g1 += v;
                        else g0 += v;
                    }
                    rq += g0 + g1;
                    CF_NOUNROLL
                    for (int c2 = 0; c2 < NVB - 1; c2 += 2) {   // two columns per trip: four independent chains
                        double s0 = 0.0, s1 = 0.0, r0 = 0.0, r1 = 0.0;
                        CF_UNROLL
                        for (int c = 0; c + 1 < CF_NX; c += 2) {
                            const cf_d2 ga = cf_ld2(GS + c * MR + c2), gb = cf_ld2(GS + (c + 1) * MR + c2);
                            s0 += t[c] * ga.x; r0 += t[c] * ga.y;
                            s1 += t[c + 1] * gb.x; r1 += t[c + 1] * gb.y;
                        }
                        const cf_d2 gl = cf_ld2(GS + (CF_NX - 1) * MR + c2);
                        s0 += t[CF_NX - 1] * gl.x; r0 += t[CF_NX - 1] * gl.y;
                        if (vl && c2 <= lane) HS[trl + c2] += s0 + s1;
                        if (vl && c2 + 1 <= lane) HS[trl + c2 + 1] += r0 + r1;
                    }
                    {   // NVB is odd: the last column
                        double s0 = 0.0, s1 = 0.0;
                        CF_UNROLL
                        for (int c = 0; c + 1 < CF_NX; c += 2) {
                            s0 += t[c] * GS[c * MR + NVB - 1];
                            s1 += t[c + 1] * GS[(c + 1) * MR + NVB - 1];
                        }
                        s0 += t[CF_NX - 1] * GS[(CF_NX - 1) * MR + NVB - 1];
                        if (lane == NVB - 1) HS[trl + NVB - 1] += s0 + s1;
                    }
                }
                cf_syncwarp();
                // G <- G A_k' + [E B_k' ; 0 ; b_k']   (A_k'[i][c] = M_k[4+i][c], element (r,c) of a record at c*18 + r)
                double gn[CF_NX];
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) {
                    const double *Ac = Mj + c * CF_MROWS + CF_NU;
                    double s0 = 0.0, s1 = 0.0;
                    if constexpr ((CF_NU & 1) == 0) {   // 16-byte aligned column start
                        CF_UNROLL
                        for (int ip = 0; ip < CF_NX / 2; ip++) {
                            const cf_d2 a2 = cf_ld2(Ac + 2 * ip);
                            s0 += grow[2 * ip] * a2.x;
                            s1 += grow[2 * ip + 1] * a2.y;
                        }
                        if (CF_NX & 1) s0 += grow[CF_NX - 1] * Ac[CF_NX - 1];
                    } else {
                        CF_UNROLL
                        for (int ip = 0; ip < CF_NX; ip++) {
                            if (ip & 1) s1 += grow[ip] * Ac[ip];
                            else s0 += grow[ip] * Ac[ip];
                        }
                    }
                    const double add = (mine || lane == NVB) ? Mrow[c * CF_MROWS] : 0.0;
                    gn[c] = (s0 + s1) + add;
                }
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) grow[c] = gn[c];
            }
            // ---- the condensed stage block
            if (vl) HS[trl + lane] += hdiag;
            if constexpr (BS == 1) {
                if (WD) {
                    // full weight matrix of stage k0 (ocp_nlp_cost_ls.c:743-772,883-912): Hessian dt (Cyt W_chol)(Cyt W_chol)'
                    // (the product comes from the host, [u;x] order), gradient dt Cyt W (y - yref).  Stage 0: its states are
                    // eliminated -- decoupled dummies here -- and leave S xbar in the input gradient (x_ocp_qp_red.c:354).
                    const double *Hd = WD + (long) k0 * (2 * CF_NV * CF_NV), *Wp = Hd + CF_NV * CF_NV;
                    const double h = dt(k0);
                    double *RS = sm + SM_V0, *XB = sm + SM_V1;
                    cf_syncwarp();
                    if (vl) {
                        RS[lane] = ul ? ug[k0 * CF_NU + lane] - yrefg[k0 * CF_NY + CF_NX + lane]
                                      : xg[k0 * CF_NX + ci] - yrefg[k0 * CF_NY + ci];
                        XB[lane] = (i == 0 && xl) ? x0g[ci] - xg[ci] : 0.0;
                    }
                    cf_syncwarp();
                    if (vl) {
                        double g = 0.0, sx = 0.0;
                        CF_NOUNROLL
                        for (int c = 0; c < NVB; c++) {
                            const double hv = h * Hd[lane * CF_NV + c];
                            g += Wp[lane * CF_NV + c] * RS[c];
                            sx += hv * XB[c];
                            const bool dummy = i == 0 && (c >= NUB || xl);       // eliminated states: unit diagonal, no coupling
                            if (c <= lane) HS[trl + c] = dummy ? (c == lane ? 1.0 : 0.0) : hv;
                        }
                        rq = (i == 0 && xl) ? 0.0 : h * g + ((i == 0) ? sx : 0.0);
                    }
                }
            }
            cf_syncwarp();
            for (int e = lane; e < HSZ; e += 32) bk[P_H + e] = HS[e];
            if (lane < MR) {
                CF_UNROLL
                for (int c = 0; c < CF_NX; c++) bk[B_M + c * MR + lane] = grow[c];
                bk[R_RQ + lane] = vl ? rq : 0.0;
                bk[R_UX + lane] = vl ? v0 : 0.0;
                bk[R_DUX + lane] = 0.0;
            }
            if (lane < CF_NX) { bk[R_PI + lane] = 0.0; bk[R_DPI + lane] = 0.0; }
            if (lane < NB2) { bk[R_DLAM + lane] = 0.0; bk[R_DT + lane] = 0.0; }
            if (ul && lane >= CF_NU * bs) {     // dummy inputs: finite, never counted
                bk[R_D + lane] = -1.0; bk[R_D + NUB + lane] = -1.0;
                bk[R_T + lane] = 1.0; bk[R_T + NUB + lane] = 1.0;
                bk[R_LAM + lane] = 1.0; bk[R_LAM + NUB + lane] = 1.0;
            }
            cf_syncwarp();   // this block's staging area, GS and HS may be rewritten
        }
        // terminal stage: diagonal Hessian (dummy inputs 1, states W_e), gradient from the preparation, no dynamics
        {
            double *bk = blk(N2);
            for (int e = lane; e < HSZ; e += 32) bk[P_H + e] = 0.0;
            cf_syncwarp();
            if (vl) {
                double hN = 1.0;
                if (xl) { const double r = sqrt(wgt(N, ci)); hN = r * r; }
                bk[P_H + trl + lane] = hN;
            }
            double gN = xl ? PREP[(long) N * CF_PREP_STAGE + CF_NU + ci] : 0.0;
            if constexpr (BS == 1) {
                if (WD) {   // full terminal weight (13 x 13 state block of table row N; terminal scaling 1)
                    const double *Hd = WD + (long) N * (2 * CF_NV * CF_NV), *Wp = Hd + CF_NV * CF_NV;
                    double *RS = sm + SM_V0;
                    cf_syncwarp();
                    if (xl) RS[lane] = xg[N * CF_NX + ci] - yref_eg[ci];
                    cf_syncwarp();
                    if (xl) {
                        double g = 0.0;
                        CF_NOUNROLL
                        for (int c = NUB; c < NVB; c++) {
                            g += Wp[lane * CF_NV + c] * RS[c];
                            if (c <= lane) bk[P_H + trl + c] = Hd[lane * CF_NV + c];
                        }
                        gN = g;
                    }
                }
            }
            if (lane < MR) {
                bk[R_RQ + lane] = gN;
                bk[R_UX + lane] = 0.0; bk[R_DUX + lane] = 0.0;
            }
            if (lane < CF_NX) { bk[R_PI + lane] = 0.0; bk[R_DPI + lane] = 0.0; }
            if (lane < NB2) {
                bk[R_DLAM + lane] = 0.0; bk[R_DT + lane] = 0.0; bk[R_D + lane] = -1.0; bk[R_T + lane] = 1.0; bk[R_LAM + lane] = 1.0;
            }
        }
        cf_syncwarp();
    }

    // x0 elimination on the staged record of stage 0 (18-row layout of the prepared store): b_0 += A_0 (x0 - x_0), A rows dropped
    CF_MEM void eliminate_x0(double *MS, const double *xg, const double *x0g)
    {
        const bool xs = lane >= CF_NU && lane < CF_NV;
        double *Mrow = MS + (lane < CF_NV ? lane : CF_NV);
        const double xbar = xs ? (x0g[lane - CF_NU] - xg[lane - CF_NU]) : 0.0;
        CF_NOUNROLL
        for (int i = 0; i < CF_NX; i++) {
            const double tot = cf_warp_sum(xs ? Mrow[i * CF_MROWS] * xbar : 0.0);
            if (lane == CF_NV) MS[i * CF_MROWS + CF_NV] = tot + MS[i * CF_MROWS + CF_NV];
            else if (xs) Mrow[i * CF_MROWS] = 0.0;
        }
        cf_syncwarp();
    }

    // =============================================================== IPM sweeps (cf_rti_warp.h with fatter stages)
    // residual_factorize: UPDATE_VAR_QP + OCP_QP_RES_COMPUTE (+ norms) + the backward Riccati factorisation
    // (x_core_qp_ipm_aux.c:220-325, x_ocp_qp_res.c:334-470,602-637, x_ocp_qp_kkt.c:573-740) in one backward sweep.
    CF_MEM void residual_factorize(const double a_raw, const bool do_factor)
    {
        const double a = step_adjust(a_raw);
        double ng = 0, nb = 0, nd = 0, nm = 0, mus = 0;
        double *PS = sm + SM_P, *PV = sm + SM_V0, *UXS = sm + SM_V1, *PIS = sm + SM_V2, *G = sm + SM_V1, *HD = sm + SM_V2;
        pass_begin();
        fetch(N2 & 1, N2, 0, B_RD);
        const bool vl = lane < NVB, xl = lane >= NUB && vl;
        const int ci = xl ? lane - NUB : 0, lv = vl ? lane : 0, trl = cf_tri(lv);
        const int fg = lane >> 2, fq = lane & 3;
        const int rl = lane < MR ? lane : MR - 1;
        int pa[4][2];   // pa[kk][h] = address of P[4kk+fq][8h+fg] in the lower-triangular shared-memory array
        CF_UNROLL
        for (int kk = 0; kk < 4; kk++)
            CF_UNROLL
            for (int hh = 0; hh < 2; hh++) {
                const int i = 4 * kk + fq, j = 8 * hh + fg;
                const bool ok = i < CF_NX && j < CF_NX;
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                pa[kk][hh] = ok ? hi * ALP + lo + PCO : -1;
            }
        int pk[PKT];    // packed lower triangle of P_{k+1}: element e = lane + 32 t
        CF_UNROLL
        for (int t = 0; t < PKT; t++) {
            const int e = lane + 32 * t;
            int i = 0;
            CF_UNROLL
            for (int q = 1; q < CF_NX; q++) i += (e >= cf_tri(q)) ? 1 : 0;
            pk[t] = (e < CF_TRI_NX) ? i * ALP + (e - cf_tri(i)) + PCO : -1;
        }
        int trr[RT];    // packed row starts of the fragment rows
        CF_UNROLL
        for (int t = 0; t < RT; t++) { const int r = 8 * t + fg; trr[t] = r < NVB ? cf_tri(r) : 0; }
        double ux_next = 0.0, pi_k = 0.0;
        CF_NOUNROLL
        for (int k = N2; k >= 0; k--) {
            const int bf = k & 1;
            const bool kl = k < N2;
            const int nbk = CF_NU * bsz(k);
            const bool bl = lane < nbk;
            const int lb = bl ? lane : 0;
            wait(bf);
            cf_syncwarp();
            if (k > 0) fetch(bf ^ 1, k - 1, 0, B_RD);
            if (do_factor && kl) {
                double *LFk = blk(k) + B_PX;
                CF_UNROLL
                for (int t = 0; t < PKT; t++)
                    if (pk[t] >= 0) LFk[lane + 32 * t] = PS[pk[t]];
            }
            double *VS = buf(bf);
            double *rk = rec(k);
            const double *HP = VS + P_H;
            // ---------------- update + residuals
            const double uxc = vl ? VS[R_UX + lv] + a * VS[R_DUX + lv] : 0.0;
            if (vl) rk[R_UX + lane] = uxc;
            const double pim = (k > 0 && xl) ? VS[R_PI + ci] + a * VS[R_DPI + ci] : 0.0;
            if (k > 0 && xl) rk[R_PI + ci] = pim;
            if (lane < MR) UXS[lane] = uxc;
            if (xl) PIS[ci] = pi_k;
            cf_syncwarp();
            double rg;
            {   // res_g = H2 ux + rq - pi_{k-1}: symmetric product from the packed lower triangle (SYMV_L)
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int c = 0; c < NVB; c++) {
                    const double hv = HP[(c <= lv) ? trl + c : cf_tri(c) + lv];
                    if (c & 1) s1 += hv * UXS[c];
                    else s0 += hv * UXS[c];
                }
                rg = (s0 + s1) + VS[R_RQ + lv] - pim;
            }
            double Gam = 0.0, gam = 0.0;
            {
                double ll = VS[R_LAM + lb] + a * VS[R_DLAM + lb], lu = VS[R_LAM + NUB + lb] + a * VS[R_DLAM + NUB + lb];
                double tl = VS[R_T + lb] + a * VS[R_DT + lb], tu = VS[R_T + NUB + lb] + a * VS[R_DT + NUB + lb];
                ll = ll <= CF_LAM_MIN ? CF_LAM_MIN : ll; lu = lu <= CF_LAM_MIN ? CF_LAM_MIN : lu;
                tl = tl <= CF_T_MIN ? CF_T_MIN : tl; tu = tu <= CF_T_MIN ? CF_T_MIN : tu;
                const double rdl = VS[R_D + lb] + tl - uxc, rdu = VS[R_D + NUB + lb] + tu + uxc;
                const double rml = ll * tl, rmu = lu * tu;
                if (bl) {
                    rk[R_LAM + lane] = ll; rk[R_LAM + NUB + lane] = lu;
                    rk[R_T + lane] = tl; rk[R_T + NUB + lane] = tu;
                    rk[R_RESD + lane] = rdl; rk[R_RESD + NUB + lane] = rdu;
                    rk[R_BKP + lane] = rml; rk[R_BKP + NUB + lane] = rmu;
                    rk[R_RESM + lane] = rml - CF_TAU_MIN; rk[R_RESM + NUB + lane] = rmu - CF_TAU_MIN;
                    rg += lu - ll;
                    mus += rml + rmu;
                    cf_amax(nd, rdl); cf_amax(nd, rdu);
                    cf_amax(nm, rml); cf_amax(nm, rmu);
                    const double til = cf_rcp(tl), tiu = cf_rcp(tu);
                    Gam = til * ll + tiu * lu;
                    gam = til * ((rml - CF_TAU_MIN) - ll * rdl) - tiu * ((rmu - CF_TAU_MIN) - lu * rdu);
                }
            }
            if (kl) {
                double *Mk = VS + B_M;
                {   // res_g += [B';A'] pi_k
                    double s0 = 0.0, s1 = 0.0;
                    CF_UNROLL
                    for (int cp = 0; cp < CF_NX / 2; cp++) {
                        const cf_d2 p2 = cf_ld2(PIS + 2 * cp);
                        s0 += Mk[(2 * cp) * MR + lv] * p2.x;
                        s1 += Mk[(2 * cp + 1) * MR + lv] * p2.y;
                    }
                    if (CF_NX & 1) s0 += Mk[(CF_NX - 1) * MR + lv] * PIS[CF_NX - 1];
                    rg += s0 + s1;
                }
                {   // res_b = (b - x+) + [A B] ux   (column ci: rows 0..NVB-1, row NVB = b)
                    const double *Mc = Mk + ci * MR;
                    double s0 = 0.0, s1 = 0.0;
                    CF_UNROLL
                    for (int rp = 0; rp < (NVB - 1) / 2; rp++) {
                        const cf_d2 m2 = cf_ld2(Mc + 2 * rp), u2 = cf_ld2(UXS + 2 * rp);
                        s0 += m2.x * u2.x;
                        s1 += m2.y * u2.y;
                    }
                    const cf_d2 m2 = cf_ld2(Mc + NVB - 1);
                    s0 += m2.x * UXS[NVB - 1];
                    const double rb = (m2.y - ux_next) + (s0 + s1);
                    cf_syncwarp();
                    if (xl) {
                        cf_amax(nb, rb);
                        rk[R_RESB + ci] = rb;
                        Mk[ci * MR + NVB] = rb;
                    }
                }
            }
            if (vl) { rk[R_RESG + lane] = rg; cf_amax(ng, rg); }
            ux_next = uxc;
            pi_k = pim;
            // ---------------- factorisation
            if (!do_factor) continue;
            if (!kl) {
                for (int i = lane; i < CF_NX * ALP; i += 32) PS[i] = 0.0;
                cf_syncwarp();
                if (xl) {
                    // P_N = state block of the terminal Hessian (+ reg); lower triangle (dense with full weight matrices)
                    CF_NOUNROLL
                    for (int c = 0; c <= ci; c++) PS[ci * ALP + c + PCO] = HP[trl + NUB + c] + (c == ci ? CF_REG_PRIM : 0.0);
                    PV[ci] = rg;
                    rk[R_DUX + lane] = rg;
                }
                continue;
            }
            const double g = vl ? rg + gam : 0.0, hd = CF_REG_PRIM + Gam;
            cf_syncwarp();
            const double *Mk = VS + B_M;
            double *WS = VS + HSZ;
            double am[RT][4], wt[RT][2][2];
            CF_UNROLL
            for (int t = 0; t < RT; t++) { wt[t][0][0] = wt[t][0][1] = wt[t][1][0] = wt[t][1][1] = 0.0; }
            CF_UNROLL
            for (int kk = 0; kk < 4; kk++) {
                const int kc = 4 * kk + fq;
                const bool kv = kc < CF_NX;
                const double b0 = pa[kk][0] >= 0 ? PS[pa[kk][0]] : 0.0;
                const double b1 = pa[kk][1] >= 0 ? PS[pa[kk][1]] : 0.0;
                CF_UNROLL
                for (int t = 0; t < RT; t++) {
                    const int r = 8 * t + fg;
                    am[t][kk] = (kv && r < MR) ? Mk[kc * MR + r] : 0.0;
                    cf_dmma(wt[t][0][0], wt[t][0][1], am[t][kk], b0);
                    cf_dmma(wt[t][1][0], wt[t][1][1], am[t][kk], b1);
                }
            }
            // row NVB: Pb = P res_b, then + p_{k+1}'
            if (fg == (NVB & 7)) {
                double *pb = rk + R_PB;
                CF_UNROLL
                for (int tp = 0; tp < 2; tp++)
                    CF_UNROLL
                    for (int e = 0; e < 2; e++) {
                        const int c = 8 * tp + 2 * fq + e;
                        if (c < CF_NX) { pb[c] = wt[NVB >> 3][tp][e]; wt[NVB >> 3][tp][e] += PV[c]; }
                    }
            }
            cf_syncwarp();
            CF_UNROLL
            for (int t = 0; t < RT; t++) {
                const int r = 8 * t + fg;
                if (r < MR) {
                    cf_st2(WS + r * WST + 2 * fq, wt[t][0][0], wt[t][0][1]);
                    cf_st2(WS + r * WST + 8 + 2 * fq, wt[t][1][0], wt[t][1][1]);
                }
            }
            if (lane < MR) { G[lane] = g; HD[lane] = hd; }
            cf_syncwarp();
            // ---- S = H2 + Gamma + W M'
            double wf[RT][4];
            CF_UNROLL
            for (int t = 0; t < RT; t++) {
                const int r = 8 * t + fg;
                CF_UNROLL
                for (int kk = 0; kk < 4; kk++) wf[t][kk] = (r < MR) ? WS[r * WST + 4 * kk + fq] : 0.0;
            }
            double sx[RT][CTV][2];
            CF_UNROLL
            for (int t = 0; t < RT; t++) {
                CF_UNROLL
                for (int tp = 0; tp < CTV; tp++) {
                    if (tp > t) continue;
                    const int r = 8 * t + fg, c0 = 8 * tp + 2 * fq;
                    double s0 = 0.0, s1 = 0.0;
                    if (r < NVB) {
                        if (c0 <= r) s0 = HP[trr[t] + c0];
                        if (c0 + 1 <= r) s1 = HP[trr[t] + c0 + 1];
                    } else if (r == NVB) {
                        if (c0 < NVB) s0 = G[c0];
                        if (c0 + 1 < NVB) s1 = G[c0 + 1];
                    }
                    if (r == c0) s0 += HD[r < MR ? r : 0];
                    if (r == c0 + 1) s1 += HD[r < MR ? r : 0];
                    CF_UNROLL
                    for (int kk = 0; kk < 4; kk++) cf_dmma(s0, s1, wf[t][kk], am[tp][kk]);
                    sx[t][tp][0] = s0; sx[t][tp][1] = s1;
                }
            }
            cf_syncwarp();
            double *LUs = WS;
            CF_UNROLL
            for (int t = 0; t < RT; t++) {
                CF_UNROLL
                for (int tp = 0; tp < CTV; tp++) {
                    if (tp > t || 8 * tp >= NUB) continue;
                    const int r = 8 * t + fg, c0 = 8 * tp + 2 * fq;
                    if (r < MR && c0 < NUB) LUs[c0 * MR + r] = sx[t][tp][0];
                    if (r < MR && c0 + 1 < NUB) LUs[(c0 + 1) * MR + r] = sx[t][tp][1];
                }
            }
            cf_syncwarp();
            // ---- POTRF_L_MN(nv+1, nu): the NUB input columns, lane = row; non-positive pivot -> 0
            {
                double o[NUB], og[NUB];
                CF_UNROLL
                for (int j = 0; j < NUB; j++) o[j] = (lane < MR) ? LUs[j * MR + rl] : 0.0;
                CF_UNROLL
                for (int j = 0; j < NUB; j++) {
                    double v = o[j];
                    CF_UNROLL
                    for (int c = 0; c < j; c++) v -= o[c] * LUs[c * MR + j];
                    const double piv = cf_shfl(v, j);
                    double dj, inv;
                    cf_sqrt_rsqrt(piv, dj, inv);
                    if (!(piv > 0.0)) { dj = 0.0; inv = 0.0; flags |= CF_FLAG_BAD_PIVOT; }
                    o[j] = (rl == j) ? dj : ((rl > j) ? v * inv : 0.0);
                    if (lane < MR) LUs[j * MR + rl] = o[j];
                    og[j] = (rl == j) ? inv : o[j];
                    cf_syncwarp();
                }
                double *LFk = blk(k) + B_LU;
                if (lane < MR) {
                    CF_UNROLL
                    for (int j = 0; j < NUB; j++) LFk[j * MR + lane] = og[j];
                }
                if (lane == NVB) {
                    CF_UNROLL
                    for (int jp = 0; jp < NUB / 2; jp++) cf_st2(rk + R_DUX + 2 * jp, o[2 * jp], o[2 * jp + 1]);
                    if (NUB & 1) rk[R_DUX + NUB - 1] = o[NUB - 1];
                }
            }
            // ---- Schur complement: S_xx -= Ls Ls' (K = NUB)
            double la[RT][KSU];
            CF_UNROLL
            for (int t = 0; t < RT; t++) {
                const int r = 8 * t + fg;
                CF_UNROLL
                for (int ks = 0; ks < KSU; ks++) la[t][ks] = (t >= T0 && r < MR && 4 * ks + fq < NUB) ? LUs[(4 * ks + fq < NUB ? 4 * ks + fq : 0) * MR + r] : 0.0;
            }
            cf_syncwarp();
            CF_UNROLL
            for (int t = T0; t < RT; t++) {
                CF_UNROLL
                for (int tp = T0; tp < CTV; tp++) {
                    if (tp > t) continue;
                    CF_UNROLL
                    for (int ks = 0; ks < KSU; ks++) cf_dmma(sx[t][tp][0], sx[t][tp][1], -la[t][ks], la[tp][ks]);
                    const int r = 8 * t + fg, c0 = 8 * tp + 2 * fq;
                    const bool xrow = r >= NUB && r < NVB;
                    if (xrow && c0 + PCO >= NUB && c0 < NUB + 16) cf_st2(PS + (r - NUB) * ALP + (c0 - NUB + PCO), sx[t][tp][0], sx[t][tp][1]);
                    if (r == NVB) {
                        CF_UNROLL
                        for (int e = 0; e < 2; e++) {
                            const int c = c0 + e;
                            if (c >= NUB && c < NVB) { PV[c - NUB] = sx[t][tp][e]; rk[R_DUX + c] = sx[t][tp][e]; }
                        }
                    }
                }
            }
        }
        nrm[0] = cf_warp_max(ng); nrm[1] = cf_warp_max(nb); nrm[2] = cf_warp_max(nd); nrm[3] = cf_warp_max(nm);
        mu = cf_warp_sum(mus) * (1.0 / (double) (2 * CF_NU * N));
        cf_syncwarp();
    }

    // Gamma / gamma of a rhs-only solve for the bound of input `lane` (lanes < nbk); see CfWarpT::bound_terms
    CF_MEM void bound_terms(int k, const double *f, int rm_mode, double sigma_mu, double &Gam, double &gam)
    {
        double ll = f[R_LAM + lane], lu = f[R_LAM + NUB + lane];
        const double til = cf_rcp(f[R_T + lane]), tiu = cf_rcp(f[R_T + NUB + lane]);
        double rml = f[R_BKP + lane], rmu = f[R_BKP + NUB + lane];
        if (rm_mode == 1) {
            rml = rml + f[R_DT + lane] * f[R_DLAM + lane] - sigma_mu;
            rmu = rmu + f[R_DT + NUB + lane] * f[R_DLAM + NUB + lane] - sigma_mu;
        } else { rml -= sigma_mu; rmu -= sigma_mu; }
        rec(k)[R_RESM + lane] = rml; rec(k)[R_RESM + NUB + lane] = rmu;
        const double gl = til * (rml - ll * f[R_RESD + lane]);
        const double gu = tiu * (rmu - lu * f[R_RESD + NUB + lane]);
        Gam = til * ll + tiu * lu;
        gam = gl - gu;
    }

    // forward substitution + dlam, dt, step length (x_ocp_qp_kkt.c:536-570,741-758, x_core_qp_ipm_aux.c:117-216)
    CF_MEM void forward(const bool need_pi)
    {
        double *XS = sm + SM_V0, *DS = sm + SM_V1;
        double dn = 1.0, dd = -1.0, pn_ = 1.0, pd_ = -1.0;
        double dxk = 0.0;
        const int VO = R_LAM, VN = (need_pi ? SB : B_PX) - R_LAM;
        pass_begin();
        if (N2 > 0) fetch(0, 0, VO, VN);
        XS[lane] = 0.0;
        const bool vl = lane < NVB, ul = lane < NUB, xl = lane >= NUB && vl;
        const int ci = xl ? lane - NUB : 0, lv = vl ? lane : 0, ju = ul ? lane : 0;
        double *PE = sm + SM_P;   // P_{k+1} expanded to full symmetric rows, stride CF_PST
        int pe_a[PKT], pe_b[PKT];
        CF_UNROLL
        for (int t = 0; t < PKT; t++) {
            const int e = lane + 32 * t;
            int i = 0;
            CF_UNROLL
            for (int q = 1; q < CF_NX; q++) i += (e >= cf_tri(q)) ? 1 : 0;
            const int j = e - cf_tri(i);
            pe_a[t] = (e < CF_TRI_NX) ? i * CF_PST + j : -1;
            pe_b[t] = j * CF_PST + i;
        }
        CF_NOUNROLL
        for (int k = 0; k < N2; k++) {
            const int bf = k & 1;
            double *rk = rec(k);
            const bool bl = lane < CF_NU * bsz(k);
            const int lb = bl ? lane : 0;
            const double pnext = need_pi ? rec(k + 1)[R_DUX + lv] : 0.0;
            wait(bf);
            cf_syncwarp();
            if (k + 1 < N2) fetch(bf ^ 1, k + 1, VO, VN);
            const double *VS = buf(bf) - VO;
            const double *Mk = VS + B_M, *LU = VS + B_LU;
            if (need_pi) {
                CF_UNROLL
                for (int t = 0; t < PKT; t++) {
                    if (pe_a[t] >= 0) {
                        const double v = VS[B_PX + lane + 32 * t];
                        PE[pe_a[t]] = v;
                        PE[pe_b[t]] = v;
                    }
                }
            }
            // ---- du = Luu^-T ( -l_u - Lxu' dx ): lane j < NUB owns column j
            double v;
            {
                double v0 = -VS[R_DUX + ju], v1 = 0.0;
                CF_UNROLL
                for (int ip = 0; ip < CF_NX / 2; ip++) {
                    const cf_d2 x2 = cf_ld2(XS + 2 * ip);
                    v0 -= LU[ju * MR + NUB + 2 * ip] * x2.x;
                    v1 -= LU[ju * MR + NUB + 2 * ip + 1] * x2.y;
                }
                if (CF_NX & 1) v0 -= LU[ju * MR + NUB + CF_NX - 1] * XS[CF_NX - 1];
                v = v0 + v1;
            }
            const double invd = LU[ju * MR + ju];
            double du = 0.0;
            CF_UNROLL
            for (int j = NUB - 1; j >= 0; j--) {
                const double duj = cf_shfl(v * invd, j);
                const double vn = v - LU[ju * MR + j] * duj;
                du = (ju == j) ? duj : du;
                v = (ju < j) ? vn : v;
            }
            const double duxk = ul ? du : dxk;
            if (vl && need_pi) rk[R_DUX + lane] = duxk;
            {
                const double ll = VS[R_LAM + lb], lu = VS[R_LAM + NUB + lb];
                const double tl = VS[R_T + lb], tu = VS[R_T + NUB + lb];
                const double rdl = VS[R_RESD + lb], rdu = VS[R_RESD + NUB + lb];
                const double rml = VS[R_RESM + lb], rmu = VS[R_RESM + NUB + lb];
                const double til = cf_rcp(tl), tiu = cf_rcp(tu);
                double dtl = du, dtu = -du;
                const double dlam_l = -til * (rml + (ll * dtl) - (ll * rdl));
                const double dlam_u = -tiu * (rmu + (lu * dtu) - (lu * rdu));
                dtl -= rdl; dtu -= rdu;
                if (bl) {
                    rk[R_DLAM + lane] = dlam_l; rk[R_DLAM + NUB + lane] = dlam_u;
                    rk[R_DT + lane] = dtl; rk[R_DT + NUB + lane] = dtu;
                    bool c;
                    c = dn * dlam_l < ll * dd; dn = c ? ll : dn; dd = c ? dlam_l : dd;
                    c = pn_ * dtl < tl * pd_; pn_ = c ? tl : pn_; pd_ = c ? dtl : pd_;
                    c = dn * dlam_u < lu * dd; dn = c ? lu : dn; dd = c ? dlam_u : dd;
                    c = pn_ * dtu < tu * pd_; pn_ = c ? tu : pn_; pd_ = c ? dtu : pd_;
                }
            }
            // ---- dx+ = [A B] dux + res_b
            if (lane < MR) DS[lane] = vl ? duxk : 0.0;
            cf_syncwarp();
            double dxn;
            {
                const double *Mc = Mk + ci * MR;
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int rp = 0; rp < (NVB - 1) / 2; rp++) {
                    const cf_d2 m2 = cf_ld2(Mc + 2 * rp), d2 = cf_ld2(DS + 2 * rp);
                    s0 += m2.x * d2.x;
                    s1 += m2.y * d2.y;
                }
                s0 += Mc[NVB - 1] * DS[NVB - 1];
                dxn = xl ? (s0 + s1) + VS[R_RESB + ci] : 0.0;
                if (xl) XS[ci] = dxn;
            }
            if (need_pi) {
                cf_syncwarp();
                const double *Li = PE + ci * CF_PST;
                double z0 = pnext, z1 = 0.0;
                CF_UNROLL
                for (int cp = 0; cp < CF_NX / 2; cp++) {
                    const cf_d2 p2 = cf_ld2(Li + 2 * cp), x2 = cf_ld2(XS + 2 * cp);
                    z0 += p2.x * x2.x;
                    z1 += p2.y * x2.y;
                }
                if (CF_NX & 1) z0 += Li[CF_NX - 1] * XS[CF_NX - 1];
                if (xl) rec(k + 1)[R_DPI + ci] = z0 + z1;
            }
            dxk = dxn;
        }
        if (vl && need_pi) rec(N2)[R_DUX + lane] = ul ? 0.0 : dxk;
        const double a_p = cf_warp_max(pn_ / pd_), a_d = cf_warp_max(dn / dd);
        alpha = -(a_p > a_d ? a_p : a_d);
        cf_syncwarp();
    }

    // rhs-only backward recursion with cached P res_b (x_ocp_qp_kkt.c:1147-1245)
    CF_MEM void backward_rhs(int rm_mode, double sigma_mu)
    {
        double *TS = sm + SM_V0;
        const int VO = R_BKP, VN = B_PX - R_BKP;
        pass_begin();
        if (N2 > 0) fetch(0, N2 - 1, VO, VN);
        const bool vl = lane < NVB, ul = lane < NUB, xl = lane >= NUB && vl;
        const int ju = ul ? lane : 0, lv = vl ? lane : 0;
        double pn = 0.0;
        if (vl) {
            pn = rec(N2)[R_RESG + lane];
            rec(N2)[R_DUX + lane] = pn;
        }
        CF_NOUNROLL
        for (int k = N2 - 1; k >= 0; k--) {
            const int bf = (N2 - 1 - k) & 1;
            const bool bl = lane < CF_NU * bsz(k);
            cf_syncwarp();
            if (k > 0) fetch(bf ^ 1, k - 1, VO, VN);
            wait(bf);
            const double *VS = buf(bf) - VO;
            const double *Mk = VS + B_M, *LU = VS + B_LU;
            double Gam = 0.0, gam = 0.0;
            if (bl) bound_terms(k, VS, rm_mode, sigma_mu, Gam, gam);
            double rhs = vl ? VS[R_RESG + lv] + gam : 0.0;
            if (xl) TS[lane - NUB] = pn + VS[R_PB + lane - NUB];
            cf_syncwarp();
            if (vl) {
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int cp = 0; cp < CF_NX / 2; cp++) {
                    const cf_d2 t2 = cf_ld2(TS + 2 * cp);
                    s0 += Mk[(2 * cp) * MR + lane] * t2.x;
                    s1 += Mk[(2 * cp + 1) * MR + lane] * t2.y;
                }
                if (CF_NX & 1) s0 += Mk[(CF_NX - 1) * MR + lane] * TS[CF_NX - 1];
                rhs += s0 + s1;
            }
            // TRSV_LNN_MN(nv, nu)
            double Lr[NUB];
            CF_UNROLL
            for (int j = 0; j < NUB; j++) Lr[j] = LU[j * MR + lv];
            const double invd = LU[ju * MR + ju];
            CF_UNROLL
            for (int j = 0; j < NUB; j++) {
                const double zj = cf_shfl(rhs * invd, j);
                if (lane == j) rhs = zj;
                else if (lane > j && vl) rhs -= Lr[j] * zj;
            }
            if (vl) rec(k)[R_DUX + lane] = rhs;
            pn = rhs;
        }
        cf_syncwarp();
    }

    // COMPUTE_MU_AFF_QP (x_core_qp_ipm_aux.c:329-352) + the complementarity norm the adjusted step would produce
    CF_MEM void compute_mu_aff()
    {
        double s0 = 0.0, s1 = 0.0, pm = 0.0;
        const double aa = step_adjust(alpha);
        const int e = lane < NB2 ? lane : 0, inp = e < NUB ? e : e - NUB;
        int k = 0;
        CF_NOUNROLL
        for (; k + 3 < N2; k += 4) {   // four stages per trip: sixteen independent loads in flight per lane
            double l[4], d[4], t[4], u[4];
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                const double *r = rec(k + q);
                l[q] = r[R_LAM + e]; d[q] = r[R_DLAM + e]; t[q] = r[R_T + e]; u[q] = r[R_DT + e];
            }
            CF_UNROLL
            for (int q = 0; q < 4; q++) {
                if (lane < NB2 && inp < CF_NU * bsz(k + q)) {
                    const double v = (l[q] + alpha * d[q]) * (t[q] + alpha * u[q]);
                    if (q & 1) s1 += v;
                    else s0 += v;
                    cf_amax(pm, (l[q] + aa * d[q]) * (t[q] + aa * u[q]));
                }
            }
        }
        CF_NOUNROLL
        for (; k < N2; k++) {
            const double *r = rec(k);
            const double l = r[R_LAM + e], d = r[R_DLAM + e], t = r[R_T + e], u = r[R_DT + e];
            if (lane < NB2 && inp < CF_NU * bsz(k)) {
                s0 += (l + alpha * d) * (t + alpha * u);
                cf_amax(pm, (l + aa * d) * (t + aa * u));
            }
        }
        mu_aff = cf_warp_sum(s0 + s1) * (1.0 / (double) (2 * CF_NU * N));
        pm_max = cf_warp_max(pm);
    }

    CF_MEM bool lin_res_ok_fact() const { return true; }
    CF_MEM bool lin_res_ok_corr() const { return true; }
    template <int MODE, int NPI> CF_MEM void forward_t(const bool need_pi) { forward(need_pi); }
    static constexpr bool HAS_REFINE = false;   // iterative refinement belongs to the linear-residual diagnostics of the uncondensed program

    // =============================================================== expansion + primal update
    // d_part_cond_qp_expand_sol (x_part_cond.c:658-742 -> EXPAND_SOL, x_cond_aux.c:1820-): block states and inputs are the
    // QP solution, the states inside a block follow from the linearised dynamics; then the full step
    // (ocp_nlp_common.c:2900-2952), x_0 taking the eliminated step x0 - x_0.
    CF_MEM void expand_update(double *xg, double *ug, const double *x0g)
    {
        double *DS = sm + SM_V1, *UXB = sm + SM_V2;
        const bool xs_l = lane >= CF_NU && lane < CF_NV;      // 18-row layout of the records: lanes 4..16 carry the state
        const int ci = xs_l ? lane - CF_NU : 0;
        pass_begin();
        // the records of block i+1 and its solution travel while block i is expanded (two staging areas, two mbarriers)
        if (lane == 0) {
            cf_bulk_expect(bar, bsz(0) * CF_PREP_STAGE * 8);
            cf_bulk_g2s_raw(sm, PREP, bsz(0) * CF_PREP_STAGE * 8, bar);
        }
        double unext = lane < NVB ? rec(0)[R_UX + lane] : 0.0;
        CF_NOUNROLL
        for (int i = 0; i < N2; i++) {
            const int bs = bsz(i), k0 = kfirst(i), bf = i & 1;
            const double *ST = sm + bf * (BS * CF_PREP_STAGE);
            const double ucur = unext;
            cf_syncwarp();   // every lane is done with block i-1: its staging area and UXB may be rewritten
            if (i + 1 < N2) {
                if (lane == 0) {
                    cf_fence_proxy_async();
                    cf_bulk_expect(bar + (bf ^ 1), bsz(i + 1) * CF_PREP_STAGE * 8);
                    cf_bulk_g2s_raw(sm + (bf ^ 1) * (BS * CF_PREP_STAGE), PREP + (long) kfirst(i + 1) * CF_PREP_STAGE,
                                    bsz(i + 1) * CF_PREP_STAGE * 8, bar + (bf ^ 1));
                }
                unext = lane < NVB ? rec(i + 1)[R_UX + lane] : 0.0;
            }
            UXB[lane] = ucur;
            wait(bf);
            cf_syncwarp();
            double xs = 0.0;
            if (xs_l) xs = (i == 0) ? x0g[ci] - xg[ci] : UXB[NUB + ci];
            CF_NOUNROLL
            for (int j = 0; j < bs; j++) {
                const double *Mj = ST + j * CF_PREP_STAGE;
                const int k = k0 + j;
                double uj = 0.0;
                if (lane < CF_NU) {
                    uj = UXB[CF_NU * (bs - 1 - j) + lane];
                    ug[k * CF_NU + lane] += uj;
                }
                if (xs_l) xg[k * CF_NX + ci] += xs;
                if (j + 1 == bs) break;     // the next block's state is a QP variable
                if (lane < CF_MROWS) DS[lane] = lane < CF_NU ? uj : (xs_l ? xs : 0.0);
                cf_syncwarp();
                const double *Mc = Mj + ci * CF_MROWS;
                double s0 = 0.0, s1 = 0.0;
                CF_UNROLL
                for (int rp = 0; rp < (CF_NV - 1) / 2; rp++) {
                    const cf_d2 m2 = cf_ld2(Mc + 2 * rp), d2 = cf_ld2(DS + 2 * rp);
                    s0 += m2.x * d2.x;
                    s1 += m2.y * d2.y;
                }
                const cf_d2 m2 = cf_ld2(Mc + CF_NV - 1);   // last variable row and the b row
                s0 += m2.x * DS[CF_NV - 1];
                xs = xs_l ? m2.y + (s0 + s1) : 0.0;
                cf_syncwarp();
            }
        }
        if (xs_l) xg[N * CF_NX + ci] += rec(N2)[R_UX + NUB + ci];
        cf_syncwarp();
    }
};

// scratch slot of the condensed program: (N2 + 1) stage blocks
template <int BS>
static inline
#if !defined(CF_SIMT_EMU)
    __host__ __device__
#endif
    long
    cf_pc_scratch_doubles(int N2)
{
    return ((long) (N2 + 1) * CfPcWarpT<BS>::SB + 15) & ~15L;
}

// Feedback phase (rti_phase 2) of instance `inst` with the QP partially condensed to blk.N2 stages.
template <int BS>
CF_DEV void cf_pcond_instance(const CfParams *Pg, const CfBatchView &bv, const CfPcBlocks &blk, int inst, double *slot, double *sm,
                              unsigned &par)
{
    typedef CfPcWarpT<BS> W;
    CfParams *P = reinterpret_cast<CfParams *>(sm + W::SM_PAR);
    {
        const int lane = cf_lane();
        const double *src = reinterpret_cast<const double *>(Pg);
        double *dst = sm + W::SM_PAR;
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
    W w;
    w.bind(P, Pg, slot, sm, bv.prep + (long) inst * bv.prep_stride, bv.dts, blk);
    w.par = par;
    w.BST = bv.bnd_stage;
    w.WTAB = bv.W_stage;
    w.WD = (BS == 1) ? bv.W_dense : nullptr;
    const int N = Pg->N;
    double *xg = bv.x + (long) inst * (N + 1) * CF_NX;
    double *ug = bv.u + (long) inst * N * CF_NU;
    const double *x0g = bv.x0 + (long) inst * CF_NX;
    const double *yrefg = bv.yref + (long) inst * N * CF_NY;
    const double *yref_eg = bv.yref_e + (long) inst * CF_NX;
    unsigned long long *prof = bv.prof;
    {
        CF_PROF_BEGIN();
        w.condense(xg, ug, x0g, yrefg, yref_eg);
        CF_PROF_END(CF_PROF_LIN);
    }
    int iters = 0;
    const int qp_status = cf_ipm_solve(w, iters, prof);
    CF_PROF_BEGIN();
    int status = CF_ACADOS_SUCCESS;
    if (qp_status == 0 || qp_status == 1) w.expand_update(xg, ug, x0g);
    else status = CF_ACADOS_QP_FAILURE;
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
}
