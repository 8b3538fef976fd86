// TEST HARNESS ONLY: runs the CUDA warp program (crazyflie_nmpc_b200/csrc/cf_rti_warp.h)
// on the host by executing the 32 lanes of a warp as lock-step fibers.  Every warp
// primitive (shuffle, __syncwarp) is a rendez-vous of all 32 fibers, so data races and
// divergent-barrier bugs in the kernel source show up here without a GPU.  This file is
// compiled only by tests/ (see tests/simt_emu/build.py); the product library never
// contains it and has no CPU path.
#define CF_SIMT_EMU 1
#include <ucontext.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "cf_rti_warp.h"
#include "cf_pcond_warp.h"

namespace cfemu {
struct Warp
{
    ucontext_t ctx[32], main_ctx;
    char *stacks[32];
    int cur = 0;
    int done[32];
    int line0 = 0;
    double xch[32];
    void (*fn)(void *) = nullptr;
    void *arg = nullptr;
};
static thread_local Warp *W = nullptr;
static const size_t STACK = 512 * 1024;

int lane() { return W->cur; }

static void switch_next()
{
    Warp *w = W;
    int from = w->cur;
    for (int step = 1; step <= 32; step++) {
        int to = (from + step) & 31;
        if (!w->done[to]) {
            if (to == from) return;
            w->cur = to;
            swapcontext(&w->ctx[from], &w->ctx[to]);
            return;
        }
    }
}

void barrier(int line)
{
    Warp *w = W;
    if (w->cur == 0) w->line0 = line;
    else if (line != w->line0) {
        fprintf(stderr, "cfemu: divergent warp barrier: lane %d at line %d, lane 0 at line %d\n", w->cur, line, w->line0);
        abort();
    }
    switch_next();
}

double exch(double v, int src, int line)
{
    Warp *w = W;
    w->xch[w->cur] = v;
    barrier(line);
    double r = w->xch[src & 31];
    barrier(-line);
    return r;
}

int atomic_add(int *p, int v)
{
    int old = *p;
    *p += v;
    return old;
}

// ---- TMA bulk copies.  global->shared: the destination is poisoned when the copy is issued and
// filled only when the warp waits on the mbarrier, so reading a staged block before waiting,
// or waiting on the wrong barrier/parity, produces NaNs instead of silently working.
struct PendingLoad { void *dst; const void *src; int bytes; uint64_t *bar; };
struct PendingStore { const void *src; std::vector<char> snapshot; };
static thread_local std::vector<PendingLoad> g_loads;
static thread_local std::vector<PendingStore> g_stores;
static thread_local std::vector<std::pair<uint64_t *, int>> g_expect;

void bulk_expect(uint64_t *bar, int bytes, int line)
{
    for (auto &e : g_expect)
        if (e.first == bar) { fprintf(stderr, "cfemu: line %d: mbarrier re-armed while a phase is pending\n", line); abort(); }
    g_expect.push_back({bar, bytes});
}
void bulk_g2s(void *dst, const void *src, int bytes, uint64_t *bar, int line)
{
    if ((((uintptr_t) dst) & 15) || (((uintptr_t) src) & 15) || (bytes & 15)) {
        fprintf(stderr, "cfemu: line %d: bulk copy needs 16-byte aligned addresses and size\n", line); abort();
    }
    double *d = (double *) dst;
    for (int i = 0; i < bytes / 8; i++) d[i] = std::nan("");
    g_loads.push_back({dst, src, bytes, bar});
}
void bulk_wait(uint64_t *bar, unsigned parity, int line)
{
    // phase bit lives in the barrier word: completed phases counted modulo 2
    barrier(line);
    if (W->cur == 0 || true) {
        // the first lane to arrive completes the phase; later lanes see it done
        if ((*bar & 1u) == parity) {
            int expect = -1, got = 0;
            for (size_t i = 0; i < g_expect.size(); i++)
                if (g_expect[i].first == bar) { expect = g_expect[i].second; g_expect.erase(g_expect.begin() + i); break; }
            if (expect < 0) { fprintf(stderr, "cfemu: line %d: wait on an mbarrier that was never armed (deadlock on the GPU)\n", line); abort(); }
            for (size_t i = 0; i < g_loads.size();) {
                if (g_loads[i].bar == bar) { memcpy(g_loads[i].dst, g_loads[i].src, g_loads[i].bytes); got += g_loads[i].bytes; g_loads.erase(g_loads.begin() + i); }
                else i++;
            }
            if (got != expect) { fprintf(stderr, "cfemu: line %d: expect_tx %d bytes but %d were copied\n", line, expect, got); abort(); }
            *bar ^= 1u;
        }
    }
    barrier(-line);
}
void bulk_s2g(void *dst, const void *src, int bytes, int line)
{
    if ((((uintptr_t) dst) & 15) || (((uintptr_t) src) & 15) || (bytes & 15)) {
        fprintf(stderr, "cfemu: line %d: bulk copy needs 16-byte aligned addresses and size\n", line); abort();
    }
    memcpy(dst, src, bytes);
    PendingStore p; p.src = src; p.snapshot.assign((const char *) src, (const char *) src + bytes);
    g_stores.push_back(std::move(p));
}
void bulk_s2g_wait(int max_pending, int line)
{
    // the shared source of a store that may still be in flight must not have been modified
    while ((int) g_stores.size() > max_pending) {
        PendingStore &p = g_stores.front();
        if (memcmp(p.src, p.snapshot.data(), p.snapshot.size()) != 0) {
            fprintf(stderr, "cfemu: line %d: shared source of a bulk store was overwritten before wait_group\n", line); abort();
        }
        g_stores.erase(g_stores.begin());
    }
}

// fp64 mma.sync.m8n8k4: every lane publishes its A and B fragment element, then computes its two outputs
static thread_local double g_fa[32], g_fb[32];
void dmma(double &d0, double &d1, double a, double b, int line)
{
    Warp *w = W;
    g_fa[w->cur] = a; g_fb[w->cur] = b;
    barrier(line);
    const int g = w->cur >> 2, q = w->cur & 3;
    double s0 = d0, s1 = d1;
    for (int k = 0; k < 4; k++) {
        const double A = g_fa[g * 4 + k];
        s0 = std::fma(A, g_fb[(2 * q) * 4 + k], s0);
        s1 = std::fma(A, g_fb[(2 * q + 1) * 4 + k], s1);
    }
    barrier(-line);
    d0 = s0; d1 = s1;
}

static void trampoline()
{
    Warp *w = W;
    w->fn(w->arg);
    w->done[w->cur] = 1;
    int from = w->cur;
    for (int step = 1; step < 32; step++) {
        int to = (from + step) & 31;
        if (!w->done[to]) { w->cur = to; setcontext(&w->ctx[to]); }
    }
    setcontext(&w->main_ctx);
}

static void run_warp(void (*fn)(void *), void *arg)
{
    Warp *w = new Warp();
    W = w;
    w->fn = fn;
    w->arg = arg;
    for (int i = 0; i < 32; i++) {
        w->done[i] = 0;
        w->stacks[i] = (char *) malloc(STACK);
        getcontext(&w->ctx[i]);
        w->ctx[i].uc_stack.ss_sp = w->stacks[i];
        w->ctx[i].uc_stack.ss_size = STACK;
        w->ctx[i].uc_link = nullptr;
        makecontext(&w->ctx[i], trampoline, 0);
    }
    w->cur = 0;
    swapcontext(&w->main_ctx, &w->ctx[0]);
    for (int i = 0; i < 32; i++) free(w->stacks[i]);
    delete w;
    W = nullptr;
}
}  // namespace cfemu

// multiplier output [B][cf_mult_stride(N)] for the next runs (NULL: off), see CfBatchView::mult
static double *g_mult = nullptr;
extern "C" void cfemu_set_mult(double *buf) { g_mult = buf; }
extern "C" long cfemu_mult_stride(int N) { return cf_mult_stride(N); }

struct Job
{
    const CfParams *P;
    CfBatchView bv;
    int inst;
    double *slot;
    double *sm;
};
#if CF_CRAZYFLIE   // the tuned uncondensed program exists for nx = 13, nu = 4 only
static void job_fn(void *a)
{
    Job *j = (Job *) a;
    cf_warp_init_smem(j->sm);
    unsigned par = 0;
    cf_rti_instance(j->P, j->bv, j->inst, j->slot, j->sm, par);
}
#endif
template <int PH, bool VDT>
static void job_fn_t(void *a)
{
    Job *j = (Job *) a;
    cf_warp_init_smem(j->sm);
    unsigned par = 0;
    cf_rti_instance<PH, VDT>(j->P, j->bv, j->inst, j->slot, j->sm, par);
}

static const double *g_bnd_stage = nullptr;
// per-stage input boxes [N][8] for the next cfemu_rti_general calls (NULL: none)
extern "C" void cfemu_set_stage_bounds(const double *tab) { g_bnd_stage = tab; }
// per-instance parameter arrays {W, W_e, lbu, ubu, lbu0, ubu0} for the next cfemu_rti_general calls (NULL: none)
static const double *const *g_per_inst = nullptr;
extern "C" void cfemu_set_per_inst(const double *const *p) { g_per_inst = p; }
// lin_res_check of the next cfemu_rti_batch calls: 1 = flags only (default), 2 = with iterative refinement
static int g_lin_res_check = 1;
extern "C" void cfemu_set_lin_res_check(int v) { g_lin_res_check = v; }
#if CF_CRAZYFLIE
extern "C" long cfemu_scratch_doubles(int N) { return cf_scratch_layout(N).total; }
extern "C" void cfemu_scratch_offsets(int N, long *out)
{
    CfScratchLayout s = cf_scratch_layout(N);
    long v[12] = {s.total, CF_SB, B_M, B_LU, B_PX, R_UX, R_PI, R_RQ, R_B, R_RESG, R_DUX, R_D};
    memcpy(out, v, sizeof v);
}

// params: Wdiag[17], WNdiag[13], lbu[4], ubu[4] (38 doubles) or NULL for the reference values
// per_inst: six pointers {W [B][17], W_e [B][13], lbu, ubu, lbu0, ubu0 [B][4]}, any of them (or the array) NULL
extern "C" int cfemu_rti_batch2(int B, int N, double Ts, const double *params, int max_ipm_iter, const double *x0,
                                const double *yref, const double *yref_e, double *x, double *u, int *status,
                                int *qp_iter, int *qp_status, int *flags, double *res, double *scratch_out,
                                int nthreads, const double *const *per_inst);
extern "C" int cfemu_rti_batch(int B, int N, double Ts, const double *params, int max_ipm_iter, const double *x0,
                               const double *yref, const double *yref_e, double *x, double *u, int *status,
                               int *qp_iter, int *qp_status, int *flags, double *res, double *scratch_out,
                               int nthreads)
{
    return cfemu_rti_batch2(B, N, Ts, params, max_ipm_iter, x0, yref, yref_e, x, u, status, qp_iter, qp_status, flags, res,
                            scratch_out, nthreads, nullptr);
}
extern "C" int cfemu_rti_batch2(int B, int N, double Ts, const double *params, int max_ipm_iter, const double *x0,
                                const double *yref, const double *yref_e, double *x, double *u, int *status,
                                int *qp_iter, int *qp_status, int *flags, double *res, double *scratch_out,
                                int nthreads, const double *const *per_inst)
{
    CfParams P;
    static const double Q[13] = {120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0};
    for (int i = 0; i < 13; i++) { P.Wdiag[i] = Q[i]; P.WNdiag[i] = 50 * Q[i]; }
    for (int i = 0; i < 4; i++) { P.Wdiag[13 + i] = 0.06; P.lbu[i] = 0; P.ubu[i] = 22; }
    if (params) {
        memcpy(P.Wdiag, params, 17 * 8); memcpy(P.WNdiag, params + 17, 13 * 8);
        memcpy(P.lbu, params + 30, 4 * 8); memcpy(P.ubu, params + 34, 4 * 8);
    }
    memcpy(P.lbu0, P.lbu, sizeof P.lbu); memcpy(P.ubu0, P.ubu, sizeof P.ubu);
    P.Ts = Ts; P.N = N; P.max_ipm_iter = max_ipm_iter > 0 ? max_ipm_iter : CF_ITER_MAX;
    P.lin_res_check = g_lin_res_check; P.pad_ = 0;
    const long stride = cf_scratch_layout(N).total;
    CfBatchView bv = {};
    bv.mult = g_mult; bv.mult_stride = cf_mult_stride(N);
    bv.B = B; bv.first = 0; bv.ready = nullptr; bv.x0 = x0; bv.yref = yref; bv.yref_e = yref_e; bv.x = x; bv.u = u; bv.status = status;
    bv.qp_iter = qp_iter; bv.qp_status = qp_status; bv.flags = flags; bv.res = res; bv.scratch = nullptr;
    bv.scratch_stride = stride; bv.counter = nullptr;
    bv.W_b = per_inst ? per_inst[0] : nullptr; bv.WN_b = per_inst ? per_inst[1] : nullptr;
    bv.lbu_b = per_inst ? per_inst[2] : nullptr; bv.ubu_b = per_inst ? per_inst[3] : nullptr;
    bv.lbu0_b = per_inst ? per_inst[4] : nullptr; bv.ubu0_b = per_inst ? per_inst[5] : nullptr;
    bv.prof = nullptr;
    bv.dts = nullptr; bv.prep = nullptr; bv.prep_stride = 0; bv.bnd_stage = nullptr;
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&]() {
            std::vector<double> slot(stride + 2, 0.0), sm(CF_SM_DOUBLES + 2, 0.0);
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= B) break;
                // poison the scratch so that reads of never-written data are visible
                for (auto &v : slot) v = std::nan("");
                double *slot_a = (double *) ((((uintptr_t) slot.data()) + 15) & ~(uintptr_t) 15), *sm_a = (double *) ((((uintptr_t) sm.data()) + 15) & ~(uintptr_t) 15);
                for (int q = 0; q < CF_SM_DOUBLES; q++) sm_a[q] = std::nan("");
                Job j{&P, bv, i, slot_a, sm_a};
                cfemu::run_warp(job_fn, &j);
                if (scratch_out) memcpy(scratch_out + (size_t) i * stride, slot_a, stride * 8);
            }
        });
    for (auto &t : th) t.join();
    return 0;
}

// General variants of the warp program: per-interval time steps `dts` (NULL = uniform Ts) and, with split != 0, the
// real-time iteration as two phases -- preparation with x0, then (scratch slot and shared memory poisoned in between, as
// another instance would have used them) feedback with x0_fb (NULL = x0).
extern "C" int cfemu_rti_general(int B, int N, double Ts, const double *dts, int split, const double *x0, const double *x0_fb,
                                 const double *yref, const double *yref_e, double *x, double *u, int *status, int *qp_iter,
                                 int *qp_status, int *flags, double *res, int nthreads)
{
    CfParams P;
    static const double Q[13] = {120, 100, 100, 1e-3, 1e-3, 1e-3, 1e-3, 0.7, 1.0, 4.0, 1e-5, 1e-5, 10.0};
    for (int i = 0; i < 13; i++) { P.Wdiag[i] = Q[i]; P.WNdiag[i] = 50 * Q[i]; }
    for (int i = 0; i < 4; i++) { P.Wdiag[13 + i] = 0.06; P.lbu[i] = P.lbu0[i] = 0; P.ubu[i] = P.ubu0[i] = 22; }
    std::vector<double> dtv(N, Ts);
    if (dts) dtv.assign(dts, dts + N);
    P.Ts = dtv[0]; P.N = N; P.max_ipm_iter = CF_ITER_MAX; P.lin_res_check = g_lin_res_check; P.pad_ = 0;
    const long stride = cf_scratch_layout(N).total, pstride = cf_prep_stride(N);
    std::vector<double> prep((size_t) B * pstride + 2, std::nan(""));
    CfBatchView bv = {};
    bv.mult = g_mult; bv.mult_stride = cf_mult_stride(N);
    memset(&bv, 0, sizeof bv);
    bv.B = B; bv.x0 = x0; bv.yref = yref; bv.yref_e = yref_e; bv.x = x; bv.u = u; bv.status = status;
    bv.qp_iter = qp_iter; bv.qp_status = qp_status; bv.flags = flags; bv.res = res; bv.scratch_stride = stride;
    bv.dts = dtv.data();
    bv.prep = (double *) ((((uintptr_t) prep.data()) + 15) & ~(uintptr_t) 15);
    bv.prep_stride = pstride;
    bv.bnd_stage = g_bnd_stage;
    if (g_per_inst) {
        bv.W_b = g_per_inst[0]; bv.WN_b = g_per_inst[1]; bv.lbu_b = g_per_inst[2]; bv.ubu_b = g_per_inst[3];
        bv.lbu0_b = g_per_inst[4]; bv.ubu0_b = g_per_inst[5];
    }
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&]() {
            std::vector<double> slot(stride + 2, 0.0), sm(CF_SM_DOUBLES + 2, 0.0);
            double *slot_a = (double *) ((((uintptr_t) slot.data()) + 15) & ~(uintptr_t) 15), *sm_a = (double *) ((((uintptr_t) sm.data()) + 15) & ~(uintptr_t) 15);
            auto poison = [&]() {
                for (long q = 0; q < stride; q++) slot_a[q] = std::nan("");
                for (int q = 0; q < CF_SM_DOUBLES; q++) sm_a[q] = std::nan("");
            };
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= B) break;
                poison();
                Job j{&P, bv, i, slot_a, sm_a};
                if (!split) cfemu::run_warp(job_fn_t<CF_PH_BOTH, true>, &j);
                else {
                    cfemu::run_warp(job_fn_t<CF_PH_PREPARATION, true>, &j);
                    poison();
                    if (x0_fb) j.bv.x0 = x0_fb;
                    cfemu::run_warp(job_fn_t<CF_PH_FEEDBACK, true>, &j);
                }
            }
        });
    for (auto &t : th) t.join();
    return 0;
}


#endif  // CF_CRAZYFLIE

// Partially condensed feedback (cf_pcond_warp.h): preparation phase of the general warp program, then -- scratch slot
// and shared memory poisoned in between -- the condensed feedback with block size BS = ceil(N / N2).
struct PcJob
{
    const CfParams *P;
    CfBatchView bv;
    CfPcBlocks blk;
    int inst;
    double *slot;
    double *sm;
};
template <int BS>
static void pc_job_fn(void *a)
{
    PcJob *j = (PcJob *) a;
    cf_warp_init_smem(j->sm + CfPcWarpT<BS>::SM_BAR - CF_SM_BAR);   // the two mbarriers of the condensed layout
    unsigned par = 0;
    cf_pcond_instance<BS>(j->P, j->bv, j->blk, j->inst, j->slot, j->sm, par);
}
// full weight matrices in the kernel's table format [(N+1)][2][17*17] ([u;x] order: Hessian factor product, W) for the next
// cfemu_rti_pcond calls (NULL: diagonal weights); with it N2 = N selects the block-size-1 program
static const double *g_wdense = nullptr;
extern "C" void cfemu_set_dense_weights(const double *tab) { g_wdense = tab; }
extern "C" int cfemu_rti_pcond(int B, int N, double Ts, int N2, const double *params, const double *x0, const double *yref,
                               const double *yref_e, double *x, double *u, int *status, int *qp_iter, int *qp_status, int *flags,
                               double *res, int nthreads, const double *const *per_inst)
{
    // block size 1 serves the full weight matrices and, for models other than the Crazyflie, the whole feedback phase
    if (N2 < 1 || N2 > N || (N2 == N && !g_wdense && CF_CRAZYFLIE)) return -1;
    const CfPcBlocks blk = cf_pc_blocks(N, N2);
    const int BS = blk.n_big ? blk.bs0 + 1 : blk.bs0;
    if (BS < 1 || BS > (CF_CRAZYFLIE ? 3 : 1)) return -2;
    CfParams P;
    for (int i = 0; i < CF_NY; i++) P.Wdiag[i] = CfSpec::W[i];     // the generated OCP description (any model)
    for (int i = 0; i < CF_NX; i++) P.WNdiag[i] = CfSpec::W_e[i];
    for (int i = 0; i < CF_NU; i++) { P.lbu[i] = CfSpec::lbu[i]; P.ubu[i] = CfSpec::ubu[i]; }
    if (params) {
        memcpy(P.Wdiag, params, CF_NY * 8); memcpy(P.WNdiag, params + CF_NY, CF_NX * 8);
        memcpy(P.lbu, params + CF_NY + CF_NX, CF_NU * 8); memcpy(P.ubu, params + CF_NY + CF_NX + CF_NU, CF_NU * 8);
    }
    memcpy(P.lbu0, P.lbu, sizeof P.lbu); memcpy(P.ubu0, P.ubu, sizeof P.ubu);
    P.Ts = Ts; P.N = N; P.max_ipm_iter = CF_ITER_MAX; P.lin_res_check = 0; P.pad_ = 0;
    const long stride0 = cf_scratch_layout(N).total, pstride = cf_prep_stride(N);
#if CF_CRAZYFLIE
    const long stride1 = BS == 3 ? cf_pc_scratch_doubles<3>(N2) : (BS == 2 ? cf_pc_scratch_doubles<2>(N2) : cf_pc_scratch_doubles<1>(N2));
    const int smd = BS == 3 ? (int) CfPcWarpT<3>::SM_DOUBLES : (BS == 2 ? (int) CfPcWarpT<2>::SM_DOUBLES : (int) CfPcWarpT<1>::SM_DOUBLES);
#else
    const long stride1 = cf_pc_scratch_doubles<1>(N2);
    const int smd = (int) CfPcWarpT<1>::SM_DOUBLES;
#endif
    const long stride = stride0 > stride1 ? stride0 : stride1;
    const int smn = smd > CF_SM_DOUBLES ? smd : CF_SM_DOUBLES;
    std::vector<double> dtv(N, Ts);
    std::vector<double> prep((size_t) B * pstride + 2, std::nan(""));
    CfBatchView bv = {};
    bv.mult = g_mult; bv.mult_stride = cf_mult_stride(N);
    memset(&bv, 0, sizeof bv);
    bv.B = B; bv.x0 = x0; bv.yref = yref; bv.yref_e = yref_e; bv.x = x; bv.u = u; bv.status = status;
    bv.qp_iter = qp_iter; bv.qp_status = qp_status; bv.flags = flags; bv.res = res; bv.scratch_stride = stride;
    bv.dts = dtv.data();
    bv.prep = (double *) ((((uintptr_t) prep.data()) + 15) & ~(uintptr_t) 15);
    bv.prep_stride = pstride;
    bv.bnd_stage = g_bnd_stage;
    bv.W_dense = g_wdense;
    if (per_inst) {
        bv.W_b = per_inst[0]; bv.WN_b = per_inst[1]; bv.lbu_b = per_inst[2]; bv.ubu_b = per_inst[3];
        bv.lbu0_b = per_inst[4]; bv.ubu0_b = per_inst[5];
    }
    if (nthreads < 1) nthreads = 1;
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&]() {
            std::vector<double> slot(stride + 2, 0.0), sm(smn + 2, 0.0);
            double *slot_a = (double *) ((((uintptr_t) slot.data()) + 15) & ~(uintptr_t) 15), *sm_a = (double *) ((((uintptr_t) sm.data()) + 15) & ~(uintptr_t) 15);
            auto poison = [&]() {
                for (long q = 0; q < stride; q++) slot_a[q] = std::nan("");
                for (int q = 0; q < smn; q++) sm_a[q] = std::nan("");
            };
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= B) break;
                poison();
                Job j{&P, bv, i, slot_a, sm_a};
                cfemu::run_warp(job_fn_t<CF_PH_PREPARATION, true>, &j);
                poison();
                PcJob pj{&P, bv, blk, i, slot_a, sm_a};
                #if CF_CRAZYFLIE
                cfemu::run_warp(BS == 3 ? pc_job_fn<3> : (BS == 2 ? pc_job_fn<2> : pc_job_fn<1>), &pj);
#else
                cfemu::run_warp(pc_job_fn<1>, &pj);
#endif
            }
        });
    for (auto &t : th) t.join();
    return 0;
}

extern "C" void cfemu_dims(int *nx, int *nu) { *nx = CF_NX; *nu = CF_NU; }
