// Host side of the batch C-ABI (include/cfnmpc.h) and the sm_100a kernels behind it.
// C++ above a C boundary, as the reference's solver glue is C
// (acados_template/c_templates_tera/acados_solver.in.c).  No torch types, no CPU path:
// every compute call ends in a kernel launch or returns CFNMPC_ECUDA.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/cfnmpc.h"
#include "cf_kernels.h"
#include "cf_loop_kernels.h"


// iterate initialisation of the generated solver (acados_solver.in.c:2323-2352)
__global__ void cf_init_iterate_kernel(double *x, double *u, long nx_tot, long nu_tot)
{
    const long t = (long) blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nx_tot) x[t] = ((t % CF_NX) == 3) ? 1.0 : 0.0;
    if (t < nu_tot) u[t] = 0.0;
}

// ------------------------------------------------------------------ host state
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(CFNMPC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

struct cfnmpc_batch
{
    int B = 0, N = 0, device = 0;
    CfParams P;
    CfBatchView bv;
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_free = nullptr;
    int *d_ready = nullptr, *h_ready = nullptr;   // upload front of cfnmpc_batch_solve_from_host (device counter, pinned staging)
    double *d_x0 = nullptr, *d_yref = nullptr, *d_yref_e = nullptr, *d_x = nullptr, *d_u = nullptr, *d_res = nullptr;
    double *d_scratch = nullptr, *d_stage = nullptr;
    // closed-loop driver state (cf_loop_kernels.h)
    int *d_policy = nullptr, *d_titer = nullptr, *d_motors = nullptr;
    double *d_setpoint = nullptr, *d_traj = nullptr, *d_euler = nullptr, *d_twist = nullptr;
    int n_traj = 0;
    double uss = 0.0;
    unsigned long long *d_prof = nullptr;
    double *d_Wb = nullptr, *d_WNb = nullptr, *d_lbub = nullptr, *d_ubub = nullptr, *d_lbu0b = nullptr, *d_ubu0b = nullptr;
    int *d_status = nullptr, *d_qp_iter = nullptr, *d_qp_status = nullptr, *d_flags = nullptr, *d_counter = nullptr;
    int grid = 0, blocks_per_sm = 0, sm_count = 0, n_slots = 0, regs = 0, minb = 3, wpb = 4;
    void (*kernel)(const CfParams, const CfBatchView) = nullptr;
    // general variants (default launch shape): per-interval time steps, split phases
    void (*kernel_vdt)(const CfParams, const CfBatchView) = nullptr;
    void (*kernel_fb_g4)(const CfParams, const CfBatchView) = nullptr;   // general feedback kernel at 4 x 4 (two-kernel step of the general path)
    int grid_fb_g4 = 0;
    void (*kernel_prep)(const CfParams, const CfBatchView) = nullptr;
    void (*kernel_fb)(const CfParams, const CfBatchView) = nullptr;
    // the two halves specialised for the uniform grid: cfnmpc_batch_solve as two launches (option "two_kernels")
    void (*kernel_prep_u)(const CfParams, const CfBatchView) = nullptr;
    void (*kernel_fb_u)(const CfParams, const CfBatchView) = nullptr;
    bool two_kernels = true;
    // partial condensing (option "qp_cond_N"): 0 = off (every block holds one stage, the reference's own configuration)
    int cond_N = 0, pc_bs = 0, pc_wpb = 4, pc_minb = 2, grid_pc = 0, pc_regs = 0, pc_blocks_per_sm = 0;
    size_t smem_pc = 0;
    CfPcBlocks pcb;
    void (*kernel_pc)(const CfParams, const CfBatchView, const CfPcBlocks) = nullptr;
    int grid_prep_u = 0, prep_minb = 3;
    int grid_fb = 0, fb_minb = 4, fb_wpb = 4, fb_regs = 0, fb_blocks_per_sm = 0;
    size_t smem_fb = 0;
    cudaEvent_t ev_mid = nullptr;     // between the two launches of a two-kernel step
    bool mid_valid = false;
    size_t smem_general = 0;
    int grid_general = 0;
    double *d_dts = nullptr, *d_prep = nullptr;
    double *h_dts = nullptr;          // host copy of the time grid (N doubles)
    double *d_bst = nullptr;          // per-stage input boxes [N][8] (allocated on first use)
    double *d_wdense = nullptr;       // full weight matrices per stage [(N+1)][2][289] (field "W_dense_table")
    bool dense_w = false;
    double *d_mult = nullptr;         // multiplier output [B][cf_mult_stride(N)] (option "multipliers")
    double *d_wst = nullptr;          // per-stage weights [N+1][17] (allocated on first use); while set, the general kernels run
    bool vdt_grid = false, wst = false, prepared = false;   // non-uniform time grid / per-stage weights: the general kernels
    bool itref = false;                                      // lin_res_check >= 1: diagnostics / iterative refinement (general kernels)
    bool vdt = false;                                        // = vdt_grid || wst || itref
    size_t smem = 0;
    long long launches = 0;
    bool timed = false;
};

static void default_params(CfParams &P, int N, double Ts)
{
    // generate_c_code.py:61-84 (Q, R), :113 (W_e = 50 Q), :133-134 (0 <= u <= 22), through tools/gen_spec.py
    for (int i = 0; i < CF_NY; i++) P.Wdiag[i] = CfSpec::W[i];
    for (int i = 0; i < CF_NX; i++) P.WNdiag[i] = CfSpec::W_e[i];
    for (int i = 0; i < CF_NU; i++) { P.lbu[i] = P.lbu0[i] = CfSpec::lbu[i]; P.ubu[i] = P.ubu0[i] = CfSpec::ubu[i]; }
    P.Ts = Ts; P.N = N; P.max_ipm_iter = CF_ITER_MAX; P.lin_res_check = 0; P.pad_ = 0;
}

extern "C" const char *cfnmpc_last_error(void) { return g_err.c_str(); }
extern "C" const char *cfnmpc_version(void) { return "crazyflie_nmpc_b200 0.2 (sm_100a)"; }
// sizes the library was generated for and the horizon / final time of its OCP description (the same entry point exists in
// the generic-model libraries, cfnmpc_generic.cu)
extern "C" int cfnmpc_model_dims(int *nx, int *nu, int *N, double *Tf)
{
    if (nx) *nx = CF_NX;
    if (nu) *nu = CF_NU;
    if (N) *N = CF_SPEC_N;
    if (Tf) *Tf = CF_SPEC_TF;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_destroy(cfnmpc_batch *h)
{
    if (!h) return CFNMPC_OK;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_x0, h->d_yref, h->d_yref_e, h->d_x, h->d_u, h->d_res, h->d_scratch, h->d_stage,
                    h->d_status, h->d_qp_iter, h->d_qp_status, h->d_flags, h->d_counter,
                    h->d_policy, h->d_titer, h->d_motors, h->d_setpoint, h->d_traj, h->d_euler, h->d_twist,
                    h->d_prof, h->d_Wb, h->d_WNb, h->d_lbub, h->d_ubub, h->d_lbu0b, h->d_ubu0b, h->d_dts, h->d_prep, h->d_bst, h->d_wst, h->d_mult, h->d_wdense};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_free) cudaEventDestroy(h->ev_free);
    if (h->ev_mid) cudaEventDestroy(h->ev_mid);
    if (h->d_ready) cudaFree(h->d_ready);
    if (h->h_ready) cudaFreeHost(h->h_ready);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    free(h->h_dts);
    delete h;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_create(int batch, int N, double Ts, int device, cfnmpc_batch **out)
{
    if (!out) return fail(CFNMPC_EINVAL, "cfnmpc_batch_create: out is NULL");
    *out = nullptr;
    if (batch < 1 || N < 1 || N > 4096 || !(Ts > 0)) return fail(CFNMPC_EINVAL, "cfnmpc_batch_create: need batch >= 1, 1 <= N <= 4096, Ts > 0");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(CFNMPC_EINVAL, "cfnmpc_batch_create: no such CUDA device");
    CK(cudaSetDevice(device));
    cfnmpc_batch *h = new cfnmpc_batch();
    h->B = batch; h->N = N; h->device = device;
    default_params(h->P, N, Ts);
#define CKH(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            cfnmpc_batch_destroy(h);                                                                \
            return fail(CFNMPC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
        }                                                                                           \
    } while (0)
    CKH(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CKH(cudaEventCreate(&h->ev0));
    CKH(cudaEventCreate(&h->ev1));
    CKH(cudaEventCreate(&h->ev_mid));
    cudaDeviceProp prop;
    CKH(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    // launch shape: CFNMPC_WARPS_PER_BLOCK x CFNMPC_MIN_BLOCKS (default 4 x 3 = 12 warps per SM at 168 registers, the
    // measured optimum: profiles/README.md)
    if (const char *e = getenv("CFNMPC_MIN_BLOCKS")) h->minb = atoi(e);
    if (const char *e = getenv("CFNMPC_WARPS_PER_BLOCK")) h->wpb = atoi(e);
    const int shape = h->wpb * 100 + h->minb;
    switch (shape) {
    case 404: h->kernel = cf_rti_kernel<4, 4>; break;
    case 405: h->kernel = cf_rti_kernel<4, 5>; break;
    case 209: h->kernel = cf_rti_kernel<2, 9>; break;

    default: h->wpb = 4; h->minb = 3; h->kernel = cf_rti_kernel<4, 3>; break;
    }
    h->smem = (size_t) h->wpb * CF_SM_DOUBLES * sizeof(double);
    CKH(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem));
    h->kernel_vdt = cf_rti_kernel<4, 3, CF_PH_BOTH, true>;
    h->kernel_prep = cf_rti_kernel<4, 3, CF_PH_PREPARATION, true>;
    h->kernel_fb = cf_rti_kernel<4, 3, CF_PH_FEEDBACK, true>;
    h->kernel_prep_u = cf_rti_kernel<4, 3, CF_PH_PREPARATION, false>;
    h->kernel_fb_u = cf_rti_kernel<4, 3, CF_PH_FEEDBACK, false>;
    // Default step = two launches (profiles/README.md, "two kernels"): the preparation at 4 x 3 blocks per SM (168 registers,
    // the linearisation needs them), the feedback at 4 x 4 (16 warps per SM at 128 registers).  CFNMPC_TWO_KERNELS=0 /
    // option "two_kernels" 0 returns to the single fused kernel, CFNMPC_FB_MIN_BLOCKS=3 to a 4 x 3 feedback kernel.
    if (const char *e = getenv("CFNMPC_FB_MIN_BLOCKS")) h->fb_minb = atoi(e);
    if (h->fb_minb == 4) h->kernel_fb_u = cf_rti_kernel<4, 4, CF_PH_FEEDBACK, false>;
    else h->fb_minb = 3;
    // other shapes of the benchmarked feedback kernel, for the occupancy experiments of profiles/README.md:
    // CFNMPC_FB_SHAPE = 100 * warps per block + blocks per SM (18 / 20 warps per SM at 96 registers: 6 % slower than 4 x 4;
    // a 112-register cap does not give 18 warps: registers are allocated per warp in units that round it up to 128)
    if (const char *e = getenv("CFNMPC_FB_SHAPE")) {
        switch (atoi(e)) {
        case 603: h->kernel_fb_u = cf_rti_kernel<6, 3, CF_PH_FEEDBACK, false>; h->fb_wpb = 6; h->fb_minb = 3; break;
        case 306: h->kernel_fb_u = cf_rti_kernel<3, 6, CF_PH_FEEDBACK, false>; h->fb_wpb = 3; h->fb_minb = 6; break;
        case 405: h->kernel_fb_u = cf_rti_kernel<4, 5, CF_PH_FEEDBACK, false>; h->fb_wpb = 4; h->fb_minb = 5; break;
        default: break;
        }
    }
    h->smem_fb = (size_t) h->fb_wpb * CF_SM_DOUBLES * sizeof(double);
    if (const char *e = getenv("CFNMPC_PREP_MIN_BLOCKS"))
        if (atoi(e) == 4) { h->kernel_prep_u = cf_rti_kernel<4, 4, CF_PH_PREPARATION, false>; h->prep_minb = 4; }
    h->smem_general = (size_t) 4 * CF_SM_DOUBLES * sizeof(double);
    CKH(cudaFuncSetAttribute(h->kernel_prep_u, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_general));
    CKH(cudaFuncSetAttribute(h->kernel_fb_u, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_fb));
    if (const char *e = getenv("CFNMPC_TWO_KERNELS")) h->two_kernels = atoi(e) != 0;
    {
        cudaFuncAttributes fb;
        CKH(cudaFuncGetAttributes(&fb, h->kernel_fb_u));
        h->fb_regs = fb.numRegs;
        CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->fb_blocks_per_sm, h->kernel_fb_u, h->fb_wpb * 32, h->smem_fb));
        if (h->fb_blocks_per_sm < 1) { cfnmpc_batch_destroy(h); return fail(CFNMPC_ECUDA, "feedback kernel does not fit on an SM"); }
    }
    CKH(cudaFuncSetAttribute(h->kernel_vdt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_general));
    h->kernel_fb_g4 = cf_rti_kernel<4, 4, CF_PH_FEEDBACK, true>;
    CKH(cudaFuncSetAttribute(h->kernel_fb_g4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_general));
    CKH(cudaFuncSetAttribute(h->kernel_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_general));
    CKH(cudaFuncSetAttribute(h->kernel_fb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_general));
    cudaFuncAttributes fa;
    CKH(cudaFuncGetAttributes(&fa, h->kernel));
    h->regs = fa.numRegs;
    CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, h->kernel, h->wpb * 32, h->smem));
    if (h->blocks_per_sm < 1) { cfnmpc_batch_destroy(h); return fail(CFNMPC_ECUDA, "kernel does not fit on an SM"); }
    // persistent grid: every SM fully occupied, but never more warps than instances
    long want = (long) h->sm_count * h->blocks_per_sm;
    long need = ((long) batch + h->wpb - 1) / h->wpb;
    h->grid = (int) (want < need ? want : need);
    h->n_slots = h->grid * h->wpb;
    // the general variants run 4 warps per block on the same scratch slots: never more blocks than n_slots / 4
    h->grid_general = h->n_slots / 4 > 0 ? h->n_slots / 4 : 1;
    if (h->n_slots < 4) h->n_slots = 4;
    {
        const long want_fb = (long) h->sm_count * h->fb_blocks_per_sm, need_fb = ((long) batch + 3) / 4;
        const long need_fbw = ((long) batch + h->fb_wpb - 1) / h->fb_wpb;
        h->grid_fb = (int) (want_fb < need_fbw ? want_fb : need_fbw);
        if (h->grid_fb * h->fb_wpb > h->n_slots) h->n_slots = h->grid_fb * h->fb_wpb;
        int bg = 0;
        CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bg, h->kernel_fb_g4, 128, h->smem_general));
        const long want_g = (long) h->sm_count * (bg > 0 ? bg : 1);
        h->grid_fb_g4 = (int) (want_g < need_fb ? want_g : need_fb);
        if (h->grid_fb_g4 * 4 > h->n_slots) h->n_slots = h->grid_fb_g4 * 4;
        const long want_p = (long) h->sm_count * h->prep_minb;
        h->grid_prep_u = (int) (want_p < need_fb ? want_p : need_fb);
        if (h->grid_prep_u * 4 > h->n_slots) h->n_slots = h->grid_prep_u * 4;
    }
    const long stride = cf_scratch_layout(N).total;
    const size_t B = batch;
    CKH(cudaMalloc(&h->d_x0, B * CF_NX * 8));
    CKH(cudaMalloc(&h->d_yref, B * N * CF_NY * 8));
    CKH(cudaMalloc(&h->d_yref_e, B * CF_NX * 8));
    CKH(cudaMalloc(&h->d_x, B * (N + 1) * CF_NX * 8));
    CKH(cudaMalloc(&h->d_u, B * N * CF_NU * 8));
    CKH(cudaMalloc(&h->d_res, B * 4 * 8));
    CKH(cudaMalloc(&h->d_stage, B * CF_NX * 8));
    CKH(cudaMalloc(&h->d_status, B * 4));
    CKH(cudaMalloc(&h->d_qp_iter, B * 4));
    CKH(cudaMalloc(&h->d_qp_status, B * 4));
    CKH(cudaMalloc(&h->d_flags, B * 4));
    CKH(cudaMalloc(&h->d_counter, 4));
    CKH(cudaMalloc(&h->d_dts, (size_t) N * 8));
    h->h_dts = (double *) malloc((size_t) N * 8);
    for (int i = 0; i < N; i++) h->h_dts[i] = Ts;
    CKH(cudaMemcpy(h->d_dts, h->h_dts, (size_t) N * 8, cudaMemcpyHostToDevice));
    CKH(cudaMalloc(&h->d_policy, B * 4));
    CKH(cudaMalloc(&h->d_titer, B * 4));
    CKH(cudaMalloc(&h->d_motors, B * CF_NU * 4));
    CKH(cudaMalloc(&h->d_setpoint, B * 3 * 8));
    CKH(cudaMalloc(&h->d_euler, B * 3 * 8));
    CKH(cudaMalloc(&h->d_twist, B * 4 * 8));
    CKH(cudaMemsetAsync(h->d_policy, 0, B * 4, h->stream));      // Regulation at (0, 0, 0) until told otherwise
    CKH(cudaMemsetAsync(h->d_titer, 0, B * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_motors, 0, B * CF_NU * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_setpoint, 0, B * 3 * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_euler, 0, B * 3 * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_twist, 0, B * 4 * 8, h->stream));
    // uss as the node computes it: float arithmetic, g0 = 9.80665 (acados_mpc.cpp:107,189,253)
    // `uss = sqrt((mq*g0)/(4*Ct))` with float mq, Ct, uss and g0 a double macro: the product and the quotient are double,
    // 4*Ct is float, only the result is rounded to float = 15.777770042419434
    h->uss = (double) (float) sqrt(((double) 0.033f * 9.80665) / (double) (4.0f * 3.25e-4f));
    CKH(cudaMalloc(&h->d_scratch, (size_t) h->n_slots * stride * 8));
    CKH(cudaMemsetAsync(h->d_scratch, 0, (size_t) h->n_slots * stride * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_x0, 0, B * CF_NX * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_yref, 0, B * N * CF_NY * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_yref_e, 0, B * CF_NX * 8, h->stream));
    CKH(cudaMemsetAsync(h->d_status, 0, B * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_qp_iter, 0, B * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_qp_status, 0, B * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_flags, 0, B * 4, h->stream));
    CKH(cudaMemsetAsync(h->d_res, 0, B * 4 * 8, h->stream));
    {
        const long nx_tot = (long) B * (N + 1) * CF_NX, nu_tot = (long) B * N * CF_NU;
        const long n = nx_tot > nu_tot ? nx_tot : nu_tot;
        cf_init_iterate_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, h->stream>>>(h->d_x, h->d_u, nx_tot, nu_tot);
        CKH(cudaGetLastError());
        h->launches++;
    }
    CfBatchView &bv = h->bv;
    bv.B = batch; bv.first = 0; bv.ready = nullptr; bv.x0 = h->d_x0; bv.yref = h->d_yref; bv.yref_e = h->d_yref_e; bv.x = h->d_x; bv.u = h->d_u;
    bv.status = h->d_status; bv.qp_iter = h->d_qp_iter; bv.qp_status = h->d_qp_status; bv.flags = h->d_flags;
    bv.res = h->d_res; bv.scratch = h->d_scratch; bv.scratch_stride = stride; bv.counter = h->d_counter;
    bv.W_b = bv.WN_b = bv.lbu_b = bv.ubu_b = bv.lbu0_b = bv.ubu0_b = nullptr;
    bv.prof = nullptr;
    bv.dts = h->d_dts; bv.prep = nullptr; bv.prep_stride = cf_prep_stride(N); bv.bnd_stage = nullptr; bv.W_stage = nullptr;
    bv.mult = nullptr; bv.mult_stride = cf_mult_stride(N); bv.W_dense = nullptr;
    CKH(cudaStreamSynchronize(h->stream));
#undef CKH
    *out = h;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_set_stream(cfnmpc_batch *h, void *cuda_stream)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t) cuda_stream : h->own_stream;
    return CFNMPC_OK;
}

struct FieldRef
{
    void *dev;
    size_t bytes;
};
static bool batch_field(cfnmpc_batch *h, const char *f, FieldRef &r)
{
    const size_t B = h->B, N = h->N;
    if (!strcmp(f, "x0")) r = {h->d_x0, B * CF_NX * 8};
    else if (!strcmp(f, "yref")) r = {h->d_yref, B * N * CF_NY * 8};
    else if (!strcmp(f, "yref_e")) r = {h->d_yref_e, B * CF_NX * 8};
    else if (!strcmp(f, "x") || !strcmp(f, "x_all")) r = {h->d_x, B * (N + 1) * CF_NX * 8};
    else if (!strcmp(f, "u") || !strcmp(f, "u_all")) r = {h->d_u, B * N * CF_NU * 8};
    else if (!strcmp(f, "status")) r = {h->d_status, B * 4};
    else if (!strcmp(f, "qp_iter")) r = {h->d_qp_iter, B * 4};
    else if (!strcmp(f, "qp_status")) r = {h->d_qp_status, B * 4};
    else if (!strcmp(f, "flags")) r = {h->d_flags, B * 4};
    else if (!strcmp(f, "res")) r = {h->d_res, B * 4 * 8};
    else if (!strcmp(f, "policy")) r = {h->d_policy, B * 4};
    else if (!strcmp(f, "traj_iter")) r = {h->d_titer, B * 4};
    else if (!strcmp(f, "setpoint")) r = {h->d_setpoint, B * 3 * 8};
    else if (!strcmp(f, "motors")) r = {h->d_motors, B * CF_NU * 4};
    else if (!strcmp(f, "euler")) r = {h->d_euler, B * 3 * 8};
    else if (!strcmp(f, "twist")) r = {h->d_twist, B * 4 * 8};
    else return false;
    return true;
}

static int set_cond_N(cfnmpc_batch *h, int N2);
extern "C" int cfnmpc_batch_set(cfnmpc_batch *h, const char *field, const void *src, int src_on_device)
{
    if (!h || !field || !src) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set: null argument");
    CK(cudaSetDevice(h->device));
    // solver-wide parameters travel as kernel arguments
    double *pdst = nullptr;
    int pn = 0;
    if (!strcmp(field, "W")) { pdst = h->P.Wdiag; pn = CF_NY; }
    else if (!strcmp(field, "W_e")) { pdst = h->P.WNdiag; pn = CF_NX; }
    else if (!strcmp(field, "lbu")) { pdst = h->P.lbu; pn = CF_NU; }
    else if (!strcmp(field, "ubu")) { pdst = h->P.ubu; pn = CF_NU; }
    else if (!strcmp(field, "lbu0")) { pdst = h->P.lbu0; pn = CF_NU; }
    else if (!strcmp(field, "ubu0")) { pdst = h->P.ubu0; pn = CF_NU; }
    if (pdst) {
        if (src_on_device) CK(cudaMemcpy(pdst, src, pn * 8, cudaMemcpyDeviceToHost));
        else memcpy(pdst, src, pn * 8);
        // "lbu"/"ubu" address every stage 0..N-1; "lbu0"/"ubu0" afterwards single out stage 0
        if (pdst == h->P.lbu) memcpy(h->P.lbu0, h->P.lbu, sizeof h->P.lbu);
        if (pdst == h->P.ubu) memcpy(h->P.ubu0, h->P.ubu, sizeof h->P.ubu);
        return CFNMPC_OK;
    }
    // per-instance parameter arrays: allocated on first use, then passed to the kernel instead of the solver-wide value
    {
        double **slot = nullptr;
        const double **view = nullptr;
        int w = 0;
        if (!strcmp(field, "W_batch")) { slot = &h->d_Wb; view = &h->bv.W_b; w = CF_NY; }
        else if (!strcmp(field, "W_e_batch")) { slot = &h->d_WNb; view = &h->bv.WN_b; w = CF_NX; }
        else if (!strcmp(field, "lbu_batch")) { slot = &h->d_lbub; view = &h->bv.lbu_b; w = CF_NU; }
        else if (!strcmp(field, "ubu_batch")) { slot = &h->d_ubub; view = &h->bv.ubu_b; w = CF_NU; }
        else if (!strcmp(field, "lbu0_batch")) { slot = &h->d_lbu0b; view = &h->bv.lbu0_b; w = CF_NU; }
        else if (!strcmp(field, "ubu0_batch")) { slot = &h->d_ubu0b; view = &h->bv.ubu0_b; w = CF_NU; }
        if (slot) {
            const size_t bytes = (size_t) h->B * w * 8;
            if (!*slot) CK(cudaMalloc(slot, bytes));
            CK(cudaMemcpyAsync(*slot, src, bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
            *view = *slot;
            return CFNMPC_OK;
        }
    }
    FieldRef r;
    if (!strcmp(field, "bounds_stage")) {
        // [N][8] = lbu(4) | ubu(4) of every stage: what a sequence of per-stage ocp_nlp_constraints_model_set calls builds
        const size_t bytes = (size_t) h->N * 8 * 8;
        if (!h->d_bst) CK(cudaMalloc(&h->d_bst, bytes));
        CK(cudaMemcpyAsync(h->d_bst, src, bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        h->bv.bnd_stage = h->d_bst;
        return CFNMPC_OK;
    }
    if (!strcmp(field, "W_stage")) {
        // [N+1][17]: diagonal of W_k per stage (cost order y = [x;u]), row N = diagonal of W_e -- what a sequence of
        // per-stage ocp_nlp_cost_model_set(.., k, "W", ..) calls builds (ocp_nlp_cost_ls.c:301-331)
        const size_t bytes = (size_t) (h->N + 1) * CF_NY * 8;
        if (!h->d_wst) CK(cudaMalloc(&h->d_wst, bytes));
        CK(cudaMemcpyAsync(h->d_wst, src, bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        h->bv.W_stage = h->d_wst;
        h->wst = true;
        h->vdt = true;
        h->prepared = false;
        return CFNMPC_OK;
    }
    if (!strcmp(field, "W_dense_table")) {
        // [N+1][17][17] row-major, cost order y = [x;u]: the full weight matrix of every stage (row N: W_e in its leading
        // 13 x 13 block) -- what per-stage ocp_nlp_cost_model_set(.., k, "W", ..) calls with non-diagonal matrices build
        // (ocp_nlp_cost_ls.c:301-331).  The reference's Hessian is scaling (Cyt W_chol)(Cyt W_chol)' with W_chol from
        // blasfeo_dpotrf_l (:743-772), its gradient uses W itself (:883-912): both are handed to the kernel in [u;x] order.
        if (h->bv.mult) return fail(CFNMPC_EINVAL, "W_dense_table: not available together with the option multipliers");
        if (h->cond_N && !h->dense_w) return fail(CFNMPC_EINVAL, "W_dense_table: not available together with qp_cond_N < N");
        const int N = h->N;
        std::vector<double> W((size_t) (N + 1) * 289), T((size_t) (N + 1) * 578, 0.0);
        if (src_on_device) CK(cudaMemcpy(W.data(), src, W.size() * 8, cudaMemcpyDeviceToHost));
        else memcpy(W.data(), src, W.size() * 8);
        auto yidx = [](int r) { return r < CF_NU ? CF_NX + r : r - CF_NU; };   // cost index of stage variable r
        for (int k = 0; k <= N; k++) {
            const int n = k < N ? CF_NY : CF_NX;
            const double *Wk = &W[(size_t) k * 289];
            double L[17][17] = {};
            for (int j = 0; j < n; j++) {
                double d = Wk[j * 17 + j];
                for (int c = 0; c < j; c++) d -= L[j][c] * L[j][c];
                if (!(d > 0.0)) return fail(CFNMPC_EINVAL, "W_dense_table: a weight matrix is not symmetric positive definite");
                d = sqrt(d);
                L[j][j] = d;
                for (int i = j + 1; i < n; i++) {
                    if (Wk[i * 17 + j] != Wk[j * 17 + i]) return fail(CFNMPC_EINVAL, "W_dense_table: a weight matrix is not symmetric");
                    double v = Wk[i * 17 + j];
                    for (int c = 0; c < j; c++) v -= L[i][c] * L[j][c];
                    L[i][j] = v / d;
                }
            }
            double *Hd = &T[(size_t) k * 578], *Wp = Hd + 289;
            for (int r = 0; r < 17; r++)
                for (int c = 0; c < 17; c++) {
                    const int i = yidx(r), j = yidx(c);
                    if (i >= n || j >= n) continue;
                    double v = 0.0;
                    for (int q = 0; q <= (i < j ? i : j); q++) v += L[i][q] * L[j][q];
                    Hd[r * 17 + c] = v;
                    Wp[r * 17 + c] = Wk[i * 17 + j];
                }
        }
        if (!h->d_wdense) CK(cudaMalloc(&h->d_wdense, T.size() * 8));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(h->d_wdense, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
        h->bv.W_dense = h->d_wdense;
        h->dense_w = true;
        h->prepared = false;
        return set_cond_N(h, 0);   // selects the block-size-1 condensed feedback kernel
    }
    if (!strcmp(field, "time_steps")) {
        // crazyflie_acados_update_time_steps (c_templates_tera/acados_solver.in.c:133-153): interval lengths = cost scalings
        std::vector<double> dt(h->N);
        if (src_on_device) CK(cudaMemcpy(dt.data(), src, (size_t) h->N * 8, cudaMemcpyDeviceToHost));
        else memcpy(dt.data(), src, (size_t) h->N * 8);
        bool uniform = true;
        for (int i = 0; i < h->N; i++) {
            if (!(dt[i] > 0)) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set: time steps must be positive");
            uniform = uniform && dt[i] == dt[0];
        }
        CK(cudaStreamSynchronize(h->stream));   // a solve in flight still reads the old grid
        memcpy(h->h_dts, dt.data(), (size_t) h->N * 8);
        CK(cudaMemcpy(h->d_dts, h->h_dts, (size_t) h->N * 8, cudaMemcpyHostToDevice));
        h->P.Ts = dt[0];
        h->vdt_grid = !uniform;
        h->vdt = h->vdt_grid || h->wst || h->itref;
        h->prepared = false;
        return CFNMPC_OK;
    }
    if (!strcmp(field, "uss")) {
        if (src_on_device) CK(cudaMemcpy(&h->uss, src, 8, cudaMemcpyDeviceToHost));
        else memcpy(&h->uss, src, 8);
        return CFNMPC_OK;
    }
    if (!strcmp(field, "x0") || !strcmp(field, "yref") || !strcmp(field, "yref_e") || !strcmp(field, "x") || !strcmp(field, "u") ||
        !strcmp(field, "policy") || !strcmp(field, "traj_iter") || !strcmp(field, "setpoint")) {
        batch_field(h, field, r);
        CK(cudaMemcpyAsync(r.dev, src, r.bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        // a preparation phase belongs to the iterate it linearised: a new iterate invalidates it (the reference would
        // silently combine the two, ocp_nlp_sqp_rti.c:545-683; here cfnmpc_batch_feedback then returns CFNMPC_ESTATE)
        if (!strcmp(field, "x") || !strcmp(field, "u")) h->prepared = false;
        return CFNMPC_OK;
    }
    return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_set: unknown field '") + field + "'");
}

// Select the launch shape of the condensed feedback kernel for block size `bs`: <warps per block, blocks per SM>
template <int BS>
static void (*pc_kernel_for(int wpb, int minb))(const CfParams, const CfBatchView, const CfPcBlocks)
{
    switch (wpb * 100 + minb) {
    case 402: return cf_pcond_kernel<4, 2, BS>;
    case 205: return cf_pcond_kernel<2, 5, BS>;
    case 303: return cf_pcond_kernel<3, 3, BS>;
    case 403: return cf_pcond_kernel<4, 3, BS>;
    case 305: return cf_pcond_kernel<3, 5, BS>;
    case 404: return cf_pcond_kernel<4, 4, BS>;
    default: return nullptr;
    }
}

// Partial condensing to N2 stages (the reference's qp_cond_N, ocp_qp_partial_condensing.c:235-258): 0 or >= N switches it
// off.  Block sizes up to 3 stages are implemented (one row of the condensed stage per lane: 4*3 + 13 + 1 = 26 <= 32).
static int set_cond_N(cfnmpc_batch *h, int N2)
{
    // full weight matrices need the dense stage Hessian of the condensed program: block size 1 stands in for "off"
    if (N2 <= 0 || N2 >= h->N) {
        if (!h->dense_w) { h->cond_N = 0; return CFNMPC_OK; }
        N2 = h->N;
    } else if (h->dense_w) return fail(CFNMPC_EINVAL, "qp_cond_N: partial condensing with full weight matrices (W_dense_table) is not implemented");
    if (h->bv.mult) return fail(CFNMPC_EINVAL, "qp_cond_N / W_dense_table: the condensed feedback program does not produce the multiplier output (option multipliers)");
    const CfPcBlocks b = cf_pc_blocks(h->N, N2);
    const int bs = b.n_big ? b.bs0 + 1 : b.bs0;
    if (bs > 3) return fail(CFNMPC_EINVAL, "qp_cond_N: blocks of more than 3 stages are not implemented (need qp_cond_N >= ceil(N / 3))");
    CK(cudaSetDevice(h->device));
    if (const char *e = getenv("CFNMPC_PC_WARPS_PER_BLOCK")) h->pc_wpb = atoi(e);
    if (const char *e = getenv("CFNMPC_PC_MIN_BLOCKS")) h->pc_minb = atoi(e);
    auto pick = [&](int w, int m) { return bs == 3 ? pc_kernel_for<3>(w, m) : (bs == 2 ? pc_kernel_for<2>(w, m) : pc_kernel_for<1>(w, m)); };
    auto k = pick(h->pc_wpb, h->pc_minb);
    if (!k) { h->pc_wpb = 4; h->pc_minb = 2; k = pick(4, 2); }
    const int smd = bs == 3 ? (int) CfPcWarpT<3>::SM_DOUBLES : (bs == 2 ? (int) CfPcWarpT<2>::SM_DOUBLES : (int) CfPcWarpT<1>::SM_DOUBLES);
    const size_t smem = (size_t) h->pc_wpb * smd * sizeof(double);
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k, h->pc_wpb * 32, smem));
    if (bps < 1) return fail(CFNMPC_ECUDA, "the condensed feedback kernel does not fit on an SM");
    const long want = (long) h->sm_count * bps, need = ((long) h->B + h->pc_wpb - 1) / h->pc_wpb;
    const int grid = (int) (want < need ? want : need);
    // scratch: every resident warp needs a slot of (N2 + 1) condensed stage blocks
    const long stride_pc = bs == 3 ? cf_pc_scratch_doubles<3>(N2) : (bs == 2 ? cf_pc_scratch_doubles<2>(N2) : cf_pc_scratch_doubles<1>(N2));
    const long stride = stride_pc > h->bv.scratch_stride ? stride_pc : h->bv.scratch_stride;
    const int slots = grid * h->pc_wpb > h->n_slots ? grid * h->pc_wpb : h->n_slots;
    if (stride != h->bv.scratch_stride || slots != h->n_slots) {
        CK(cudaStreamSynchronize(h->stream));
        double *ns = nullptr;
        CK(cudaMalloc(&ns, (size_t) slots * stride * 8));
        CK(cudaMemsetAsync(ns, 0, (size_t) slots * stride * 8, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        cudaFree(h->d_scratch);
        h->d_scratch = ns; h->bv.scratch = ns; h->bv.scratch_stride = stride; h->n_slots = slots;
    }
    h->kernel_pc = k; h->smem_pc = smem; h->grid_pc = grid; h->pc_regs = fa.numRegs; h->pc_blocks_per_sm = bps;
    h->pcb = b; h->pc_bs = bs; h->cond_N = N2;
    h->prepared = false;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_set_option(cfnmpc_batch *h, const char *option, int value)
{
    if (!h || !option) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set_option: null argument");
    if (!strcmp(option, "qp_cond_N")) return set_cond_N(h, value);
    if (!strcmp(option, "lin_res_check")) {
        h->P.lin_res_check = value <= 0 ? 0 : (value == 4 ? 4 : (value >= 2 ? 2 : 1));
        // the linear-residual diagnostics and the iterative refinement are compiled into the general kernel variants only
        h->itref = h->P.lin_res_check >= 1;
        h->vdt = h->vdt_grid || h->wst || h->itref;
        h->prepared = false;
    }
    else if (!strcmp(option, "two_kernels")) h->two_kernels = value != 0;
    else if (!strcmp(option, "max_ipm_iter")) h->P.max_ipm_iter = (value > 0 && value < CF_ITER_MAX) ? value : CF_ITER_MAX;
    else if (!strcmp(option, "multipliers")) {
        // keep pi / lam / t of every solved instance (ocp_nlp_out_get "pi" / "lam" / "t"): 8 (29 N + 182) bytes per instance
        if (value && h->cond_N) return fail(CFNMPC_EINVAL, "option multipliers: not available with partial condensing (qp_cond_N < N) or full weight matrices");
        CK(cudaSetDevice(h->device));
        if (value && !h->d_mult) {
            const size_t bytes = (size_t) h->B * h->bv.mult_stride * 8;
            CK(cudaMalloc(&h->d_mult, bytes));
            CK(cudaMemsetAsync(h->d_mult, 0, bytes, h->stream));
        }
        h->bv.mult = value ? h->d_mult : nullptr;
    }
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_set_option: unknown option '") + option + "'");
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_clear(cfnmpc_batch *h, const char *field)
{
    if (!h || !field) return fail(CFNMPC_EINVAL, "cfnmpc_batch_clear: null argument");
    if (!strcmp(field, "W_batch")) h->bv.W_b = nullptr;
    else if (!strcmp(field, "W_e_batch")) h->bv.WN_b = nullptr;
    else if (!strcmp(field, "lbu_batch")) h->bv.lbu_b = nullptr;
    else if (!strcmp(field, "ubu_batch")) h->bv.ubu_b = nullptr;
    else if (!strcmp(field, "lbu0_batch")) h->bv.lbu0_b = nullptr;
    else if (!strcmp(field, "ubu0_batch")) h->bv.ubu0_b = nullptr;
    else if (!strcmp(field, "bounds_stage")) h->bv.bnd_stage = nullptr;
    else if (!strcmp(field, "W_stage")) { h->bv.W_stage = nullptr; h->wst = false; h->vdt = h->vdt_grid || h->itref; h->prepared = false; }
    else if (!strcmp(field, "W_dense_table")) { h->bv.W_dense = nullptr; h->dense_w = false; h->cond_N = 0; h->prepared = false; }
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_clear: '") + field + "' is not a per-instance parameter array");
    return CFNMPC_OK;
}

static int ensure_prep_store(cfnmpc_batch *h);
extern "C" int cfnmpc_batch_solve(cfnmpc_batch *h, int n_rti)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    if (n_rti < 1) return fail(CFNMPC_EINVAL, "cfnmpc_batch_solve: n_rti must be >= 1");
    CK(cudaSetDevice(h->device));
    if (h->cond_N) { if (int rc = ensure_prep_store(h)) return rc; }   // the condensed feedback needs the prepared linearisations
    else if (h->two_kernels && ensure_prep_store(h) != CFNMPC_OK) h->two_kernels = false;   // no room: the fused kernel
    h->mid_valid = false;
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < n_rti; r++) {
        CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
        if (h->cond_N) {
            // preparation (uniform or per-interval grid), then the feedback on the partially condensed QP
            if (h->vdt) h->kernel_prep<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, h->bv);
            else h->kernel_prep_u<<<h->grid_prep_u, 128, h->smem_general, h->stream>>>(h->P, h->bv);
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
            if (r == n_rti - 1) { CK(cudaEventRecord(h->ev_mid, h->stream)); h->mid_valid = n_rti == 1; }
            h->kernel_pc<<<h->grid_pc, h->pc_wpb * 32, h->smem_pc, h->stream>>>(h->P, h->bv, h->pcb);
            h->launches++;
        } else if (h->vdt && !h->two_kernels) h->kernel_vdt<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, h->bv);
        else if (h->vdt) {
            // the general path (per-interval steps, per-stage weights, linear-residual diagnostics) as two launches as well
            h->kernel_prep<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, h->bv);
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
            if (r == n_rti - 1) { CK(cudaEventRecord(h->ev_mid, h->stream)); h->mid_valid = n_rti == 1; }
            h->kernel_fb_g4<<<h->grid_fb_g4, 128, h->smem_general, h->stream>>>(h->P, h->bv);
            h->launches++;
        } else if (h->two_kernels) {
            h->kernel_prep_u<<<h->grid_prep_u, 128, h->smem_general, h->stream>>>(h->P, h->bv);
            CK(cudaGetLastError());
            CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
            if (r == n_rti - 1) { CK(cudaEventRecord(h->ev_mid, h->stream)); h->mid_valid = n_rti == 1; }
            h->kernel_fb_u<<<h->grid_fb, h->fb_wpb * 32, h->smem_fb, h->stream>>>(h->P, h->bv);
            h->launches++;
        } else h->kernel<<<h->grid, h->wpb * 32, h->smem, h->stream>>>(h->P, h->bv);
        CK(cudaGetLastError());
        h->launches++;
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->prepared = false;   // the iterate moved: an earlier preparation no longer belongs to it
    return CFNMPC_OK;
}

// Split real-time iteration: rti_phase 1 / 2 of the reference (ocp_nlp_sqp_rti.c:189-198,1213-1237).
static int ensure_prep_store(cfnmpc_batch *h)
{
    if (!h->d_prep) {
        size_t bytes = (size_t) h->B * h->bv.prep_stride * 8;
        if (getenv("CFNMPC_TEST_FAIL_PREP_ALLOC")) bytes = (size_t) 1 << 60;   // tests: a real failed allocation
        cudaError_t e = cudaMalloc(&h->d_prep, bytes);
        if (e != cudaSuccess) {
            h->d_prep = nullptr;
            (void) cudaGetLastError();   // the failed allocation must not surface as the error of the fused-kernel launch that follows
            return fail(CFNMPC_ECUDA, std::string("cannot allocate the prepared linearisations (") + std::to_string(bytes >> 20) +
                                          " MiB): " + cudaGetErrorString(e));
        }
        h->bv.prep = h->d_prep;
    }
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_prepare(cfnmpc_batch *h)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    if (int rc = ensure_prep_store(h)) return rc;
    h->mid_valid = false;
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    h->kernel_prep<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, h->bv);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->prepared = true;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_feedback(cfnmpc_batch *h)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    if (!h->prepared) return fail(CFNMPC_ESTATE, "cfnmpc_batch_feedback: no preparation phase belongs to the current iterate");
    CK(cudaSetDevice(h->device));
    h->mid_valid = false;
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    if (h->cond_N) h->kernel_pc<<<h->grid_pc, h->pc_wpb * 32, h->smem_pc, h->stream>>>(h->P, h->bv, h->pcb);
    else h->kernel_fb<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, h->bv);
    CK(cudaGetLastError());
    h->launches++;
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->prepared = false;
    return CFNMPC_OK;
}

// One control tick fed from HOST buffers: ONE launch of the persistent kernel starts at once; the inputs (x0, yref,
// yref_e) follow in n_chunks contiguous chunks on a second stream, each chunk followed by a 4-byte copy that advances
// the device-side "upload front"; a warp that pulls an instance beyond the front waits for it.  Upload and solve overlap
// without the per-launch tails that chunked launches would add.
#define CF_MAX_CHUNKS 32
extern "C" int cfnmpc_batch_solve_from_host(cfnmpc_batch *h, const double *x0, const double *yref, const double *yref_e, int n_chunks)
{
    if (!h || !x0 || !yref || !yref_e) return fail(CFNMPC_EINVAL, "cfnmpc_batch_solve_from_host: null argument");
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > CF_MAX_CHUNKS) n_chunks = CF_MAX_CHUNKS;
    if (n_chunks > h->B) n_chunks = h->B;
    CK(cudaSetDevice(h->device));
    if (h->cond_N) { if (int rc = ensure_prep_store(h)) return rc; }
    else if (h->two_kernels && ensure_prep_store(h) != CFNMPC_OK) h->two_kernels = false;
    h->mid_valid = false;
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_free, cudaEventDisableTiming));
        CK(cudaMalloc(&h->d_ready, 4));
        CK(cudaHostAlloc(&h->h_ready, CF_MAX_CHUNKS * sizeof(int), cudaHostAllocDefault));
    }
    // the staging counters of the previous call must have been consumed before they are rewritten
    CK(cudaStreamSynchronize(h->copy_stream));
    CK(cudaEventRecord(h->ev0, h->stream));
    CK(cudaMemsetAsync(h->d_ready, 0, 4, h->stream));
    CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    // the copies may only overwrite the inputs once everything enqueued so far on the solve stream is done with them
    CK(cudaEventRecord(h->ev_free, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_free, 0));
    const size_t N = h->N;
    cudaError_t e = cudaSuccess;
    for (int c = 0; c < n_chunks && e == cudaSuccess; c++) {
        const size_t first = (size_t) h->B * c / n_chunks, last = (size_t) h->B * (c + 1) / n_chunks, cnt = last - first;
        e = cudaMemcpyAsync(h->d_x0 + first * CF_NX, x0 + first * CF_NX, cnt * CF_NX * 8, cudaMemcpyHostToDevice, h->copy_stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h->d_yref + first * N * CF_NY, yref + first * N * CF_NY, cnt * N * CF_NY * 8, cudaMemcpyHostToDevice, h->copy_stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h->d_yref_e + first * CF_NX, yref_e + first * CF_NX, cnt * CF_NX * 8, cudaMemcpyHostToDevice, h->copy_stream);
        h->h_ready[c] = (int) last;
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_ready, h->h_ready + c, 4, cudaMemcpyHostToDevice, h->copy_stream);
    }
    if (e != cudaSuccess)
        return fail(CFNMPC_ECUDA, std::string("cfnmpc_batch_solve_from_host: upload failed: ") + cudaGetErrorString(e));
    // the copies are enqueued BEFORE the launch: a synchronous launch (profilers, CUDA_LAUNCH_BLOCKING) cannot dead-lock
    CfBatchView bv = h->bv;
    bv.ready = h->d_ready;
    if (h->cond_N) {
        bv.prep = h->d_prep;
        if (h->vdt) h->kernel_prep<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, bv);
        else h->kernel_prep_u<<<h->grid_prep_u, 128, h->smem_general, h->stream>>>(h->P, bv);
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
        bv.ready = nullptr;
        h->kernel_pc<<<h->grid_pc, h->pc_wpb * 32, h->smem_pc, h->stream>>>(h->P, bv, h->pcb);
        h->launches++;
    } else if (h->vdt && !h->two_kernels) h->kernel_vdt<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, bv);
    else if (h->vdt) {
        bv.prep = h->d_prep;
        h->kernel_prep<<<h->grid_general, 128, h->smem_general, h->stream>>>(h->P, bv);
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
        bv.ready = nullptr;
        h->kernel_fb_g4<<<h->grid_fb_g4, 128, h->smem_general, h->stream>>>(h->P, bv);
        h->launches++;
    } else if (h->two_kernels) {
        // the preparation follows the upload front; by the time it has finished every input is in place
        bv.prep = h->d_prep;
        h->kernel_prep_u<<<h->grid_prep_u, 128, h->smem_general, h->stream>>>(h->P, bv);
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
        bv.ready = nullptr;
        h->kernel_fb_u<<<h->grid_fb, h->fb_wpb * 32, h->smem_fb, h->stream>>>(h->P, bv);
        h->launches++;
    } else h->kernel<<<h->grid, h->wpb * 32, h->smem, h->stream>>>(h->P, bv);
    CK(cudaGetLastError());
    h->launches++;
    h->prepared = false;
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_sync(cfnmpc_batch *h)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_get(cfnmpc_batch *h, const char *field, int stage, void *dst, int dst_on_device)
{
    if (!h || !field || !dst) return fail(CFNMPC_EINVAL, "cfnmpc_batch_get: null argument");
    CK(cudaSetDevice(h->device));
    const cudaMemcpyKind kind = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    const bool is_u = !strcmp(field, "u"), is_x = !strcmp(field, "x");
    const bool is_pi = !strcmp(field, "pi"), is_lam = !strcmp(field, "lam"), is_t = !strcmp(field, "t"), is_l0 = !strcmp(field, "lam_x0");
    const bool all_pi = !strcmp(field, "pi_all"), all_lam = !strcmp(field, "lam_all"), all_t = !strcmp(field, "t_all");
    if (all_pi || all_lam || all_t) {   // every stage at once: [B][N][13] / [B][N][8]
        if (!h->bv.mult) return fail(CFNMPC_ESTATE, "cfnmpc_batch_get: set option multipliers before the solve");
        const size_t w = (size_t) h->N * (all_pi ? CF_NX : 8) * 8;
        const long sec = all_pi ? 0 : (all_lam ? (long) h->N * 13 : (long) h->N * 21);
        CK(cudaMemcpy2DAsync(dst, w, h->d_mult + sec, (size_t) h->bv.mult_stride * 8, w, h->B, kind, h->stream));
    } else if (is_pi || is_lam || is_t || is_l0) {
        // multipliers of the iterate: pi_k [B][13]; lam_k / t_k [B][8] = lower(4) | upper(4) of the input box of stage k;
        // lam_x0 [B][13] signed multiplier of x_0 = x0 (CfBatchView::mult)
        if (!h->bv.mult) return fail(CFNMPC_ESTATE, "cfnmpc_batch_get: set option multipliers before the solve");
        if (stage < 0 || stage >= (is_l0 ? 1 : h->N)) return fail(CFNMPC_EINVAL, "cfnmpc_batch_get: stage out of range");
        const int w = (is_pi || is_l0) ? CF_NX : 8;
        const long sec = is_pi ? 0 : (is_lam ? (long) h->N * 13 : (is_t ? (long) h->N * 21 : (long) h->N * 29));
        const int n = h->B * w;
        cf_gather_stage_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->d_mult + sec, h->d_stage, h->B, (int) h->bv.mult_stride, stage, w);
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaMemcpyAsync(dst, h->d_stage, (size_t) n * 8, kind, h->stream));
    } else if (is_u || is_x) {
        const int w = is_u ? CF_NU : CF_NX, nst = is_u ? h->N : h->N + 1;
        if (stage < 0 || stage >= nst) return fail(CFNMPC_EINVAL, "cfnmpc_batch_get: stage out of range");
        const int n = h->B * w;
        cf_gather_stage_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(is_u ? h->d_u : h->d_x, h->d_stage, h->B, nst * w, stage, w);
        CK(cudaGetLastError());
        h->launches++;
        CK(cudaMemcpyAsync(dst, h->d_stage, (size_t) n * 8, kind, h->stream));
    } else {
        FieldRef r;
        if (!batch_field(h, field, r))
            return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_get: unknown field '") + field + "'");
        CK(cudaMemcpyAsync(dst, r.dev, r.bytes, kind, h->stream));
    }
    if (!dst_on_device) CK(cudaStreamSynchronize(h->stream));
    return CFNMPC_OK;
}

// ------------------------------------------------------------------ closed-loop driver
extern "C" int cfnmpc_batch_set_trajectory(cfnmpc_batch *h, const double *table, int n_rows, int src_on_device)
{
    if (!h || !table) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set_trajectory: null argument");
    if (n_rows <= h->N) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set_trajectory: the table needs more than N rows");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_traj) { CK(cudaFree(h->d_traj)); h->d_traj = nullptr; }
    const size_t bytes = (size_t) n_rows * CF_NY * 8;
    CK(cudaMalloc(&h->d_traj, bytes));
    CK(cudaMemcpyAsync(h->d_traj, table, bytes, src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    h->n_traj = n_rows;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_update_reference(cfnmpc_batch *h)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    const long n = (long) h->B * (h->N + 1) * CF_NY;
    cf_reference_window_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, h->stream>>>(h->B, h->N, h->d_policy, h->d_titer, h->d_setpoint,
                                                                                h->d_traj, h->n_traj, h->uss, h->d_yref, h->d_yref_e);
    CK(cudaGetLastError());
    cf_policy_advance_kernel<<<(h->B + 255) / 256, 256, 0, h->stream>>>(h->B, h->N, h->d_policy, h->d_titer, h->n_traj);
    CK(cudaGetLastError());
    h->launches += 2;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_commands(cfnmpc_batch *h, int motors_from_u1)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    cf_command_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->N, h->d_x, h->d_u, motors_from_u1, h->d_motors, h->d_euler, h->d_twist);
    CK(cudaGetLastError());
    h->launches++;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_plant_step(cfnmpc_batch *h, double dt, int n_steps, int truncated_motors)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    if (!(dt > 0) || n_steps < 1) return fail(CFNMPC_EINVAL, "cfnmpc_batch_plant_step: need dt > 0 and n_steps >= 1");
    CK(cudaSetDevice(h->device));
    // input of the plant: the first control of the current solution (u_0 of every instance), or the int32 motor command
    if (!truncated_motors) {
        cf_gather_stage_kernel<<<(h->B * CF_NU + 255) / 256, 256, 0, h->stream>>>(h->d_u, h->d_stage, h->B, h->N * CF_NU, 0, CF_NU);
        CK(cudaGetLastError());
        h->launches++;
    }
    cf_predict_kernel<<<(h->B + CF_PRED_THREADS - 1) / CF_PRED_THREADS, CF_PRED_THREADS, 0, h->stream>>>(
        h->d_x0, h->d_stage, truncated_motors ? h->d_motors : nullptr, nullptr, dt, n_steps, h->B, h->d_x0);
    CK(cudaGetLastError());
    h->launches++;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_solve(cfnmpc_batch *h, int n_rti);
extern "C" int cfnmpc_batch_tick(cfnmpc_batch *h, int motors_from_u1)
{
    int rc = cfnmpc_batch_update_reference(h);
    if (rc == CFNMPC_OK) rc = cfnmpc_batch_solve(h, 1);
    if (rc == CFNMPC_OK) rc = cfnmpc_batch_commands(h, motors_from_u1);
    return rc;
}

// ------------------------------------------------------------------ batched state predictor object
struct cfnmpc_sim
{
    int B = 0, device = 0, n_steps = 1, sens_forw = 0;
    bool T_per_instance = false;
    double T_all = 0.015;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    double *d_x = nullptr, *d_u = nullptr, *d_T = nullptr, *d_xn = nullptr, *d_S = nullptr;
    long long launches = 0;
};

extern "C" int cfnmpc_sim_destroy(cfnmpc_sim *s)
{
    if (!s) return CFNMPC_OK;
    cudaSetDevice(s->device);
    void *ptrs[] = {s->d_x, s->d_u, s->d_T, s->d_xn, s->d_S};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_create(int batch, int device, cfnmpc_sim **out)
{
    if (!out) return fail(CFNMPC_EINVAL, "cfnmpc_sim_create: out is NULL");
    *out = nullptr;
    if (batch < 1) return fail(CFNMPC_EINVAL, "cfnmpc_sim_create: need batch >= 1");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(CFNMPC_EINVAL, "cfnmpc_sim_create: no such CUDA device");
    CK(cudaSetDevice(device));
    cfnmpc_sim *s = new cfnmpc_sim();
    s->B = batch; s->device = device;
    const size_t B = batch;
    cudaError_t e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_x, B * CF_NX * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_u, B * CF_NU * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_T, B * 8);
    if (e == cudaSuccess) e = cudaMalloc(&s->d_xn, B * CF_NX * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_x, 0, B * CF_NX * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_u, 0, B * CF_NU * 8);
    if (e == cudaSuccess) e = cudaMemset(s->d_xn, 0, B * CF_NX * 8);
    if (e != cudaSuccess) {
        cfnmpc_sim_destroy(s);
        return fail(CFNMPC_ECUDA, std::string("cfnmpc_sim_create: ") + cudaGetErrorString(e));
    }
    s->stream = s->own_stream;
    *out = s;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_set_stream(cfnmpc_sim *s, void *cuda_stream)
{
    if (!s) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->stream));
    s->stream = cuda_stream ? (cudaStream_t) cuda_stream : s->own_stream;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_opts_set(cfnmpc_sim *s, const char *field, int value)
{
    if (!s || !field) return fail(CFNMPC_EINVAL, "cfnmpc_sim_opts_set: null argument");
    if (!strcmp(field, "num_steps")) {
        if (value < 1) return fail(CFNMPC_EINVAL, "cfnmpc_sim_opts_set: num_steps must be >= 1");
        s->n_steps = value;
    } else if (!strcmp(field, "sens_forw")) s->sens_forw = value != 0;
    else if (!strcmp(field, "num_stages")) {
        if (value != 4) return fail(CFNMPC_EINVAL, "cfnmpc_sim_opts_set: only the 4-stage ERK of the reference configuration is implemented");
    } else return fail(CFNMPC_EINVAL, std::string("cfnmpc_sim_opts_set: unknown option '") + field + "'");
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_set(cfnmpc_sim *s, const char *field, const void *src, int src_on_device)
{
    if (!s || !field || !src) return fail(CFNMPC_EINVAL, "cfnmpc_sim_set: null argument");
    CK(cudaSetDevice(s->device));
    const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t B = s->B;
    if (!strcmp(field, "x")) CK(cudaMemcpyAsync(s->d_x, src, B * CF_NX * 8, kind, s->stream));
    else if (!strcmp(field, "u")) CK(cudaMemcpyAsync(s->d_u, src, B * CF_NU * 8, kind, s->stream));
    else if (!strcmp(field, "T_batch")) { CK(cudaMemcpyAsync(s->d_T, src, B * 8, kind, s->stream)); s->T_per_instance = true; }
    else if (!strcmp(field, "T")) {
        if (src_on_device) CK(cudaMemcpy(&s->T_all, src, 8, cudaMemcpyDeviceToHost));
        else memcpy(&s->T_all, src, 8);
        s->T_per_instance = false;
    } else return fail(CFNMPC_EINVAL, std::string("cfnmpc_sim_set: unknown field '") + field + "'");
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_solve(cfnmpc_sim *s)
{
    if (!s) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(s->device));
    const double *T_b = s->T_per_instance ? s->d_T : nullptr;
    if (s->sens_forw) {
        if (!s->d_S) CK(cudaMalloc(&s->d_S, (size_t) s->B * CF_NX * CF_NV * 8));
        const long threads = (long) s->B * 32;
        cf_predict_sens_kernel<<<(unsigned) ((threads + 127) / 128), 128, 0, s->stream>>>(s->d_x, s->d_u, T_b, s->T_all, s->n_steps, s->B, s->d_xn, s->d_S);
    } else {
        cf_predict_kernel<<<(s->B + CF_PRED_THREADS - 1) / CF_PRED_THREADS, CF_PRED_THREADS, 0, s->stream>>>(s->d_x, s->d_u, nullptr, T_b, s->T_all,
                                                                                                     s->n_steps, s->B, s->d_xn);
    }
    CK(cudaGetLastError());
    s->launches++;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_get(cfnmpc_sim *s, const char *field, void *dst, int dst_on_device)
{
    if (!s || !field || !dst) return fail(CFNMPC_EINVAL, "cfnmpc_sim_get: null argument");
    CK(cudaSetDevice(s->device));
    const cudaMemcpyKind kind = dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (!strcmp(field, "xn")) CK(cudaMemcpyAsync(dst, s->d_xn, (size_t) s->B * CF_NX * 8, kind, s->stream));
    else if (!strcmp(field, "S_forw")) {
        if (!s->sens_forw || !s->d_S) return fail(CFNMPC_ESTATE, "cfnmpc_sim_get: S_forw needs the option sens_forw = 1 and a solve");
        CK(cudaMemcpyAsync(dst, s->d_S, (size_t) s->B * CF_NX * CF_NV * 8, kind, s->stream));
    } else return fail(CFNMPC_EINVAL, std::string("cfnmpc_sim_get: unknown field '") + field + "'");
    if (!dst_on_device) CK(cudaStreamSynchronize(s->stream));
    return CFNMPC_OK;
}

extern "C" int cfnmpc_sim_launches(cfnmpc_sim *s, long long *n)
{
    if (!s || !n) return fail(CFNMPC_EINVAL, "null argument");
    *n = s->launches;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_device_ptr(cfnmpc_batch *h, const char *field, void **ptr)
{
    if (!h || !field || !ptr) return fail(CFNMPC_EINVAL, "cfnmpc_batch_device_ptr: null argument");
    FieldRef r;
    if (!batch_field(h, field, r)) return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_device_ptr: unknown field '") + field + "'");
    *ptr = r.dev;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_info(cfnmpc_batch *h, const char *what, long long *value)
{
    if (!h || !what || !value) return fail(CFNMPC_EINVAL, "cfnmpc_batch_info: null argument");
    if (!strcmp(what, "batch")) *value = h->B;
    else if (!strcmp(what, "N")) *value = h->N;
    else if (!strcmp(what, "n_slots")) *value = h->n_slots;
    else if (!strcmp(what, "sm_count")) *value = h->sm_count;
    else if (!strcmp(what, "warps_per_block")) *value = h->wpb;
    else if (!strcmp(what, "blocks_per_sm")) *value = h->blocks_per_sm;
    else if (!strcmp(what, "grid")) *value = h->grid;
    else if (!strcmp(what, "regs_per_thread")) *value = h->regs;
    else if (!strcmp(what, "smem_per_block")) *value = (long long) h->smem;
    else if (!strcmp(what, "scratch_bytes")) *value = (long long) h->n_slots * h->bv.scratch_stride * 8;
    else if (!strcmp(what, "launches")) *value = h->launches;
    else if (!strcmp(what, "two_kernels")) *value = h->cond_N ? 1 : h->two_kernels;
    else if (!strcmp(what, "qp_cond_N")) *value = h->cond_N ? h->cond_N : h->N;
    else if (!strcmp(what, "pcond_block_size")) *value = h->cond_N ? h->pc_bs : 1;
    else if (!strcmp(what, "pcond_regs_per_thread")) *value = h->pc_regs;
    else if (!strcmp(what, "pcond_blocks_per_sm")) *value = h->pc_blocks_per_sm;
    else if (!strcmp(what, "pcond_warps_per_block")) *value = h->pc_wpb;
    else if (!strcmp(what, "pcond_grid")) *value = h->grid_pc;
    else if (!strcmp(what, "feedback_regs_per_thread")) *value = h->fb_regs;
    else if (!strcmp(what, "feedback_blocks_per_sm")) *value = h->fb_blocks_per_sm;
    else if (!strcmp(what, "feedback_grid")) *value = h->grid_fb;
    else if (!strcmp(what, "feedback_warps_per_block")) *value = h->fb_wpb;
    else if (!strcmp(what, "preparation_grid")) *value = h->grid_prep_u;
    else if (!strcmp(what, "prepared_bytes")) *value = h->d_prep ? (long long) h->B * h->bv.prep_stride * 8 : 0;
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_info: unknown property '") + what + "'");
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_last_solve_ms(cfnmpc_batch *h, double *ms)
{
    if (!h || !ms) return fail(CFNMPC_EINVAL, "null argument");
    if (!h->timed) return fail(CFNMPC_ESTATE, "cfnmpc_batch_last_solve_ms: no solve has been enqueued");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = f;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_last_phase_ms(cfnmpc_batch *h, double *ms2)
{
    if (!h || !ms2) return fail(CFNMPC_EINVAL, "null argument");
    if (!h->timed) return fail(CFNMPC_ESTATE, "cfnmpc_batch_last_phase_ms: no solve has been enqueued");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev1));
    float f = 0;
    if (h->mid_valid) {
        CK(cudaEventElapsedTime(&f, h->ev0, h->ev_mid));
        ms2[0] = f;
        CK(cudaEventElapsedTime(&f, h->ev_mid, h->ev1));
        ms2[1] = f;
    } else {
        CK(cudaEventElapsedTime(&f, h->ev0, h->ev1));
        ms2[0] = 0.0;
        ms2[1] = f;
    }
    return CFNMPC_OK;
}

extern "C" int cfnmpc_debug_scratch(cfnmpc_batch *h, double *dst, size_t max_doubles, size_t *n_doubles, long long *offsets12)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CfScratchLayout s = cf_scratch_layout(h->N);
    if (offsets12) {
        long long v[12] = {s.total, CF_SB, B_M, B_LU, B_PX, R_UX, R_PI, R_RQ, R_B, R_RESG, R_DUX, R_D};
        memcpy(offsets12, v, sizeof v);
    }
    if (n_doubles) *n_doubles = (size_t) s.total;
    if (dst) {
        if (h->B != 1) return fail(CFNMPC_ESTATE, "cfnmpc_debug_scratch: only meaningful for batch == 1");
        if (max_doubles < (size_t) s.total) return fail(CFNMPC_EINVAL, "cfnmpc_debug_scratch: buffer too small");
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        // any warp of the (single) block may have pulled instance 0: find the slot that was written
        int slot = 0;
        for (int w = 0; w < h->n_slots; w++) {
            double probe[CF_CMSZ];
            CK(cudaMemcpy(probe, h->d_scratch + (size_t) w * s.total + B_M, sizeof probe, cudaMemcpyDeviceToHost));
            bool used = false;
            for (int i = 0; i < CF_CMSZ; i++) used |= probe[i] != 0.0;
            if (used) { slot = w; break; }
        }
        CK(cudaMemcpy(dst, h->d_scratch + (size_t) slot * s.total, (size_t) s.total * 8, cudaMemcpyDeviceToHost));
    }
    return CFNMPC_OK;
}

// Profiling aid: enable (cycles_calls != NULL) per-pass counters and read them back: for each of the passes
// {linearisation, residual+factorisation, forward, rhs-backward, mu_aff, primal update} the sum over warps of elapsed SM
// cycles and the number of calls since the counters were enabled.  Costs two clock reads per pass while enabled.
extern "C" int cfnmpc_debug_pass_cycles(cfnmpc_batch *h, unsigned long long *cycles_calls12)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (!h->d_prof) {
        CK(cudaMalloc(&h->d_prof, 2 * CF_PROF_N * 8));
        CK(cudaMemset(h->d_prof, 0, 2 * CF_PROF_N * 8));
        h->bv.prof = h->d_prof;
    }
    if (cycles_calls12) {
        CK(cudaMemcpy(cycles_calls12, h->d_prof, 2 * CF_PROF_N * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemset(h->d_prof, 0, 2 * CF_PROF_N * 8));
    }
    return CFNMPC_OK;
}

extern "C" int cfnmpc_debug_max_ipm_iter(cfnmpc_batch *h, int max_iter)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    h->P.max_ipm_iter = (max_iter > 0 && max_iter < CF_ITER_MAX) ? max_iter : CF_ITER_MAX;
    return CFNMPC_OK;
}

// ------------------------------------------------------------------ measured fp64 peak (roofline denominator)
// Dependent-chain-free DFMA stream: 8 independent accumulators per thread, all SMs filled; FLOP = 2 per FMA.
__global__ void __launch_bounds__(256) cf_fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = a + threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = fma(v[i], b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[i];
    if (s == 12345.678) out[0] = s;   // never true: keeps the loop alive
}

extern "C" int cfnmpc_measure_fp64_peak(int device, double *tflops)
{
    if (!tflops) return fail(CFNMPC_EINVAL, "null argument");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double *d = nullptr;
    CK(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        CK(cudaEventRecord(e0));
        cf_fp64_peak_kernel<<<blocks, 256>>>(d, iters, 1.0000001, 0.9999999);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double tf = 2.0 * 8.0 * (double) iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return CFNMPC_OK;
}
