// Batch C-ABI (include/cfnmpc.h, core subset) for ANY generated OCP description: the same kernel sources compiled with
// -DCF_SPEC_HEADER='"cf_spec_<model>.h"' (tools/gen_spec.py --model <model>) into crazyflie_nmpc_b200/libcfnmpc_<model>.so
// (other nx, nu compile without touching kernels).
//
// One RTI step = two launches: the preparation program of cf_rti_warp.h (ERK4 + forward sensitivities, Gauss-Newton
// gradient; rti_phase 1) and the dense-stage feedback program of cf_pcond_warp.h with block size 1 (x0 elimination,
// Mehrotra IPM with the classical Riccati recursion on the fp64 tensor cores, primal update; rti_phase 2).  The hand-tuned
// uncondensed feedback program, the node-specific kernels and the single-instance acados surfaces exist for the Crazyflie
// sizes only (cfnmpc_api.cu, acados_shim.cpp).
//
// Entry points implemented here: cfnmpc_batch_create / destroy / set / set_option / solve / prepare / feedback / sync / get /
// last_solve_ms / info, cfnmpc_last_error, cfnmpc_version, and cfnmpc_model_dims.  No CPU path.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/cfnmpc.h"
#include "cf_kernels.h"

static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail(CFNMPC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct cfnmpc_batch
{
    int B = 0, N = 0, device = 0;
    CfParams P;
    CfBatchView bv;
    CfPcBlocks pcb;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_x0 = nullptr, *d_yref = nullptr, *d_yref_e = nullptr, *d_x = nullptr, *d_u = nullptr, *d_res = nullptr;
    double *d_scratch = nullptr, *d_prep = nullptr, *d_dts = nullptr, *d_bst = nullptr, *d_wst = nullptr;
    int *d_status = nullptr, *d_qp_iter = nullptr, *d_qp_status = nullptr, *d_flags = nullptr, *d_counter = nullptr;
    int grid = 0, sm_count = 0, regs_prep = 0, regs_fb = 0;
    size_t smem_prep = 0, smem_fb = 0;
    long long launches = 0;
    bool prepared = false, timed = false;
};

extern "C" const char *cfnmpc_last_error(void) { return g_err.c_str(); }
extern "C" const char *cfnmpc_version(void) { return "crazyflie_nmpc_b200 0.2 generic-model build (sm_100a)"; }
// sizes this library was generated for (and the horizon / final time of its OCP description)
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
    void *ptrs[] = {h->d_x0, h->d_yref, h->d_yref_e, h->d_x, h->d_u, h->d_res, h->d_scratch, h->d_prep, h->d_dts, h->d_bst, h->d_wst,
                    h->d_status, h->d_qp_iter, h->d_qp_status, h->d_flags, h->d_counter};
    for (void *p : ptrs) if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
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
    // the generated OCP description: weights, input box (tools/gen_spec.py)
    for (int i = 0; i < CF_NY; i++) h->P.Wdiag[i] = CfSpec::W[i];
    for (int i = 0; i < CF_NX; i++) h->P.WNdiag[i] = CfSpec::W_e[i];
    for (int i = 0; i < CF_NU; i++) { h->P.lbu[i] = h->P.lbu0[i] = CfSpec::lbu[i]; h->P.ubu[i] = h->P.ubu0[i] = CfSpec::ubu[i]; }
    h->P.Ts = Ts; h->P.N = N; h->P.max_ipm_iter = CF_ITER_MAX; h->P.lin_res_check = 0; h->P.pad_ = 0;
#define CKH(call)                                                                                   \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            cfnmpc_batch_destroy(h);                                                                \
            return fail(CFNMPC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
        }                                                                                           \
    } while (0)
    CKH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKH(cudaEventCreate(&h->ev0));
    CKH(cudaEventCreate(&h->ev1));
    cudaDeviceProp prop;
    CKH(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    auto kp = cf_rti_kernel<4, 3, CF_PH_PREPARATION, true>;
    auto kf = cf_pcond_kernel<4, 3, 1>;
    h->smem_prep = (size_t) 4 * CF_SM_DOUBLES * sizeof(double);
    h->smem_fb = (size_t) 4 * CfPcWarpT<1>::SM_DOUBLES * sizeof(double);
    CKH(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_prep));
    CKH(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) h->smem_fb));
    cudaFuncAttributes fa;
    CKH(cudaFuncGetAttributes(&fa, kp));
    h->regs_prep = fa.numRegs;
    CKH(cudaFuncGetAttributes(&fa, kf));
    h->regs_fb = fa.numRegs;
    int bp = 0, bf = 0;
    CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bp, kp, 128, h->smem_prep));
    CKH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, kf, 128, h->smem_fb));
    if (bp < 1 || bf < 1) { cfnmpc_batch_destroy(h); return fail(CFNMPC_ECUDA, "a kernel does not fit on an SM"); }
    const int bps = bp < bf ? bp : bf;
    const long want = (long) h->sm_count * bps, need = ((long) batch + 3) / 4;
    h->grid = (int) (want < need ? want : need);
    const int n_slots = h->grid * 4;
    h->pcb = cf_pc_blocks(N, N);
    const long s0 = cf_scratch_layout(N).total, s1 = cf_pc_scratch_doubles<1>(N);
    const long stride = s0 > s1 ? s0 : s1;
    const size_t B = batch;
    CKH(cudaMalloc(&h->d_x0, B * CF_NX * 8));
    CKH(cudaMalloc(&h->d_yref, B * N * CF_NY * 8));
    CKH(cudaMalloc(&h->d_yref_e, B * CF_NX * 8));
    CKH(cudaMalloc(&h->d_x, B * (N + 1) * CF_NX * 8));
    CKH(cudaMalloc(&h->d_u, B * N * CF_NU * 8));
    CKH(cudaMalloc(&h->d_res, B * 4 * 8));
    CKH(cudaMalloc(&h->d_status, B * 4));
    CKH(cudaMalloc(&h->d_qp_iter, B * 4));
    CKH(cudaMalloc(&h->d_qp_status, B * 4));
    CKH(cudaMalloc(&h->d_flags, B * 4));
    CKH(cudaMalloc(&h->d_counter, 4));
    CKH(cudaMalloc(&h->d_dts, (size_t) N * 8));
    CKH(cudaMalloc(&h->d_scratch, (size_t) n_slots * stride * 8));
    CKH(cudaMalloc(&h->d_prep, B * cf_prep_stride(N) * 8));
    std::vector<double> dts(N, Ts);
    CKH(cudaMemcpy(h->d_dts, dts.data(), (size_t) N * 8, cudaMemcpyHostToDevice));
    CKH(cudaMemset(h->d_scratch, 0, (size_t) n_slots * stride * 8));
    for (void *p : {(void *) h->d_x0, (void *) h->d_yref_e}) CKH(cudaMemset(p, 0, B * CF_NX * 8));
    CKH(cudaMemset(h->d_yref, 0, B * N * CF_NY * 8));
    CKH(cudaMemset(h->d_u, 0, B * N * CF_NU * 8));
    CKH(cudaMemset(h->d_res, 0, B * 4 * 8));
    for (int *p : {h->d_status, h->d_qp_iter, h->d_qp_status, h->d_flags}) CKH(cudaMemset(p, 0, B * 4));
    {   // initial iterate: x_k = x0 of the OCP description, u = 0 (acados_solver.in.c:2323-2352)
        std::vector<double> xi((size_t) B * (N + 1) * CF_NX);
        for (size_t i = 0; i < xi.size(); i++) xi[i] = CfSpec::x0[i % CF_NX];
        CKH(cudaMemcpy(h->d_x, xi.data(), xi.size() * 8, cudaMemcpyHostToDevice));
    }
    CfBatchView &bv = h->bv;
    memset(&bv, 0, sizeof bv);
    bv.B = batch; bv.x0 = h->d_x0; bv.yref = h->d_yref; bv.yref_e = h->d_yref_e; bv.x = h->d_x; bv.u = h->d_u;
    bv.status = h->d_status; bv.qp_iter = h->d_qp_iter; bv.qp_status = h->d_qp_status; bv.flags = h->d_flags; bv.res = h->d_res;
    bv.scratch = h->d_scratch; bv.scratch_stride = stride; bv.counter = h->d_counter;
    bv.dts = h->d_dts; bv.prep = h->d_prep; bv.prep_stride = cf_prep_stride(N); bv.mult_stride = 0;
#undef CKH
    *out = h;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_set(cfnmpc_batch *h, const char *field, const void *src, int src_on_device)
{
    if (!h || !field || !src) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set: null argument");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B, N = h->N;
    const cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
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
        if (pdst == h->P.lbu) memcpy(h->P.lbu0, h->P.lbu, sizeof h->P.lbu);
        if (pdst == h->P.ubu) memcpy(h->P.ubu0, h->P.ubu, sizeof h->P.ubu);
        h->prepared = false;
        return CFNMPC_OK;
    }
    if (!strcmp(field, "time_steps")) {
        std::vector<double> dt(N);
        CK(cudaMemcpy(dt.data(), src, N * 8, src_on_device ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
        for (size_t i = 0; i < N; i++) if (!(dt[i] > 0)) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set: time steps must be positive");
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(h->d_dts, dt.data(), N * 8, cudaMemcpyHostToDevice));
        h->P.Ts = dt[0];
        h->prepared = false;
        return CFNMPC_OK;
    }
    if (!strcmp(field, "bounds_stage") || !strcmp(field, "W_stage")) {
        const bool bs = field[0] == 'b';
        const size_t bytes = bs ? N * 2 * CF_NU * 8 : (N + 1) * CF_NY * 8;
        double **slot = bs ? &h->d_bst : &h->d_wst;
        if (!*slot) CK(cudaMalloc(slot, bytes));
        CK(cudaMemcpyAsync(*slot, src, bytes, kind, h->stream));
        if (bs) h->bv.bnd_stage = h->d_bst; else h->bv.W_stage = h->d_wst;
        h->prepared = false;
        return CFNMPC_OK;
    }
    double *dst = nullptr;
    size_t bytes = 0;
    if (!strcmp(field, "x0")) { dst = h->d_x0; bytes = B * CF_NX * 8; }
    else if (!strcmp(field, "yref")) { dst = h->d_yref; bytes = B * N * CF_NY * 8; }
    else if (!strcmp(field, "yref_e")) { dst = h->d_yref_e; bytes = B * CF_NX * 8; }
    else if (!strcmp(field, "x")) { dst = h->d_x; bytes = B * (N + 1) * CF_NX * 8; h->prepared = false; }
    else if (!strcmp(field, "u")) { dst = h->d_u; bytes = B * N * CF_NU * 8; h->prepared = false; }
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_set: unknown field '") + field + "' (generic-model build)");
    CK(cudaMemcpyAsync(dst, src, bytes, kind, h->stream));
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_set_option(cfnmpc_batch *h, const char *option, int value)
{
    if (!h || !option) return fail(CFNMPC_EINVAL, "cfnmpc_batch_set_option: null argument");
    if (!strcmp(option, "max_ipm_iter")) h->P.max_ipm_iter = (value > 0 && value < CF_ITER_MAX) ? value : CF_ITER_MAX;
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_set_option: unknown option '") + option + "' (generic-model build)");
    return CFNMPC_OK;
}

static int launch_prep(cfnmpc_batch *h)
{
    CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    cf_rti_kernel<4, 3, CF_PH_PREPARATION, true><<<h->grid, 128, h->smem_prep, h->stream>>>(h->P, h->bv);
    CK(cudaGetLastError());
    h->launches++;
    return CFNMPC_OK;
}
static int launch_fb(cfnmpc_batch *h)
{
    CK(cudaMemsetAsync(h->d_counter, 0, 4, h->stream));
    cf_pcond_kernel<4, 3, 1><<<h->grid, 128, h->smem_fb, h->stream>>>(h->P, h->bv, h->pcb);
    CK(cudaGetLastError());
    h->launches++;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_solve(cfnmpc_batch *h, int n_rti)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    if (n_rti < 1) return fail(CFNMPC_EINVAL, "cfnmpc_batch_solve: n_rti must be >= 1");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int r = 0; r < n_rti; r++) {
        if (int rc = launch_prep(h)) return rc;
        if (int rc = launch_fb(h)) return rc;
    }
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->prepared = false;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_prepare(cfnmpc_batch *h)
{
    if (!h) return fail(CFNMPC_EINVAL, "null handle");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev0, h->stream));
    if (int rc = launch_prep(h)) return rc;
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
    CK(cudaEventRecord(h->ev0, h->stream));
    if (int rc = launch_fb(h)) return rc;
    CK(cudaEventRecord(h->ev1, h->stream));
    h->timed = true;
    h->prepared = false;
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
    const size_t B = h->B, N = h->N;
    const void *src = nullptr;
    size_t bytes = 0;
    if (!strcmp(field, "x") || !strcmp(field, "u")) {   // one stage of every instance: a strided 2-D copy
        const bool is_u = field[0] == 'u';
        const size_t w = (is_u ? CF_NU : CF_NX) * 8, nst = is_u ? N : N + 1;
        if (stage < 0 || (size_t) stage >= nst) return fail(CFNMPC_EINVAL, "cfnmpc_batch_get: stage out of range");
        CK(cudaMemcpy2DAsync(dst, w, (const char *) (is_u ? h->d_u : h->d_x) + (size_t) stage * w, nst * w, w, B, kind, h->stream));
    } else {
        if (!strcmp(field, "x_all")) { src = h->d_x; bytes = B * (N + 1) * CF_NX * 8; }
        else if (!strcmp(field, "u_all")) { src = h->d_u; bytes = B * N * CF_NU * 8; }
        else if (!strcmp(field, "status")) { src = h->d_status; bytes = B * 4; }
        else if (!strcmp(field, "qp_iter")) { src = h->d_qp_iter; bytes = B * 4; }
        else if (!strcmp(field, "qp_status")) { src = h->d_qp_status; bytes = B * 4; }
        else if (!strcmp(field, "flags")) { src = h->d_flags; bytes = B * 4; }
        else if (!strcmp(field, "res")) { src = h->d_res; bytes = B * 4 * 8; }
        else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_get: unknown field '") + field + "' (generic-model build)");
        CK(cudaMemcpyAsync(dst, src, bytes, kind, h->stream));
    }
    if (!dst_on_device) CK(cudaStreamSynchronize(h->stream));
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_last_solve_ms(cfnmpc_batch *h, double *ms)
{
    if (!h || !ms) return fail(CFNMPC_EINVAL, "null argument");
    if (!h->timed) return fail(CFNMPC_ESTATE, "no solve has been enqueued yet");
    CK(cudaSetDevice(h->device));
    CK(cudaEventSynchronize(h->ev1));
    float f = 0;
    CK(cudaEventElapsedTime(&f, h->ev0, h->ev1));
    *ms = f;
    return CFNMPC_OK;
}

extern "C" int cfnmpc_batch_info(cfnmpc_batch *h, const char *what, long long *value)
{
    if (!h || !what || !value) return fail(CFNMPC_EINVAL, "null argument");
    if (!strcmp(what, "launches")) *value = h->launches;
    else if (!strcmp(what, "grid")) *value = h->grid;
    else if (!strcmp(what, "regs_preparation")) *value = h->regs_prep;
    else if (!strcmp(what, "regs_feedback")) *value = h->regs_fb;
    else if (!strcmp(what, "nx")) *value = CF_NX;
    else if (!strcmp(what, "nu")) *value = CF_NU;
    else return fail(CFNMPC_EINVAL, std::string("cfnmpc_batch_info: unknown item '") + what + "'");
    return CFNMPC_OK;
}
