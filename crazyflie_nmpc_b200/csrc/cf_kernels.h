// The persistent-warp kernels around the warp programs (cf_rti_warp.h, cf_pcond_warp.h), shared by the Crazyflie library
// (cfnmpc_api.cu) and the generic-model library (cfnmpc_generic.cu).
#pragma once
#include <cuda_runtime.h>

#include "cf_rti_warp.h"
#include "cf_pcond_warp.h"

// ------------------------------------------------------------------ kernels
// Persistent warps: each warp owns a scratch slot and pulls instance ids from a global
// counter until the batch is exhausted (IPM trip counts differ per instance: 4..11).
// Launch shape <WPB, MINB>: WPB warps per block (warps never cooperate, the block is only a container), MINB resident
// blocks per SM the register allocation is bounded for: warps per SM = WPB * MINB, registers <= 65536 / (32 * WPB * MINB).
template <int WPB, int PH, bool VDT>
__device__ __forceinline__ void cf_rti_kernel_body(const CfParams &P, const CfBatchView &bv)
{
    extern __shared__ __align__(128) double cf_smem[];
    const int warp = threadIdx.x >> 5;
    double *sm = cf_smem + warp * CF_SM_DOUBLES;
    double *slot = bv.scratch + (long) (blockIdx.x * WPB + warp) * bv.scratch_stride;
    cf_warp_init_smem(sm);
    unsigned par = 0;
    for (;;) {
        int inst = 0;
        if ((threadIdx.x & 31) == 0) inst = atomicAdd(bv.counter, 1);
        inst = __shfl_sync(0xffffffffu, inst, 0);
        if (inst >= bv.B) break;
        if (bv.ready) {
            // inputs of this instance may still be on their way from the host: wait for the upload front to pass it
            // (bounded: under a profiler that replays the kernel without the copies, or if the host died, give up after
            // 5 s and mark the instance instead of hanging the device)
            int late = 0;
            if ((threadIdx.x & 31) == 0) {
                int r;
                unsigned long long t0 = 0, t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                for (;;) {
                    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(r) : "l"(bv.ready) : "memory");
                    if (r > inst) break;
                    __nanosleep(500);
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                    if (t1 - t0 > 5000000000ull) { late = 1; break; }
                }
            }
            late = __shfl_sync(0xffffffffu, late, 0);
            cf_rti_instance<PH, VDT>(&P, bv, bv.first + inst, slot, sm, par);
            if (late && (threadIdx.x & 31) == 0) bv.flags[bv.first + inst] |= CF_FLAG_INPUT_LATE;
            continue;
        }
        cf_rti_instance<PH, VDT>(&P, bv, bv.first + inst, slot, sm, par);
    }
}
// <PH, VDT>: preparation + feedback (0) / preparation (1) / feedback (2); uniform or per-interval time steps.  The
// benchmarked path is <.., 0, false>.
template <int WPB, int MINB, int PH = CF_PH_BOTH, bool VDT = false>
__global__ void __launch_bounds__(WPB * 32, MINB)
cf_rti_kernel(const __grid_constant__ CfParams P, const __grid_constant__ CfBatchView bv)
{
    cf_rti_kernel_body<WPB, PH, VDT>(P, bv);
}
// Feedback phase with the QP partially condensed to blk.N2 stages (cf_pcond_warp.h): same persistent-warp scheme; needs
// the prepared linearisations of a preparation kernel.  BS = stages per block (the reference's qp_cond_N = ceil(N / BS)..).
template <int WPB, int MINB, int BS>
__global__ void __launch_bounds__(WPB * 32, MINB)
cf_pcond_kernel(const __grid_constant__ CfParams P, const __grid_constant__ CfBatchView bv, const __grid_constant__ CfPcBlocks blk)
{
    extern __shared__ __align__(128) double cf_smem[];
    const int warp = threadIdx.x >> 5;
    double *sm = cf_smem + warp * CfPcWarpT<BS>::SM_DOUBLES;
    double *slot = bv.scratch + (long) (blockIdx.x * WPB + warp) * bv.scratch_stride;
    {
        uint64_t *bar = reinterpret_cast<uint64_t *>(sm + CfPcWarpT<BS>::SM_BAR);
        if ((threadIdx.x & 31) == 0) { cf_mbar_init(bar); cf_mbar_init(bar + 1); }
        __syncwarp();
    }
    unsigned par = 0;
    for (;;) {
        int inst = 0;
        if ((threadIdx.x & 31) == 0) inst = atomicAdd(bv.counter, 1);
        inst = __shfl_sync(0xffffffffu, inst, 0);
        if (inst >= bv.B) break;
        cf_pcond_instance<BS>(&P, bv, blk, bv.first + inst, slot, sm, par);
    }
}
// out[i][0:w] = src[i][stage*w : stage*w + w]   (ocp_nlp_out_get for every instance at once)
__global__ void cf_gather_stage_kernel(const double *__restrict__ src, double *__restrict__ out, int B, int per_inst, int stage, int w)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < B * w) {
        const int i = t / w, j = t - i * w;
        out[t] = src[(long) i * per_inst + stage * w + j];
    }
}

