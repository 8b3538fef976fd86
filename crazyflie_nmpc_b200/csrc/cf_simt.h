// Warp-level primitives used by the RTI warp program (cf_rti_warp.h).
//
// Product build (nvcc, sm_100a): thin wrappers over the CUDA warp intrinsics.
//
// CF_SIMT_EMU build (g++, tests only): the same warp program is compiled for the
// host and its 32 lanes are run as lock-step fibers (tests/simt_emu/).  This is a
// debugging/verification harness for the kernel SOURCE on machines without a GPU;
// it is not reachable from the library's API and is not a CPU fallback.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(CF_SIMT_EMU)
// ---------------------------------------------------------------- host emulation
#define CF_DEV static inline
#define CF_MEM inline
#define CF_DEV_NOINLINE static
#define CF_UNROLL _Pragma("GCC unroll 32")
namespace cfemu {
int lane();
void barrier(int line);
double exch(double v, int src, int line);
int atomic_add(int *p, int v);
}  // namespace cfemu
CF_DEV int cf_lane() { return cfemu::lane(); }
#define cf_syncwarp() cfemu::barrier(__LINE__)
#define cf_shfl(v, src) cfemu::exch((v), (src), __LINE__)
CF_DEV int cf_atomic_add(int *p, int v) { return cfemu::atomic_add(p, v); }
CF_DEV double cf_ldg(const double *p) { return *p; }
CF_DEV double cf_rsqrt(double x) { return 1.0 / sqrt(x); }
CF_DEV double cf_rcp(double x) { return 1.0 / x; }
#else
// ---------------------------------------------------------------- CUDA
#define CF_DEV __device__ __forceinline__
#define CF_MEM __device__ __forceinline__
#define CF_DEV_NOINLINE __device__ __noinline__
#define CF_UNROLL _Pragma("unroll")
CF_DEV int cf_lane() { return threadIdx.x & 31; }
CF_DEV void cf_syncwarp_() { __syncwarp(); }
#define cf_syncwarp() cf_syncwarp_()
CF_DEV double cf_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
CF_DEV int cf_atomic_add(int *p, int v) { return atomicAdd(p, v); }
CF_DEV double cf_ldg(const double *p) { return __ldg(p); }
CF_DEV double cf_rsqrt(double x) { return rsqrt(x); }
CF_DEV double cf_rcp(double x) { return 1.0 / x; }
#endif

// butterfly reductions (all lanes get the result)
CF_DEV double cf_warp_sum(double v)
{
    CF_UNROLL
    for (int o = 16; o > 0; o >>= 1) v += cf_shfl(v, cf_lane() ^ o);
    return v;
}
CF_DEV double cf_warp_max(double v)
{
    CF_UNROLL
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, cf_shfl(v, cf_lane() ^ o));
    return v;
}
CF_DEV double cf_warp_min(double v)
{
    CF_UNROLL
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, cf_shfl(v, cf_lane() ^ o));
    return v;
}
