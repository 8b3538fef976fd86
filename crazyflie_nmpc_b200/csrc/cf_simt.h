// Warp-level primitives used by the RTI warp program (cf_rti_warp.h).
//
// Product build (nvcc, sm_100a): thin wrappers over the CUDA warp intrinsics, mbarrier and
// the TMA bulk-copy engine (cp.async.bulk, SASS UBLKCP).
//
// CF_SIMT_EMU build (g++, tests only): the same warp program is compiled for the host and
// its 32 lanes are run as lock-step fibers (tests/simt_emu/).  This is a
// debugging/verification harness for the kernel SOURCE on machines without a GPU; it is
// not reachable from the library's API and is not a CPU fallback.
#pragma once

#include <math.h>
#include <stdint.h>

struct cf_d2  // two consecutive doubles, 16-byte aligned (LDS.128 / LDG.128)
{
    double x, y;
};

#if defined(CF_SIMT_EMU)
// ---------------------------------------------------------------- host emulation
#define CF_DEV static inline
#define CF_MEM inline
#define CF_NOINLINE __attribute__((noinline))
#define CF_UNROLL _Pragma("GCC unroll 32")
#define CF_NOUNROLL _Pragma("GCC unroll 1")
namespace cfemu {
int lane();
void barrier(int line);
double exch(double v, int src, int line);
int atomic_add(int *p, int v);
void bulk_expect(uint64_t *bar, int bytes, int line);
void bulk_g2s(void *dst, const void *src, int bytes, uint64_t *bar, int line);
void bulk_wait(uint64_t *bar, unsigned parity, int line);
void bulk_s2g(void *dst, const void *src, int bytes, int line);
void bulk_s2g_wait(int max_pending, int line);
void dmma(double &d0, double &d1, double a, double b, int line);
}  // namespace cfemu
CF_DEV int cf_lane() { return cfemu::lane(); }
#define cf_syncwarp() cfemu::barrier(__LINE__)
#define cf_shfl(v, src) cfemu::exch((v), (src), __LINE__)
CF_DEV int cf_atomic_add(int *p, int v) { return cfemu::atomic_add(p, v); }
CF_DEV cf_d2 cf_ld2(const double *p)
{
    if (((uintptr_t) p) & 15) __builtin_trap();  // a misaligned 128-bit access faults on the GPU
    cf_d2 r = {p[0], p[1]};
    return r;
}
CF_DEV void cf_st2(double *p, double a, double b)
{
    if (((uintptr_t) p) & 15) __builtin_trap();
    p[0] = a; p[1] = b;
}
CF_DEV void cf_mbar_init(uint64_t *bar) { *bar = 0; }
CF_DEV double cf_rsqrt_seed(double x) { return (double) (float) (1.0 / sqrt(x)); }  // ~2^-23, like MUFU.RSQ64H
CF_DEV double cf_rcp_seed(double x) { return (double) (float) (1.0 / x); }            // ~2^-23, like MUFU.RCP64H
#define cf_bulk_expect(bar, bytes) cfemu::bulk_expect((bar), (bytes), __LINE__)
#define cf_bulk_g2s_raw(dst, src, bytes, bar) cfemu::bulk_g2s((dst), (src), (bytes), (bar), __LINE__)
#define cf_bulk_wait(bar, parity) cfemu::bulk_wait((bar), (parity), __LINE__)
#define cf_bulk_s2g(dst, src, bytes) cfemu::bulk_s2g((dst), (src), (bytes), __LINE__)
#define cf_bulk_s2g_wait_all() cfemu::bulk_s2g_wait(0, __LINE__)
#define cf_bulk_s2g_wait_read1() cfemu::bulk_s2g_wait(1, __LINE__)
#define cf_bulk_s2g_wait_read0() cfemu::bulk_s2g_wait(0, __LINE__)
CF_DEV void cf_fence_proxy_async() {}
CF_DEV void cf_keep(int &) {}
#define cf_dmma(d0, d1, a, b) cfemu::dmma((d0), (d1), (a), (b), __LINE__)
#else
// ---------------------------------------------------------------- CUDA (sm_100a)
#define CF_DEV __device__ __forceinline__
#define CF_MEM __device__ __forceinline__
#define CF_NOINLINE __noinline__
#define CF_UNROLL _Pragma("unroll")
#define CF_NOUNROLL _Pragma("unroll 1")
CF_DEV int cf_lane() { return threadIdx.x & 31; }
CF_DEV void cf_syncwarp_() { __syncwarp(); }
#define cf_syncwarp() cf_syncwarp_()
CF_DEV double cf_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
CF_DEV int cf_atomic_add(int *p, int v) { return atomicAdd(p, v); }
CF_DEV cf_d2 cf_ld2(const double *p)
{
    const double2 v = *reinterpret_cast<const double2 *>(p);
    cf_d2 r = {v.x, v.y};
    return r;
}
CF_DEV void cf_st2(double *p, double a, double b) { *reinterpret_cast<double2 *>(p) = make_double2(a, b); }

CF_DEV double cf_rsqrt_seed(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
CF_DEV double cf_rcp_seed(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}
CF_DEV uint32_t cf_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
// mbarrier with one expected arrival (the lane that issues the bulk copies)
CF_DEV void cf_mbar_init(uint64_t *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(cf_smem_u32(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async proxy
}
// One lane: announce `bytes` and start the TMA bulk copy global -> shared (completes on `bar`).
CF_DEV void cf_bulk_expect(uint64_t *bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cf_smem_u32(bar)), "r"(bytes) : "memory");
}
CF_DEV void cf_bulk_g2s_raw(void *dst, const void *src, int bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(cf_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(cf_smem_u32(bar))
                 : "memory");
}
// All lanes: wait until the phase with the given parity of `bar` has completed.
CF_DEV void cf_bulk_wait(uint64_t *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CF_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra CF_WAIT_%=;\n"
        "}\n" ::"r"(cf_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// One lane: TMA bulk copy shared -> global, committed as its own bulk group.
CF_DEV void cf_bulk_s2g(void *dst, const void *src, int bytes)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(cf_smem_u32(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// Issuing lane: all its bulk stores have completed (global writes performed, shared source free).
CF_DEV void cf_bulk_s2g_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// Issuing lane: all but the most recent bulk store have finished READING their shared source.
CF_DEV void cf_bulk_s2g_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// Issuing lane: every bulk store has finished READING its shared source (the staging area may be refilled).
CF_DEV void cf_bulk_s2g_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// fp64 tensor-core tile product D(8x8) += A(8x4) * B(4x8), fragments as in the PTX ISA for m8n8k4:
// a = A[lane>>2][lane&3], b = B[lane&3][lane>>2], d0/d1 = D[lane>>2][2*(lane&3) + {0,1}]   (SASS DMMA)
CF_DEV void cf_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// Order earlier generic-proxy accesses before later async-proxy (TMA) accesses.
CF_DEV void cf_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Make a loop-invariant index opaque to the compiler, so that it is kept (register or local memory) instead of being
// recomputed inside a stage loop (ptxas rematerialised 26 instructions per stage for three packed-triangle indices).
CF_DEV void cf_keep(int &v) { asm volatile("" : "+r"(v)); }
#endif

// sqrt(x) and 1/sqrt(x) for x > 0 in the normal range: hardware seed (2^-22) refined by two coupled
// Goldschmidt steps and one final correction of the root -- branch-free, ~14 instructions instead of the
// ~55 of an IEEE sqrt followed by an IEEE divide; both results are within 1-2 ulp.
CF_DEV void cf_sqrt_rsqrt(double x, double &root, double &inv)
{
    const double y = cf_rsqrt_seed(x);
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, x);
    root = fma(d, h, g);
    inv = h + h;
}

// 1/x for x in the normal range (slacks, multipliers, pivots): hardware seed (2^-23) + two Newton steps; branch-free,
// no slow-path call, within 1 ulp of the IEEE quotient.
CF_DEV double cf_rcp(double x)
{
    double y = cf_rcp_seed(x);
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// acc = max(acc, |x|) as a compare + select (fmax costs a NaN-aware sequence per call); a NaN in x is ignored, as by fmax
CF_DEV void cf_amax(double &acc, double x)
{
    const double ax = fabs(x);
    acc = (ax > acc) ? ax : acc;
}

// butterfly reductions (all lanes get the result).  Called a dozen times per interior-point iteration, never inside a
// stage loop: ONE out-of-line copy each -- inlined and unrolled they were 3 000 of the feedback kernel's 11 000 SASS
// instructions (every 64-bit exchange carries its divergent-warp fallback), and the kernel is sensitive to code size
// (profiles/README.md, v18/v19).
#if defined(CF_SIMT_EMU)
#define CF_OUTLINE static inline
#else
#define CF_OUTLINE static __device__ __noinline__
#endif
CF_OUTLINE double cf_warp_sum(double v)
{
    CF_NOUNROLL
    for (int o = 16; o > 0; o >>= 1) v += cf_shfl(v, cf_lane() ^ o);
    return v;
}
CF_OUTLINE double cf_warp_max(double v)
{
    CF_NOUNROLL
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, cf_shfl(v, cf_lane() ^ o));
    return v;
}
