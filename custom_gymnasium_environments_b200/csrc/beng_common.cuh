// Shared host/device helpers of libbeng: launch accounting, error mapping, PTX wrappers for the
// bulk asynchronous copy engine (TMA, non-tensor form) used to drain observation tiles.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdlib>

#include "../../include/beng.h"

namespace beng {

extern std::atomic<uint64_t> g_launch_count;

inline int finish_launch() {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

int device_sm_count();

// ---- bulk async copy (cp.async.bulk, SASS UBLKCP) shared -> global -------------------------
// One thread issues a contiguous copy of `bytes` (multiple of 16, both addresses 16-B aligned)
// from this CTA's shared memory to global memory; completion is tracked per issuing thread in
// bulk async-groups.
__device__ __forceinline__ void bulk_store_s2g(void *gdst, const void *ssrc, uint32_t bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(s), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until at most N of this thread's bulk groups still have to READ their shared-memory source.
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// Wait until at most N of this thread's bulk groups are incomplete (writes performed).
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// Make generic-proxy writes to shared memory visible to the async proxy (the copy engine).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// A step kernel depends on the previous step's state writes.  Launched with programmatic stream serialization, its
// CTAs may become resident while the previous grid drains; everything before pdl_wait() (index math, shared-memory
// setup) overlaps that tail, and pdl_wait() returns once the previous grid has completed and flushed its memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch `kern<<<grid, block, smem, stream>>>(arg)` with the PDL attribute (BENG_NO_PDL=1 disables it).
template <typename Kern, typename Arg>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const Arg &arg) {
    static const bool use_pdl = getenv("BENG_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

// streaming (evict-first) 128-bit / 64-bit global loads for read-once inputs
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// (coherent, not .nc: under programmatic dependent launch the producer of the action tensor may still be running when
// this grid becomes resident, and .nc data must be read-only for the whole lifetime of the grid)
__device__ __forceinline__ long long ld_stream_s64(const long long *p) {
    long long r;
    asm volatile("ld.global.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}

// Correctly rounded x / B (float64) for a compile-time integer B without the IEEE division subroutine (a ~60-instruction
// call on sm_100a): y = RN(1 / B), q = RN(x * y), r = x - B * q exactly (FMA), result RN(q + r * y) -- Markstein's
// final iteration, which returns the correctly rounded quotient when y is the correctly rounded reciprocal and q is
// within an ulp of x / B.  For finite x whose quotient neither overflows nor is subnormal (the callers divide sums of
// small integers).  tests/test_traffic_oracle.py restates the sequence in exact rational arithmetic against x / B.
template <int B>
__device__ __forceinline__ double div_const(double x) {
    constexpr double y = 1.0 / (double)B;
    const double q = x * y;
    const double r = fma(-(double)B, q, x);
    return fma(r, y, q);
}

}  // namespace beng
