// Shared host/device helpers for the sm_100a kernels of miniframe_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/miniframe_b200.h"

struct mf_handle_s {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t smem_optin;
    // scratch of the cooperative (several CTAs per matrix) factorisation, allocated once by
    // mf_create: per group 8 synchronisation words and the 8 negated 8x8 inverses (512 doubles)
    unsigned* coop_sync;
    double* coop_dinv;
    int coop_groups;
    char err[512];
};

#define MF_CUDA_CHECK(h, call)                                                        \
    do {                                                                              \
        cudaError_t _e = (call);                                                      \
        if (_e != cudaSuccess) {                                                      \
            snprintf((h)->err, sizeof((h)->err), "%s:%d %s -> %s", __FILE__, __LINE__, \
                     #call, cudaGetErrorString(_e));                                  \
            return MF_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

#define MF_FAIL(h, code, ...)                                  \
    do {                                                       \
        snprintf((h)->err, sizeof((h)->err), __VA_ARGS__);     \
        return (code);                                         \
    } while (0)

// ---- layout -----------------------------------------------------------------
__host__ __device__ inline int64_t mf_ld_of(int n) { return ((int64_t)n + 15) / 16 * 16; }
// rows allocated per matrix: the n covariance rows + the z row, rounded up to whole 8-row groups
// (the TMA box of the factorisation kernel moves 8-row groups)
__host__ __device__ inline int64_t mf_rows_of(int n) { return ((int64_t)n + 1 + 7) / 8 * 8; }

// ---- device primitives --------------------------------------------------------
#ifdef __CUDACC__

// FP64 tensor-core MMA, D(8x8) = A(8x4) * B(4x8) + C.  Lane l = 4*g + t holds
// A[g][t], B[t][g], C/D[g][2t], C/D[g][2t+1].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 16-byte async global->shared copy, L2 only (.cg); src_bytes = 0 zero-fills.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
