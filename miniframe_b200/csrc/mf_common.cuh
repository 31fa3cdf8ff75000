// Shared host/device helpers for the sm_100a kernels of miniframe_b200.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/miniframe_b200.h"

// Tensor maps (TMA descriptors) are encoded on the host; a handle keeps the last few, keyed by
// everything that goes into the encoding, so a sampler that calls with the same workspace
// thousands of times encodes once.
struct mf_tmap_key {
    int kind;                 // 1: factorisation ring (5-D load), 2: covariance tiles (3-D store)
    const void* base;
    int64_t ld, stride;
    int n, batch;
};
struct mf_tmap_slot {
    mf_tmap_key key;
    alignas(64) CUtensorMap map;
};
constexpr int MF_TMAP_SLOTS = 8;

// options a caller (the tests, the tuning scripts) may set through mf_set_option
struct mf_options {
    int potrf_tile;           // 0 = choose, 1 = 64-row tiles, 2 = 128-row tiles
    int potrf_group;          // 0 = choose, 1 = one CTA per matrix, G > 1 = G CTAs per matrix
    int potrf_map;            // 0 / 1 = tiles walk the CTAs backwards (it - jb), 2 = row blocks stay with a CTA
    int potrf_lookahead;      // 0 = choose, 1 = plain one-CTA-per-matrix kernel, 2 = look-ahead kernel
    int cov_ctas;             // 0 = choose, else CTAs per sample of the covariance build
    int experiment;           // timing-experiment variant (only with -DMF_EXPERIMENTS builds)
};

struct mf_handle_s {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t smem_optin;
    // Device scratch owned by the handle, allocated once by mf_create (2 MB): the hand-out
    // counter of the persistent factorisation and, for its cooperative variant (several CTAs per
    // matrix), per group 8 synchronisation words followed by the pool of `rowdone` words (one per
    // 64-row block of the group's matrix).  Launches on one handle share it: one handle per stream.
    unsigned* coop_sync;
    int coop_groups;
    int coop_words;           // words in the rowdone pool (behind the 8 x coop_groups header words)
    mf_options opt;
    mf_tmap_slot tmaps[MF_TMAP_SLOTS];
    int tmap_next;
    char err[512];
};

// cuTensorMapEncodeTiled through the runtime's driver entry point query (no link-time libcuda)
typedef CUresult (*mf_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
mf_tmap_encode_fn mf_tmap_encoder(mf_handle_s* h);
// cached lookup: returns the slot holding `key`'s map, or nullptr with *fresh pointing at the slot
// the caller must encode into
const CUtensorMap* mf_tmap_find(mf_handle_s* h, const mf_tmap_key& key, mf_tmap_slot** fresh);

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
