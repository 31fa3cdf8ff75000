// Batched blocked Cholesky fused with forward substitution, log-determinant and
// the quadratic form: one launch replaces cho_factor + cho_solve + sum(log(diag))
// of the reference's log_likelihood (BIGgp.py:305-309, BIGgpPlusJitt.py:311-315,
// SMALLgp.py:269-273) for a whole batch of hyper-parameter sets.
//
// Design (B200: 148 SMs, FP64 DMMA, 126 MB L2, HBM3e):
//   * one matrix per CTA, two CTAs resident per SM (three, on 64-row tiles, for matrices
//     of order <= 1100), persistent over the batch with the matrices handed out through an
//     atomic counter: no inter-CTA synchronisation; the serial diagonal-block phases of
//     one CTA are hidden behind the other CTAs' tensor-core work.  Batches smaller than
//     the GPU: several CTAs per matrix (COOP, see the kernel's comment);
//   * LEFT-looking by 64-wide block columns: tile C(128 x 64) = K - L_left L_top^T
//     is a dense contraction over the already-final columns, accumulated in
//     registers with mma.sync m8n8k4 f64 (SASS DMMA.8x8x4).  The operands stream
//     through a 3-stage shared-memory ring fed by TMA (cp.async.bulk.tensor.5d,
//     SASS UTMALDG.5D, 128-byte hardware swizzle): every warp waits on the stage's
//     mbarrier, and the last warp to release a stage refills it - no CTA-wide barrier
//     and no producer wait inside the loop.
//     Every tile of K is read once and every tile of L written once (HBM traffic
//     ~ n^3/(6*64)*8 B per matrix, arithmetic intensity 16 flop/B);
//   * each warp owns complete rows (two 8-row fragments x 64 columns), so the triangular
//     solve against the diagonal block runs entirely in registers on the tensor
//     cores: the C-fragment layout (row g, columns 2t, 2t+1) is reused directly as
//     the A operand with the k index permuted, the matching B operand being one
//     16-byte shared-memory load;
//   * the residual vector rides along as row n of the matrix: its row of L is
//     z = L^-1 r, so the forward substitution costs one extra row in the last row
//     tile, and -1/2 z.z - sum(log L_ii) - n/2 log(2 pi) falls out at the end.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "mf_common.cuh"

#ifdef MF_EXPERIMENTS
#define MF_EXPERIMENT_OF(h) ((h)->opt.experiment)
#else
#define MF_EXPERIMENT_OF(h) 0
#endif

namespace pf {

constexpr int TM = 128;       // rows per tile (8 warps x 16)
constexpr int NB = 64;        // block-column width
constexpr int KC = 16;        // contraction chunk per pipeline stage
constexpr int THREADS = 256;
constexpr int STAGE_ROWS = TM + NB;
constexpr int STAGE_BYTES = STAGE_ROWS * KC * 8;     // 24576
constexpr int AW_LD = 65;
// L_jj is kept as its 36 lower 8x8 blocks, 512 B each (block (bi, bj) at bi(bi+1)/2 + bj):
// a 16-byte fragment load of rows g, g+1 then covers one 128-byte line, conflict-free.
constexpr int SZ_LD = 36 * 64 * 8;
constexpr int SZ_DINV = 8 * 64 * 8;                  // minus-inverses of the 8 diagonal 8x8 blocks
constexpr int SZ_COL = 2 * 72 * 8;                   // double-buffered pivot column: 64 values + status word
constexpr int SZ_RED = 64 * 8;                       // reduction scratch, next-matrix word, mbarriers, release counters
constexpr int smem_bytes(int nstage, int nmi = 2) { return nstage * (64 * nmi + NB) * KC * 8 + SZ_LD + SZ_DINV + SZ_COL + SZ_RED + 1024; }
static_assert(NB * AW_LD * 8 <= 3 * (64 + NB) * KC * 8,
              "potf2 staging array must fit in the stage area");
// 128-row tiles (NMI = 2): 3 stages, 97,792 B, two CTAs per SM; 64-row tiles (NMI = 1): 73,216 B, three

struct Params {
    // 5-D tensor map over the chunk of matrices for the TMA-fed ring:
    // {column, row 0/2/4/6 of an 8-row group, row parity, 8-row group, matrix}
    alignas(64) CUtensorMap tmap;
    // the same view with boxes of 128 rows: the row tile of the 128-row kernel in ONE TMA instruction
    alignas(64) CUtensorMap tmap128;
    int batch;
    int n;
    double* K;
    int64_t ld;
    int64_t stride;
    const double* resid;
    int64_t resid_stride;
    double* ll;
    int32_t* info;
    int sm_count;
    // cooperative variant (COOP): `group` CTAs share one matrix
    int group;
    unsigned* sync;       // per group 8 words: [0] barrier arrivals, [2] failed matrix ordinal + 1,
                          // [3] info of the failure
    unsigned* rowdone;    // per group rd_stride words: block columns finished per 64-row block
    int rd_stride;
    int map_back;         // tile `it` of block column jb goes to CTA (it - jb) mod G instead of (it + jb) mod G
    int* counter;         // next matrix to hand out (one CTA per matrix, zeroed before the launch)
    int quota;            // most matrices one CTA may take (0: no limit), see mf_potrf_loglike
};

__device__ __forceinline__ double flip(double x) {      // -x on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}
__device__ __forceinline__ int ldblk(int bi, int bj) { return (bi * (bi + 1) / 2 + bj) * 64; }

// ---- mbarrier primitives (shared::cta) ---------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MF_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra MF_WAIT_%=;\n\t}"
        ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one 64-row x 16-column box of matrix `mat`, rows 8 * blk8 .., columns c0 ..; lands 128B-swizzled
__device__ __forceinline__ void tma_box(uint32_t dst, const CUtensorMap* tm, int c0, int blk8, int mat, uint32_t bar,
                                        uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
        ::"r"(dst), "l"(tm), "r"(c0), "r"(0), "r"(0), "r"(blk8), "r"(mat), "r"(bar), "l"(policy)
        : "memory");
}
// One chunk of the ring - expect_tx and the three boxes - issued by ONE elected lane of a fully
// active warp whose operands are warp-uniform: no divergent branch around it, so the instructions
// can be scheduled between the tensor-core instructions of the warp that issues them.
__device__ __forceinline__ void tma_chunk_elect(uint32_t dst, const CUtensorMap* tmA, const CUtensorMap* tmB, int c0,
                                                int blkA, int blkB, int mat, uint32_t bar, uint32_t bytes,
                                                uint32_t b_off) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 pf, pl;\n\t.reg .b32 d3;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "createpolicy.fractional.L2::evict_first.b64 pf, 1.0;\n\t"
        "createpolicy.fractional.L2::evict_last.b64 pl, 1.0;\n\t"
        "add.u32 d3, %0, %9;\n\t"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%6], %7;\n\t"
        "@p cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, 0, 0, %3, %5}], [%6], pf;\n\t"
        "@p cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[d3], [%8, {%2, 0, 0, %4, %5}], [%6], pl;\n\t}"
        ::"r"(dst), "l"(tmA), "r"(c0), "r"(blkA), "r"(blkB), "r"(mat), "r"(bar), "r"(bytes), "l"(tmB), "r"(b_off)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() {
    asm volatile("fence.proxy.async;" ::: "memory");
}

// lane 0 adds one to a shared-memory counter without a branch; the old value is looked at later
__device__ __forceinline__ void atoms_inc_lane0(unsigned& old, uint32_t addr, int lane) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.s32 p, %2, 0;\n\t"
        "@p atom.shared.add.u32 %0, [%1], 1;\n\t}"
        : "+r"(old)
        : "r"(addr), "r"(lane)
        : "memory");
}

// MINUS the inverses of the eight diagonal 8 x 8 blocks of L_jj (in Ld), the operand of the row
// tiles' triangular solve: thread 8 bb + cc computes column cc of block bb.  One routine for the
// CTA that factorised L_jj and for the CTAs that fetched it: the same bits everywhere.
__device__ __forceinline__ void dinv_blocks(const double* Ld, double* Dinv, int tid) {
    if (tid < 64) {
        const int bb = tid >> 3, cc = tid & 7;
        const double* Lb = Ld + ldblk(bb, bb);
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double sum = (i == cc) ? -1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) sum -= Lb[i * 8 + k] * x[k];
            x[i] = sum / Lb[i * 8 + i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Dinv[bb * 64 + i * 8 + cc] = x[i];
    }
}

// Factorisation of the 64 x 64 diagonal block C (in Aw, row stride AW_LD) into Ld (8 x 8 blocks)
// and Dinv (minus the inverses of the diagonal 8 x 8 blocks); a failed pivot is reported through
// *failword (1-based global index).  Called by all 256 threads; the caller's barrier follows.
//
// Warp p owns the 8-column panel p, LEFT-looking, 8 x 8 blocks in the tensor-core accumulator
// layout (lane 4g + t: row g, columns 2t, 2t + 1):
//   1. diagonal block (p, p): panels q < p are applied as they are published (one mbarrier per
//      panel, two DMMAs each), then eight pivots with warp shuffles only - no CTA-wide barrier
//      per pivot; the critical path of the whole block is 64 x (shuffle, rsqrt, multiply,
//      shuffle, fma);
//   2. minus the inverse of the factored diagonal block, eight lanes, one column each;
//   3. the blocks below it, two at a time: apply the panels q < p, multiply by the inverse
//      (the same tensor-core triangular solve the row tiles use), store, publish the panel.
// Register needs stay small (one 8 x 8 block per warp at a time) on purpose: this code shares
// the register allocation of the contraction loop.
template <int DBG>
__device__ __forceinline__ void potf2_panels(const double* Aw, double* Ld, double* Dinv, double* rsbuf,
                                             uint32_t pbar, volatile int* failword, int nv, int j0,
                                             uint32_t parity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const unsigned FULL = 0xffffffffu;
    const int c0 = 8 * warp + 2 * t;
    // block (bi, warp) of C, negated (the DMMAs ADD L L^T)
    auto load_neg = [&](int bi, double& v0, double& v1) {
        const int r = 8 * bi + g;
        v0 = flip(Aw[r * AW_LD + c0]);
        v1 = flip(Aw[r * AW_LD + c0 + 1]);
    };
    // columns beyond the matrix order are replaced by the identity once the finished panels have
    // been applied (the residual row, which follows the last matrix row, would otherwise leak
    // into them); that keeps the inverses finite and those columns out of every update
    const bool live0 = c0 < nv, live1 = c0 + 1 < nv;
    double dg0, dg1;
    load_neg(warp, dg0, dg1);
    bool dead = (DBG & 32) != 0;
    for (int q = 0; q < warp && !dead; ++q) {
        mbar_wait(pbar + 8 * q, parity);
        if (*failword) {
            dead = true;
            break;
        }
        const double2 bq = *reinterpret_cast<const double2*>(Ld + ldblk(warp, q) + g * 8 + 2 * t);
        dmma884(dg0, dg1, bq.x, bq.x);
        dmma884(dg0, dg1, bq.y, bq.y);
    }
    dg0 = live0 ? flip(dg0) : (8 * warp + g == c0 ? 1.0 : 0.0);
    dg1 = live1 ? flip(dg1) : (8 * warp + g == c0 + 1 ? 1.0 : 0.0);
    // Eight pivots, one dependent chain: broadcast of the pivot, rsqrt, scale, broadcast of the
    // column, fma.  Nothing else may sit on it: the LAPACK pivot test (d > 0) is only recorded - a
    // failed pivot poisons what follows with NaNs that nobody uses - and looked at after the
    // eighth pivot, and lane c keeps 1 / L_cc in a register until then (a store from inside the
    // chain made the compiler evaluate rsqrt twice per pivot, once for lane 0 alone).
    int bad = 0;
    double myrs = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double d = __shfl_sync(FULL, (c & 1) ? dg1 : dg0, 4 * c + (c >> 1));
        if (!DBG && bad == 0 && !(d > 0.0)) bad = c + 1;
        const double rs = rsqrt(d);
        if (lane == c) myrs = rs;
        double& col = (c & 1) ? dg1 : dg0;
        if (t == (c >> 1)) col = (g > c) ? col * rs : (g == c ? d * rs : 0.0);
        if (c < 7) {
            const double l0 = __shfl_sync(FULL, col, 8 * t + (c >> 1));
            const double l1 = __shfl_sync(FULL, col, 8 * t + 4 + (c >> 1));
            const double lr = __shfl_sync(FULL, col, 4 * g + (c >> 1));
            if (2 * t > c && live0) dg0 = fma(-lr, l0, dg0);
            if (2 * t + 1 > c && live1) dg1 = fma(-lr, l1, dg1);
        }
    }
    if (dead) {
    } else if (bad) {
        if (lane == 0) *failword = j0 + 8 * warp + bad;
        dead = true;
    } else if (lane < 8) {
        rsbuf[8 * warp + lane] = myrs;
    }
    if (!dead) {
        if (2 * t > g) dg0 = 0.0;           // upper part of the diagonal 8 x 8 block
        if (2 * t + 1 > g) dg1 = 0.0;
        *reinterpret_cast<double2*>(Ld + ldblk(warp, warp) + g * 8 + 2 * t) = make_double2(dg0, dg1);
    }
    __syncwarp();
    if (!dead && lane < 8) {
        // column `lane` of MINUS the inverse of the diagonal block (1 / L_ii is the pivot's rsqrt)
        const double* Lb = Ld + ldblk(warp, warp);
        double x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double sum = (i == lane) ? -1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) sum -= Lb[i * 8 + k] * x[k];
            x[i] = sum * rsbuf[8 * warp + i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) Dinv[warp * 64 + i * 8 + lane] = x[i];
    }
    __syncwarp();
    if (!dead) {
        const double2 dv = *reinterpret_cast<const double2*>(Dinv + warp * 64 + g * 8 + 2 * t);
        for (int bi = warp + 1; bi < 8; bi += 2) {
            const bool two = bi + 1 < 8;
            double a0, a1, b0 = 0.0, b1 = 0.0;
            load_neg(bi, a0, a1);
            if (two) load_neg(bi + 1, b0, b1);
            for (int q = 0; q < warp; ++q) {
                const double2 bq = *reinterpret_cast<const double2*>(Ld + ldblk(warp, q) + g * 8 + 2 * t);
                const double2 aq = *reinterpret_cast<const double2*>(Ld + ldblk(bi, q) + g * 8 + 2 * t);
                dmma884(a0, a1, aq.x, bq.x);
                dmma884(a0, a1, aq.y, bq.y);
                if (two) {
                    const double2 cq = *reinterpret_cast<const double2*>(Ld + ldblk(bi + 1, q) + g * 8 + 2 * t);
                    dmma884(b0, b1, cq.x, bq.x);
                    dmma884(b0, b1, cq.y, bq.y);
                }
            }
            if (!live0) a0 = b0 = 0.0;
            if (!live1) a1 = b1 = 0.0;
            // a = -(C - sum), dv = -inv(L_pp): x = a dv^T = (C - sum) L_pp^-T
            double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
            dmma884(x0, x1, a0, dv.x);
            dmma884(x0, x1, a1, dv.y);
            *reinterpret_cast<double2*>(Ld + ldblk(bi, warp) + g * 8 + 2 * t) = make_double2(x0, x1);
            if (two) {
                dmma884(y0, y1, b0, dv.x);
                dmma884(y0, y1, b1, dv.y);
                *reinterpret_cast<double2*>(Ld + ldblk(bi + 1, warp) + g * 8 + 2 * t) = make_double2(y0, y1);
            }
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(pbar + 8 * warp);       // panel `warp` is in Ld (or the block failed)
}

// lane 0 adds `inc` to a shared-memory counter without a branch; the old value is looked at later
__device__ __forceinline__ void atoms_add_lane0(unsigned& old, uint32_t addr, unsigned inc, int lane) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.eq.s32 p, %3, 0;\n\t"
        "@p atom.shared.add.u32 %0, [%1], %2;\n\t}"
        : "+r"(old)
        : "r"(addr), "r"(inc), "r"(lane)
        : "memory");
}
// The whole 64 x 64 diagonal block by ONE warp, in place in Ld (8 x 8 blocks holding MINUS C on
// entry, L_jj on exit) - potf2_panels with the warp index replaced by a loop over the panels and
// the mbarriers by __syncwarp; the same operations on the same values in the same order.
// Returns 0 or the 1-based global index of the first non-positive pivot.
template <int DBG>
__device__ __forceinline__ int potf2_one_warp(double* Ld, double* Dinv, double* rsbuf, int nv, int j0) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const unsigned FULL = 0xffffffffu;
    if (DBG & 32) return 0;
#pragma unroll 1
    for (int pn = 0; pn < 8; ++pn) {
        const int c0 = 8 * pn + 2 * t;
        const bool live0 = c0 < nv, live1 = c0 + 1 < nv;
        double2 v = *reinterpret_cast<const double2*>(Ld + ldblk(pn, pn) + g * 8 + 2 * t);
        double dg0 = v.x, dg1 = v.y;
        for (int q = 0; q < pn; ++q) {
            const double2 bq = *reinterpret_cast<const double2*>(Ld + ldblk(pn, q) + g * 8 + 2 * t);
            dmma884(dg0, dg1, bq.x, bq.x);
            dmma884(dg0, dg1, bq.y, bq.y);
        }
        dg0 = live0 ? flip(dg0) : (8 * pn + g == c0 ? 1.0 : 0.0);
        dg1 = live1 ? flip(dg1) : (8 * pn + g == c0 + 1 ? 1.0 : 0.0);
        int bad = 0;
        double myrs = 0.0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const double d = __shfl_sync(FULL, (c & 1) ? dg1 : dg0, 4 * c + (c >> 1));
            if (bad == 0 && !(d > 0.0)) bad = c + 1;           // LAPACK dpotrf pivot test
            const double rs = rsqrt(d);
            if (lane == c) myrs = rs;
            double& col = (c & 1) ? dg1 : dg0;
            if (t == (c >> 1)) col = (g > c) ? col * rs : (g == c ? d * rs : 0.0);
            if (c < 7) {
                const double l0 = __shfl_sync(FULL, col, 8 * t + (c >> 1));
                const double l1 = __shfl_sync(FULL, col, 8 * t + 4 + (c >> 1));
                const double lr = __shfl_sync(FULL, col, 4 * g + (c >> 1));
                if (2 * t > c && live0) dg0 = fma(-lr, l0, dg0);
                if (2 * t + 1 > c && live1) dg1 = fma(-lr, l1, dg1);
            }
        }
        if (bad) return j0 + 8 * pn + bad;
        if (lane < 8) rsbuf[8 * pn + lane] = myrs;
        if (2 * t > g) dg0 = 0.0;           // upper part of the diagonal 8 x 8 block
        if (2 * t + 1 > g) dg1 = 0.0;
        *reinterpret_cast<double2*>(Ld + ldblk(pn, pn) + g * 8 + 2 * t) = make_double2(dg0, dg1);
        __syncwarp();
        if (lane < 8) {
            // column `lane` of MINUS the inverse of the diagonal block (1 / L_ii is the pivot's rsqrt)
            const double* Lb = Ld + ldblk(pn, pn);
            double x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double sum = (i == lane) ? -1.0 : 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) sum -= Lb[i * 8 + k] * x[k];
                x[i] = sum * rsbuf[8 * pn + i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) Dinv[pn * 64 + i * 8 + lane] = x[i];
        }
        __syncwarp();
        const double2 dv = *reinterpret_cast<const double2*>(Dinv + pn * 64 + g * 8 + 2 * t);
        for (int bi = pn + 1; bi < 8; bi += 2) {
            const bool two = bi + 1 < 8;
            const double2 va = *reinterpret_cast<const double2*>(Ld + ldblk(bi, pn) + g * 8 + 2 * t);
            double a0 = va.x, a1 = va.y, b0 = 0.0, b1 = 0.0;
            if (two) {
                const double2 vb = *reinterpret_cast<const double2*>(Ld + ldblk(bi + 1, pn) + g * 8 + 2 * t);
                b0 = vb.x;
                b1 = vb.y;
            }
            for (int q = 0; q < pn; ++q) {
                const double2 bq = *reinterpret_cast<const double2*>(Ld + ldblk(pn, q) + g * 8 + 2 * t);
                const double2 aq = *reinterpret_cast<const double2*>(Ld + ldblk(bi, q) + g * 8 + 2 * t);
                dmma884(a0, a1, aq.x, bq.x);
                dmma884(a0, a1, aq.y, bq.y);
                if (two) {
                    const double2 cq = *reinterpret_cast<const double2*>(Ld + ldblk(bi + 1, q) + g * 8 + 2 * t);
                    dmma884(b0, b1, cq.x, bq.x);
                    dmma884(b0, b1, cq.y, bq.y);
                }
            }
            if (!live0) a0 = b0 = 0.0;
            if (!live1) a1 = b1 = 0.0;
            double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
            dmma884(x0, x1, a0, dv.x);
            dmma884(x0, x1, a1, dv.y);
            *reinterpret_cast<double2*>(Ld + ldblk(bi, pn) + g * 8 + 2 * t) = make_double2(x0, x1);
            if (two) {
                dmma884(y0, y1, b0, dv.x);
                dmma884(y0, y1, b1, dv.y);
                *reinterpret_cast<double2*>(Ld + ldblk(bi + 1, pn) + g * 8 + 2 * t) = make_double2(y0, y1);
            }
        }
        __syncwarp();
    }
    return 0;
}

// DBG (timing experiments, compiled only with -DMF_EXPERIMENTS; results are garbage): bit mask,
//   1024 = no ring traffic (no TMA loads, no waits: the loop reads whatever the stages hold),
//   4096 = TMA loads issued as usual but nobody waits for them,
//   16384 = info[] receives the index of the CTA that took the matrix (how the hand-out went),
//   8 = consume the ring without computing, 32 = skip the diagonal-block factorisation,
//   64 = skip the triangular solve, 128 = do not load the K tiles, 256 = stream the operands
//   of every matrix from the first four (L2-resident), 512 = do not store the rows of L
// COOP = 1: `p.group` CTAs (all co-resident: cooperative launch) factorise one matrix together,
// as a DATAFLOW over row tiles: tile `it` of block column jb goes to CTA (jb + it) mod G, every
// CTA walks its tiles in order, and there is no barrier between block columns.  A tile needs
// (i) its own rows and the block column's diagonal rows finished through the columns a chunk
// reads - the CTA that issues the chunk's TMA load first waits for `rowdone` of those 64-row
// blocks (one monotone word per row block: block columns finished, release/acquire; the check
// is a register compare once a row block is known to be complete) - and (ii) L_jj for its
// triangular solve, which is the diagonal tile's rowdone word.  CTAs that have no tile left in a
// block column run ahead into the next ones, up to the chunks that do depend on unfinished rows:
// the tail of a large matrix (fewer tiles than CTAs, long contractions) keeps all SMs busy.
// Per-element arithmetic is the one-CTA-per-matrix order (the factor is bit-identical); the
// minus-inverses of L_jj's diagonal blocks are recomputed by every CTA from L_jj with the same
// routine.  One group barrier per MATRIX.  Few large matrices (n = 30000, B = 1 or 64) or a small
// ensemble then fill the GPU instead of one SM per matrix.
// LA = 1 (one CTA per matrix): LOOK-AHEAD on the diagonal block, described for the 128-row tiles
// (NMI = 2; on 64-row tiles tile 0 IS the diagonal block, nothing is parked, tile 1 has 56 rows).  Without it the
// 64 x 64 diagonal block of every block column is a serial phase of the whole CTA: eight warps take
// turns on a chain of 64 dependent pivots while their accumulators wait (ncu: 8 % of all warp time is
// spent waiting inside that phase, and the co-resident CTA alone cannot keep the tensor pipe full
// meanwhile).  With it, tile 0 is contracted as before (diagonal block + the 64 rows below it); then
// the diagonal block goes to shared memory in the block layout and ONE warp (warp 7) factorises it in
// place - the same arithmetic, panel by panel (potf2_one_warp).  Warps 0-6 PARK their fragment of the
// 64 lower rows (MINUS C, not yet solved) in global memory, in the very place its X will be written -
// 32 KB per block column that nobody else reads - and go on to tile 1 with all their accumulators:
// 112 rows, seven warps x two fragments 56 rows apart, through the same TMA ring.  They meet warp 7
// again at tile 1's triangular solve, which needs L_jj (mbarrier bar_ljj), then read their parked
// fragment back and finish it; warp 7 finishes the one it kept in registers.  From tile 2 on all
// eight warps work on 128-row tiles.  A chunk of tile 1 is released by seven warps; warp 6 counts
// twice, so "the eighth release refills the stage" holds for every tile.  Warp 7 steps its ring
// position over tile 1's chunks, but only once warps 0-6 are through them (bar_t1): a ring wait names
// its phase by parity alone, and a warp two phases away from the others would read the wrong one.
// Per-element arithmetic is unchanged: factors and log-likelihoods are bit-identical (tested).
// Measured (2368 matrices of order 1500): 86.5 ms against 88.1 ms; the diagonal-block phase is gone
// from the profile (barrier stalls 5.3 -> 0.9 %).  Two earlier designs lost: the diagonal block as
// its own 64-row tile (91.8 ms: per-chunk costs against half the tensor work), and the fragment
// parked in registers with a 56-row single-fragment tile 1 (88.0 ms: the half tile is as inefficient
// as the phase it hides is long).
template <int NSTAGE, int MINB, int DBG, int COOP, int NMI, int LA = 0>
__global__ void __launch_bounds__(THREADS, MINB) mf_potrf_ll_kernel(const __grid_constant__ Params p) {
    static_assert(!LA || !COOP, "look-ahead: one CTA per matrix");
    constexpr int T1 = 56 * NMI;               // rows of the look-ahead tile of warps 0-6 (7 x NMI fragments)
    // NMI 8-row fragments per warp: tiles of 64 * NMI rows (NMI = 1: 80 registers, three CTAs per SM)
    constexpr int TM = 64 * NMI;
    constexpr int STAGE_BYTES = (TM + NB) * KC * 8;
    constexpr int OFF_LD = NSTAGE * STAGE_BYTES;
    constexpr int OFF_DINV = OFF_LD + SZ_LD;
    constexpr int OFF_COL = OFF_DINV + SZ_DINV;
    constexpr int OFF_RED = OFF_COL + SZ_COL;
    // 128B-swizzled TMA boxes need 1 KB alignment.  Declared, not computed: the kernel has no static
    // shared memory, so the array sits at a link-time constant offset of the shared window (aligning a
    // generic pointer by hand cost a dozen integer instructions per chunk).
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* const smem = smem_raw;
    double* const Aw = reinterpret_cast<double*>(smem);                 // aliases the stages
    double* const Ld = reinterpret_cast<double*>(smem + OFF_LD);
    double* const Dinv = reinterpret_cast<double*>(smem + OFF_DINV);
    double* const colbuf = reinterpret_cast<double*>(smem + OFF_COL);
    double* const red = reinterpret_cast<double*>(smem + OFF_RED);
    const uint32_t stage_u32 = smem_u32(smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int n = p.n;
    const int64_t ld = p.ld;

    // A stage holds 192 rows (128 of the row tile, 64 of the block column's own rows) of 16
    // doubles, 128 bytes each.  The tensor map delivers the rows of every 8-row group in the
    // order 0,2,4,6,1,3,5,7 and the hardware 128B swizzle XORs the 16-byte chunk index with the
    // row position mod 8, so the 16-byte fragment loads of a quarter warp (lane groups g, g+1:
    // positions 4 apart, chunk index t) hit eight different 16-byte bank groups: conflict-free.
    const int gp = (g >> 1) | ((g & 1) << 2);      // position of row g inside its 8-row group
    // potf2 ownership: rows tr + 16a, columns tc + 16b (a, b = 0..3) live in registers
    const int tr = tid & 15, tc = tid >> 4;
    // 8-row block of this warp inside each 64-row half tile.  The warp index goes through a warp
    // reduction: unchanged, but ptxas then knows it is the same in every lane and keeps it - and the
    // tile bookkeeping derived from it - in uniform registers (no spills, uniform branches)
    const int warp_u = __reduce_max_sync(0xffffffffu, warp);
    const int tbw = warp_u < 4 ? warp_u : 11 - warp_u;

    // Operand ring over the stages, no CTA-wide barrier and no producer wait in the
    // contraction loop:
    //   full[s]     - mbarrier: one arrival (the issuer's expect_tx) + the 24,576 bytes the TMA
    //                 delivers; every warp waits on it before reading stage s;
    //   released[s] - plain counter: a warp adds one once it has consumed its fragments of
    //                 stage s; the warp that brings it to a multiple of 8 knows the stage is free
    //                 and refills it itself, NSTAGE chunks ahead.
    // Chunks are numbered over the whole kernel lifetime: chunk c lives in stage c % NSTAGE and
    // is that stage's (c / NSTAGE)-th use (use_s, use_k: the next chunk this warp consumes).
    const uint32_t bar_full = smem_u32(smem + OFF_RED + 256);
    int* const released = reinterpret_cast<int*>(smem + OFF_RED + 416);
    if (tid == 0) {
        for (int s0 = 0; s0 < NSTAGE; ++s0) {
            mbar_init(bar_full + 8 * s0, 1);
            released[s0] = 0;
        }
        // diagonal-block factorisation: one mbarrier per 8-column panel, the failed-pivot word
        for (int q = 0; q < 8; ++q) mbar_init(smem_u32(smem + OFF_RED + 320) + 8 * q, 1);
        *reinterpret_cast<int*>(smem + OFF_RED + 384) = 0;
        if (LA) {
            mbar_init(smem_u32(smem + OFF_RED + 448), 1);       // L_jj is factorised (warp 7 arrives)
            mbar_init(smem_u32(smem + OFF_RED + 456), 7);       // warps 0-6 are through tile 1's contraction
        }
    }
    const uint32_t bar_ljj = smem_u32(smem + OFF_RED + 448), bar_t1 = smem_u32(smem + OFF_RED + 456);
    volatile int* const la_failword = reinterpret_cast<volatile int*>(smem + OFF_RED + 384);
    uint32_t ljj_par = 0, t1_par = 0;
    __syncthreads();
    uint32_t use_s = 0, use_k = 0;
    uint32_t potf2_parity = 0;

    const int G = COOP ? p.group : 1;
    const int grp = COOP ? (int)blockIdx.x / G : (int)blockIdx.x;
    const int rank = COOP ? (int)blockIdx.x % G : 0;
    const int ngroups = COOP ? (int)gridDim.x / G : (int)gridDim.x;
    unsigned* const gsync = COOP ? p.sync + 8 * grp : nullptr;
    unsigned* const grow = COOP ? p.rowdone + (size_t)grp * p.rd_stride : nullptr;
    // block columns known to be finished for the rows of the current tile / for the diagonal rows
    // of the current block column (CTA-wide cache of rowdone, so that most chunk issues cost a
    // shared-memory compare) and the level this thread has fenced into the async proxy
    volatile int* const known = reinterpret_cast<volatile int*>(smem + OFF_RED + 400);
    unsigned bar_target = 0;          // arrivals that complete the next group barrier
    unsigned ordinal = 0;             // matrices this group has started before this one
    auto group_barrier = [&]() {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            atomicAdd(gsync, 1u);
            bar_target += (unsigned)G;
            while ((int)(ld_acquire(gsync) - bar_target) < 0) {
            }
        }
        __syncthreads();
    };

    // rows a tile has finished are announced (rowdone) at the next CTA-wide barrier after it
    int pend_rb = 0, pend_n = 0;
    unsigned pend_val = 0;
    auto publish_pending = [&]() {
        if (pend_n > 0 && tid == 0) {
            __threadfence();
            for (int k = 0; k < pend_n; ++k) st_release(grow + pend_rb + k, pend_val);
        }
        pend_n = 0;
    };

    // One CTA per matrix: the matrices are handed out through an atomic counter rather than by
    // striding, so a CTA that gets ahead of its neighbour on the SM simply takes more of them and
    // all CTAs finish together.  (Cooperative groups and the timing-experiment builds stride.)
    constexpr bool DYNAMIC = !COOP;
    int* const next_mat = reinterpret_cast<int*>(red + 16);
    int taken = 0;
    for (int mat = grp;; mat += ngroups, ++ordinal) {
        if (DYNAMIC) {
            if (p.quota > 0 && taken >= p.quota) break;       // its fair share is done (CTA-uniform)
            ++taken;
            if (tid == 0) *next_mat = atomicAdd(p.counter, 1);
            __syncthreads();
            mat = *next_mat;
            // the same value in every lane - and ptxas knows it (a warp reduction lands in a uniform
            // register), so the TMA coordinates below stay on the uniform datapath
            mat = __reduce_max_sync(0xffffffffu, mat);
        }
        if (mat >= p.batch) break;
        double* const Kb = p.K + (size_t)mat * p.stride;
        const double* const rb = p.resid ? p.resid + (size_t)mat * p.resid_stride : nullptr;
        int fail = 0;
        const int ncb = (n + NB - 1) / NB;
        {
            // the TMA boxes move whole 8-row groups: rows n+1 .. 8*ceil((n+1)/8)-1 must read as zero
            const int64_t pad = (mf_rows_of(n) - (n + 1)) * ld;
            double* zp = Kb + (size_t)(n + 1) * ld;
            for (int64_t k = tid; k < pad; k += THREADS) zp[k] = 0.0;
            fence_async_proxy();
        }

        for (int jb = 0; jb < ncb; ++jb) {
            // a failed pivot is learnt here by the CTAs that did not meet it (one thread looks, the
            // barrier below makes the answer the same for the whole CTA)
            if (COOP && tid == 0)
                known[3] = (jb > 0 && ld_acquire(gsync + 2) == ordinal + 1) ? (int)ld_acquire(gsync + 3) : 0;
            const int j0 = jb * NB;
            const int nv = min(NB, n - j0);
            // look-ahead: tile 0 = 128 rows, tile 1 = up to 112 rows (seven warps), then 128-row tiles
            const int la_below = n + 1 - (j0 + TM);
            const int ntile = LA ? 1 + (la_below > 0 ? 1 + (la_below > T1 ? (la_below - T1 + TM - 1) / TM : 0) : 0)
                                 : (n + 1 - j0 + TM - 1) / TM;
            const int nk = j0 / KC;
            const int nicols = (nv + 7) >> 3;     // 8-column groups that hold real columns
            // rows written by the slowest warp in the previous block column must be in
            // memory before anyone prefetches them (the first chunks are issued ahead of
            // the first barrier of the contraction loop); also fences the stage area
            __syncthreads();
            if (COOP && known[3]) fail = known[3];
            if (LA) fail = *la_failword;          // a failed pivot of the previous diagonal block (warp 7 met it)
            if (fail) break;
            // The rows of L this block column streams were written through the generic proxy by
            // all threads (ordered by the barrier above); a thread that issues TMA loads must
            // itself order them before its async-proxy reads, and any warp's lane 0 may be the
            // issuer.  One fence per block column is enough: nothing read inside a block column
            // is written inside it.  (Measured: with writer-side fences only, ~13 % of the
            // matrices of a full wave come out wrong.)
            fence_async_proxy();

            if (COOP) publish_pending();      // (the barrier above: the last tile of the previous block column is complete)
            // Which CTA takes which tile.  (it + jb) mod G keeps a 64-row block with one CTA for the
            // whole factorisation - and makes the CTAs that own the last row blocks of a large matrix
            // do twice the average work (sum of rb^2 over their row blocks).  (it - jb) mod G walks
            // the row blocks across the CTAs instead: the tiles of the late, long block columns rotate
            // over all CTAs, and the CTA that owns a diagonal tile has nothing else to do in the block
            // column before it once there are fewer tiles than CTAs.
            const int it_first = !COOP ? 0 : (p.map_back ? (rank + jb) % G : (rank - jb % G + G) % G);
            const unsigned seq0 = ordinal * (unsigned)(ncb + 2);      // rowdone = seq0 + finished block columns
            bool have_L = !COOP;
            if (COOP && tid == 0) known[1] = 0;                        // diagonal rows of this block column
            // rowdone of row block `rbk` has reached `cols` finished block columns (polled by ONE
            // thread; gives up when the matrix has failed - the chunk is then issued regardless,
            // the ring must stay in step)
            auto wait_rows = [&](int rbk, int cols) -> int {
                unsigned v;
                while ((int)((v = ld_acquire(grow + rbk)) - (seq0 + (unsigned)cols)) < 0)
                    if (ld_acquire(gsync + 2) == ordinal + 1) return cols;
                return (int)(v - seq0);
            };
            double acc[2][8][2];                   // (declared here: fragment 1 of tile 0 outlives the tile in LA)
            for (int it = it_first; it < ntile; it += G) {
                const bool half = LA && it == 1;                       // the tile of warps 0-6
                if (half && warp_u == 7) continue;                     // warp 7 is in the diagonal block
                const int i0 = LA ? (it == 0 ? j0 : (it == 1 ? j0 + TM : j0 + TM + T1 + TM * (it - 2))) : j0 + it * TM;
                // first row of the NEXT tile of the block column
                const int r_next = LA ? (it == 0 ? j0 + TM : (it == 1 ? j0 + TM + T1 : i0 + TM)) : i0 + G * TM;
                const int rb8 = half ? warp_u : tbw;                   // 8-row block of fragment 0 in this tile
                const int fs = half ? 56 : 64;                         // rows between the warp's two fragments
                // Rows of the tile are dealt to the warps in 8-row fragments: fragment mi of warp w
                // is rows 8 tb(w) + 64 mi .. + 7 with tb = 0,1,2,3,7,6,5,4.  Warps w and w + 4 share
                // an SM sub-partition, so every sub-partition gets the same tensor work both in the
                // diagonal tile (triangle blocks s and 7 - s: s + 1 and 8 - s column groups) and in
                // the ragged last tile of a block column (valid fragments spread round-robin).
                const int wrow0 = i0 + 8 * rb8;                // first row of fragment 0; fragment 1 is 64 below
                // 8-column groups this warp has to produce: none if its rows lie beyond the
                // residual row, only the lower triangle inside the diagonal block
                int nil0 = (wrow0 > n) ? 0 : nicols;
                const int nil1 = (NMI < 2 || wrow0 + fs > n) ? 0 : nicols;
                if (it == 0) nil0 = min(nil0, tbw + 1);
                int fenced = 0;                                        // per thread, per tile
                if (COOP) {
                    // every warp is done with the previous tile: its rows are final
                    __syncthreads();
                    publish_pending();
                    // the ring tail of the previous tile looked the rows of THIS tile up in slot 2
                    if (tid == 0) {
                        known[0] = (it == it_first) ? 0 : known[2];
                        known[2] = 0;
                    }
                    __syncthreads();
                    // rows this tile finishes: the row blocks below the diagonal block
                    pend_rb = (i0 >> 6) + (it == 0 ? 1 : 0);
                    pend_n = NMI - (it == 0 ? 1 : 0);
                    pend_val = seq0 + (unsigned)jb + 1u;
                }

                // put chunk kc of the row tile at r0 into stage st: three 64-row TMA boxes, issued by
                // one thread (rows beyond the allocation are zero-filled by the hardware)
                auto tma_chunk = [&](uint32_t st, int kc, int r0) {
                    const uint32_t bar = bar_full + 8 * st;
                    const uint32_t dst = stage_u32 + st * STAGE_BYTES;
                    if (COOP) {
                        // columns 16 kc .. of block column c: the tile's rows and the diagonal rows
                        // must be finished through c + 1 block columns
                        const int need = (kc * KC) / NB + 1;
                        const int slot = (r0 == i0) ? 0 : 2;          // this tile / the CTA's next tile
                        int have = min((int)known[slot], (int)known[1]);
                        if (have < need) {
                            int a = wait_rows(r0 >> 6, need);
                            if (NMI == 2 && r0 + 64 <= n) a = min(a, wait_rows((r0 >> 6) + 1, need));
                            const int b = wait_rows(jb, need);
                            known[slot] = a;
                            known[1] = b;
                            have = min(a, b);
                        }
                        if (fenced < need) {
                            fence_async_proxy();                       // acquired rows -> this thread's TMA reads
                            fenced = have;
                        }
                    }
                    if (DBG & 1024) return;          // timing experiment: no ring traffic at all
                    mbar_expect_tx(bar, STAGE_BYTES);
                    const int lmat = (DBG & 256) ? (mat & 3) : mat;
                    // the rows of the tile are streamed once, the block column's own rows are read
                    // again by every tile of the block column: tell L2 which to keep
                    tma_box(dst, &p.tmap, kc * KC, r0 >> 3, lmat, bar, l2_policy_evict_first());
                    if (NMI == 2) tma_box(dst + 8192, &p.tmap, kc * KC, (r0 >> 3) + 8, lmat, bar, l2_policy_evict_first());
                    tma_box(dst + 8192 * NMI, &p.tmap, kc * KC, j0 >> 3, lmat, bar, l2_policy_evict_last());
                };
                // The warp has consumed chunk kc (stage use_s).  The LAST of the eight warps to
                // say so refills the stage at once with the chunk NSTAGE further on - of this tile,
                // else of the next tile of the block column (not from the diagonal tile, whose
                // factorisation scratch aliases the stages).  Nobody ever waits for a free stage.
                auto release_stage = [&](int kc) {
                    if (!COOP) {
                        // lane 0 counts the warp out (predicated, no branch); the old value reaches
                        // every lane through a warp reduction, so the "last one out" test and the
                        // refill are warp-uniform code
                        __syncwarp();
                        unsigned old = 0;
                        const unsigned inc = (half && warp_u == 6) ? 2u : 1u;      // (also for warp 7, see LA)
                        atoms_add_lane0(old, smem_u32(released) + 4 * use_s, inc, lane);
                        old = __reduce_max_sync(0xffffffffu, old);
                        if (((old + inc) & 7u) == 0u) {
                            int tk = kc + NSTAGE, r0 = i0;
                            bool go = true;
                            if (tk >= nk) {
                                tk -= nk;
                                r0 = r_next;
                                go = (LA || it > 0) && (it + 1 < ntile);
                            }
                            if (go && !(DBG & 1024))
                                tma_chunk_elect(stage_u32 + use_s * STAGE_BYTES, NMI == 2 ? &p.tmap128 : &p.tmap, &p.tmap,
                                                tk * KC, r0 >> 3, j0 >> 3, (DBG & 256) ? (mat & 3) : mat,
                                                bar_full + 8 * use_s, STAGE_BYTES, 8192 * NMI);
                        }
                        return;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        const int old = atomicAdd(released + use_s, 1);
                        if ((old & 7) == 7) {
                            int tk = kc + NSTAGE, r0 = i0;
                            bool go = true;
                            if (tk >= nk) {
                                tk -= nk;
                                r0 = i0 + G * TM;
                                go = (it > 0) && (it + G < ntile);
                            }
                            if (go) tma_chunk(use_s, tk, r0);
                        }
                    }
                };

                // pipeline prologue of a block column: fill the whole ring (every later tile finds
                // its first chunks already issued by the tail of its predecessor, see below)
                if (it == it_first && tid == 0) {
#pragma unroll
                    for (int s0 = 0; s0 < NSTAGE; ++s0)
                        if (s0 < nk) tma_chunk((use_s + s0) % NSTAGE, s0, i0);
                }
                // pull the NEXT tile of K (same block column, else the head of the next one)
                // into L2 while this tile is being contracted: one 512-byte row per thread
                // (issued here, at the start; 4 - 16 chunks before the end of the contraction was
                // measured 0.7 % slower)
                auto prefetch_next_k = [&]() {
                    if (tid < TM) {
                        const bool same = it + G < ntile;
                        const int pj = same ? j0 : j0 + NB;
                        const int prow = (same ? r_next : j0 + NB) + tid;
                        if (prow < n && pj < n) {
                            const unsigned bytes = (unsigned)min(NB, (int)ld - pj) * 8u;
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Kb + (size_t)prow * ld + pj), "r"(bytes) : "memory");
                        }
                    }
                };
                // (cooperative groups only: with one CTA per matrix the lines were gone from L2 again by
                // the time the tile started and the prefetch only added traffic - 0.9 % slower, measured)
                if (COOP) prefetch_next_k();

                // accumulators start at -K(tile); the residual enters as row n
                const bool interior = (it > 0) && (i0 + (half ? T1 : TM) <= n) && (j0 + NB <= n);
                if ((DBG & 128) && interior) {
#pragma unroll
                    for (int mi = 0; mi < NMI; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = -1.0 - (double)lane;
                } else if (interior) {
#pragma unroll
                    for (int mi = 0; mi < NMI; ++mi) {
                        const double* src = Kb + (size_t)(wrow0 + fs * mi + g) * ld + j0 + 2 * t;
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const double2 v = *reinterpret_cast<const double2*>(src + 8 * ni);
                            acc[mi][ni][0] = flip(v.x);
                            acc[mi][ni][1] = flip(v.y);
                        }
                    }
                } else {
                    // Diagonal, ragged and last-column tiles: entries above the diagonal, beyond the
                    // matrix or in rows beyond the residual start at zero.  Every load is issued
                    // unconditionally (an entry that does not count reads a harmless address and is
                    // dropped by a select): guarded loads serialised 32 memory latencies per thread
                    // here, 2.2 % of all warp time (ncu).
#pragma unroll
                    for (int mi = 0; mi < NMI; ++mi) {
                        const int row = wrow0 + fs * mi + g;
                        const bool mrow = row < n, rrow = (row == n) && (rb != nullptr);
                        const double* const rowp = mrow ? Kb + (size_t)row * ld : (rrow ? rb : Kb);
                        double v[8][2];
                        bool ok[8][2];
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int col = j0 + 8 * ni + 2 * t + e;
                                ok[ni][e] = col < n && ((mrow && col <= row) || rrow);
                                v[ni][e] = rowp[ok[ni][e] ? col : 0];
                            }
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            acc[mi][ni][0] = flip(ok[ni][0] ? v[ni][0] : 0.0);
                            acc[mi][ni][1] = flip(ok[ni][1] ? v[ni][1] : 0.0);
                        }
                    }
                }

                // The triangular solve against L_jj and the store of the finished rows, for fragment 0 at
                // row r0 and fragment 1 at row r1 (a0 / a1: the fragment holds rows that need it).
                auto finish_rows = [&](int r0, bool a0, int r1, bool a1) {
                    const bool dead = LA && *la_failword != 0;     // the diagonal block failed: nothing to finish
                    const bool act[2] = {a0 && !dead, a1 && !dead};
                    if (!act[0] && !act[1]) return;

                // ---- X = C L_jj^-T on the tensor cores, 8-column blocks, in registers.
                // acc holds -C; Dinv holds -inv(L_bb), so x = acc Dinv^T = +X, and the later
                // column groups (still negated) absorb +X L^T ----
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    if (DBG & 64) break;
                    const double2 dv = *reinterpret_cast<const double2*>(Dinv + nb * 64 + g * 8 + 2 * t);
#pragma unroll
                    for (int mi = 0; mi < NMI; ++mi) {
                        if (!act[mi]) continue;
                        double x0 = 0.0, x1 = 0.0;
                        dmma884(x0, x1, acc[mi][nb][0], dv.x);
                        dmma884(x0, x1, acc[mi][nb][1], dv.y);
                        acc[mi][nb][0] = x0;
                        acc[mi][nb][1] = x1;
                    }
#pragma unroll
                    for (int nb2 = nb + 1; nb2 < 8; ++nb2) {
                        const double2 lv = *reinterpret_cast<const double2*>(Ld + ldblk(nb2, nb) + g * 8 + 2 * t);
#pragma unroll
                        for (int mi = 0; mi < NMI; ++mi) {
                            if (!act[mi]) continue;
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], acc[mi][nb][0], lv.x);
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], acc[mi][nb][1], lv.y);
                        }
                    }
                }

                // ---- rows of L (and z) are final: write them once ----
#pragma unroll
                for (int mi = 0; mi < NMI; ++mi) {
                    const int row = (mi ? r1 : r0) + g;
                    if ((DBG & 512) && row < n - 1) continue;
                    if (act[mi] && row <= n) {
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const int col = j0 + 8 * ni + 2 * t;
                            double* dst = Kb + (size_t)row * ld + col;
                            if (col + 1 < n)
                                *reinterpret_cast<double2*>(dst) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                            else if (col < n)
                                dst[0] = acc[mi][ni][0];
                        }
                    }
                }
                fence_async_proxy();              // X -> async proxy
                };
                // ---- contraction over the finished columns k < j0 (tensor cores) ----
                for (int kc = 0; kc < nk; ++kc) {
                    if (DBG & 8) {          // timing experiment: consume the ring without computing
                        mbar_wait(bar_full + 8 * use_s, use_k & 1);
                        release_stage(kc);
                        if (++use_s == NSTAGE) { use_s = 0; ++use_k; }
                        continue;
                    }
                    if (!(DBG & (1024 | 4096))) mbar_wait(bar_full + 8 * use_s, use_k & 1);      // chunk kc has landed
                    const unsigned char* sb = smem + use_s * STAGE_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double2 a[2];
                        const int choff = ((4 * h + t) ^ gp) << 4;
#pragma unroll
                        for (int mi = 0; mi < NMI; ++mi)
                            a[mi] = *reinterpret_cast<const double2*>(sb + (8 * rb8 + fs * mi + gp) * 128 + choff);
#pragma unroll
                        for (int nq = 0; nq < 2; ++nq) {
                            // warp-uniform block-level branches (not per-instruction predicates): they
                            // keep ptxas from hoisting the next group's loads into this group
                            double2 b[4];
                            const bool d0 = 4 * nq < nil0, d1 = 4 * nq < nil1;
                            if (d0 || d1) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    b[u] = *reinterpret_cast<const double2*>(sb + (TM + 8 * (4 * nq + u) + gp) * 128 + choff);
                            }
                            if (d0 && d1) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
#pragma unroll
                                    for (int mi = 0; mi < NMI; ++mi)
                                        dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].x, b[u].x);
                            } else if (d0) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) dmma884(acc[0][4 * nq + u][0], acc[0][4 * nq + u][1], a[0].x, b[u].x);
                            } else if (d1) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) dmma884(acc[1][4 * nq + u][0], acc[1][4 * nq + u][1], a[1].x, b[u].x);
                            }
                            if (h == 1 && nq == 1) {
                                // Hand the stage back.  The arrival must not overtake the fragment
                                // loads: it is placed after tensor-core instructions that consume every
                                // register those loads fill, so the loads have completed by then (an
                                // arrive right after the loads let the TMA overwrite a stage that was
                                // still being read).
                                release_stage(kc);
                            }
                            if (d0 && d1) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
#pragma unroll
                                    for (int mi = 0; mi < NMI; ++mi)
                                        dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].y, b[u].y);
                            } else if (d0) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) dmma884(acc[0][4 * nq + u][0], acc[0][4 * nq + u][1], a[0].y, b[u].y);
                            } else if (d1) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) dmma884(acc[1][4 * nq + u][0], acc[1][4 * nq + u][1], a[1].y, b[u].y);
                            }
                        }
                    }
                    if (++use_s == NSTAGE) {
                        use_s = 0;
                        ++use_k;
                    }
                }
                // acc now holds S - K = -C

                if (LA && it == 0) {
                    // ---- look-ahead: MINUS C, lower 8 x 8 blocks, straight into the layout the
                    // factorisation works in; warp 7 factorises, the others go on to tile 1 ----
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni)
                        if (ni <= tbw)
                            *reinterpret_cast<double2*>(Ld + ldblk(tbw, ni) + g * 8 + 2 * t) = make_double2(acc[0][ni][0], acc[0][ni][1]);
                    const int prow1 = wrow0 + 64;              // first row of fragment 1
                    if (NMI == 2 && ntile > 1 && warp_u != 7 && prow1 + g <= n) {
                        // Warps 0-6 need all their accumulators for the next tile: fragment 1 (MINUS C,
                        // not yet solved) is parked where its X will go; the same thread reads it back.
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const int col = j0 + 8 * ni + 2 * t;
                            double* dst = Kb + (size_t)(prow1 + g) * ld + col;
                            if (col + 1 < n)
                                *reinterpret_cast<double2*>(dst) = make_double2(acc[1][ni][0], acc[1][ni][1]);
                            else if (col < n)
                                dst[0] = acc[1][ni][0];
                        }
                    }
                    __syncthreads();
                    if (warp_u == 7) {
                        const int f7 = potf2_one_warp<DBG>(Ld, Dinv, colbuf, nv, j0);
                        if (f7) {
                            if (lane == 0) *la_failword = f7;
                        } else {
                            // the inverses the row tiles multiply by, then L_jj (and the z entries inside
                            // it, when the block column is the last one) back to global memory
                            __syncwarp();
                            dinv_blocks(Ld, Dinv, lane);
                            dinv_blocks(Ld, Dinv, lane + 32);
                            const int cp = 2 * lane, col = j0 + cp;
                            for (int rr = 0; rr < NB; ++rr) {
                                const int row = j0 + rr;
                                if (row <= n && col < n) {
                                    double2 v = make_double2(0.0, 0.0);
                                    if ((cp >> 3) <= (rr >> 3))
                                        v = *reinterpret_cast<const double2*>(Ld + ldblk(rr >> 3, cp >> 3) + (rr & 7) * 8 + (cp & 7));
                                    double* dst = Kb + (size_t)row * ld + col;
                                    if (col + 1 < n)
                                        *reinterpret_cast<double2*>(dst) = v;
                                    else
                                        dst[0] = v.x;
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_ljj);   // L_jj and the inverses are in place (or the block failed)
                        finish_rows(0, false, prow1, NMI == 2 && prow1 <= n);
                        if (ntile > 1) {                        // step over the chunks of tile 1
                            mbar_wait(bar_t1, t1_par);
                            const uint32_t tot = use_s + (uint32_t)nk;
                            use_k += tot / NSTAGE;
                            use_s = tot % NSTAGE;
                        }
                        continue;
                    }
                    if (ntile > 1) continue;        // tile 1 first; the parked fragment is finished after it
                    mbar_wait(bar_ljj, ljj_par);
                    finish_rows(0, false, prow1, NMI == 2 && prow1 <= n);
                    continue;
                }
                if (it == 0) {
                    // ---- diagonal block: Cholesky of C[0:64, 0:64] ----
                    // (a CTA that ran ahead of a failed block column must not factorise garbage)
                    if (COOP && tid == 0) known[3] = (ld_acquire(gsync + 2) == ordinal + 1) ? (int)ld_acquire(gsync + 3) : 0;
                    __syncthreads();              // every warp is done with the stages Aw aliases
                    if (COOP && known[3]) {
                        fail = known[3];
                        break;
                    }
                    // fragment 0 of every warp is its 8-row block of the diagonal block
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) {
                        double* d = Aw + (8 * tbw + g) * AW_LD + 8 * ni + 2 * t;
                        d[0] = flip(acc[0][ni][0]);
                        d[1] = flip(acc[0][ni][1]);
                    }
                    // these generic-proxy writes land in memory the TMA (async proxy) refills later
                    fence_async_proxy();
                    __syncthreads();
#ifndef MF_POTF2_OLD
                    potf2_panels<DBG>(Aw, Ld, Dinv, colbuf, smem_u32(smem + OFF_RED + 320),
                                      reinterpret_cast<volatile int*>(smem + OFF_RED + 384), nv, j0, potf2_parity);
                    volatile int* const failword = reinterpret_cast<volatile int*>(smem + OFF_RED + 384);
                    potf2_parity ^= 1u;
                    __syncthreads();
                    if (*failword) {
                        fail = *failword;
                        __syncthreads();
                        if (tid == 0) *failword = 0;
                    }
                    if (fail) {
                        if (COOP && tid == 0) {       // tell the group, then release whoever waits for L_jj
                            gsync[3] = (unsigned)fail;
                            __threadfence();
                            st_release(gsync + 2, ordinal + 1);
                        }
                        break;
                    }
#else
                    double e[4][4];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) e[a][b] = Aw[(tr + 16 * a) * AW_LD + tc + 16 * b];
#pragma unroll
                    for (int cb = 0; cb < 4; ++cb) {
                        for (int cl = 0; cl < 16; ++cl) {
                            const int c = 16 * cb + cl;
                            if (c >= nv || (DBG & 32)) break;
                            double* col = colbuf + (c & 1) * 72;
                            if (tc == cl) {
                                // the 16 threads that own column c sit in one half-warp; lane cl of
                                // it owns the pivot
                                const double d = __shfl_sync(0xffffu << (16 * (lane >> 4)), e[cb][cb], (lane & 16) + cl);
                                const bool bad = !DBG && !(d > 0.0);   // LAPACK dpotrf pivot test
                                const double rs = bad ? 0.0 : rsqrt(d);
                                if (tr == cl) col[64] = bad ? -1.0 : 1.0;
#pragma unroll
                                for (int a = 0; a < 4; ++a) {
                                    const int r = tr + 16 * a;
                                    const double v = (r > c) ? e[a][cb] * rs : (r == c ? d * rs : 0.0);
                                    col[r] = v;
                                    if ((r >> 3) >= (c >> 3)) Ld[ldblk(r >> 3, c >> 3) + (r & 7) * 8 + (c & 7)] = v;
                                }
                            }
                            __syncthreads();
                            if (col[64] < 0.0) {
                                fail = j0 + c + 1;
                                break;
                            }
                            double lr[4], lk[4];
#pragma unroll
                            for (int a = 0; a < 4; ++a) lr[a] = col[tr + 16 * a];
#pragma unroll
                            for (int b = 0; b < 4; ++b) lk[b] = col[tc + 16 * b];
#pragma unroll
                            for (int a = 0; a < 4; ++a)
#pragma unroll
                                for (int b = 0; b < 4; ++b)
                                    if (b >= cb && tc + 16 * b > c && tr + 16 * a >= tc + 16 * b)
                                        e[a][b] = fma(-lr[a], lk[b], e[a][b]);
                        }
                        if (fail) break;
                    }
                    if (fail) {
                        if (COOP && tid == 0) {       // tell the group, then release whoever waits for L_jj
                            gsync[3] = (unsigned)fail;
                            __threadfence();
                            st_release(gsync + 2, ordinal + 1);
                        }
                        break;
                    }
                    // columns beyond the matrix order: identity, keeps the 8x8 inverses finite
                    if (tid < 64)
                        for (int c = nv; c < NB; ++c)
                            if ((tid >> 3) >= (c >> 3))
                                Ld[ldblk(tid >> 3, c >> 3) + (tid & 7) * 8 + (c & 7)] = (tid == c) ? 1.0 : 0.0;
                    __syncthreads();
#endif
                    // the inverses the row tiles multiply by (the factorisation above kept its own,
                    // built from the pivots' reciprocal square roots, for the blocks inside L_jj)
                    dinv_blocks(Ld, Dinv, tid);
                    // L_jj (and, when the block column is the last one, the z entries that
                    // live inside it) go back to global memory, one 512-byte row per warp pass
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int idx = tid + THREADS * q;
                        const int rr = idx >> 5, cp = (idx & 31) * 2;
                        const int row = j0 + rr, col = j0 + cp;
                        if (row <= n && col < n) {
                            double2 v = make_double2(0.0, 0.0);
                            if ((cp >> 3) <= (rr >> 3))
                                v = *reinterpret_cast<const double2*>(Ld + ldblk(rr >> 3, cp >> 3) + (rr & 7) * 8 + (cp & 7));
                            double* dst = Kb + (size_t)row * ld + col;
                            if (col + 1 < n)
                                *reinterpret_cast<double2*>(dst) = v;
                            else
                                dst[0] = v.x;
                        }
                    }
                    fence_async_proxy();          // L_jj -> async proxy
                    __syncthreads();
                    if (COOP && tid == 0) {
                        // publish: the rows of L_jj are final (row block jb has jb + 1 block columns)
                        __threadfence();
                        st_release(grow + jb, seq0 + (unsigned)jb + 1u);
                    }
                    have_L = true;
                } else if (COOP && !have_L) {
                    // first tile of a CTA that does not own the diagonal block: wait for it
                    if (tid == 0) {
                        wait_rows(jb, jb + 1);
                        known[3] = (ld_acquire(gsync + 2) == ordinal + 1) ? (int)ld_acquire(gsync + 3) : 0;
                    }
                    __syncthreads();
                    if (known[3]) {
                        fail = known[3];
                        // the tail of this tile has already put the first chunks of this CTA's next
                        // tile into the ring: take them out again so the ring stays in step
                        if (it + G < ntile && nk > 0) {
                            for (int s0 = 0; s0 < NSTAGE; ++s0) {
                                mbar_wait(bar_full + 8 * use_s, use_k & 1);
                                __syncwarp();
                                if (lane == 0) atomicAdd(released + use_s, 1);
                                if (++use_s == NSTAGE) { use_s = 0; ++use_k; }
                            }
                        }
                        break;
                    }
                    for (int idx = tid; idx < 36 * 64; idx += THREADS) {
                        const int blk = idx >> 6, r8 = (idx >> 3) & 7, c8 = idx & 7;
                        int bi = 0;
                        while ((bi + 1) * (bi + 2) / 2 <= blk) ++bi;
                        const int bj = blk - bi * (bi + 1) / 2;
                        const int r = 8 * bi + r8, c = 8 * bj + c8;
                        Ld[idx] = (r < nv && c < nv) ? __ldcg(Kb + (size_t)(j0 + r) * ld + j0 + c)
                                                     : (r == c ? 1.0 : 0.0);
                    }
                    __syncthreads();
                    dinv_blocks(Ld, Dinv, tid);
                    __syncthreads();
                    have_L = true;
                }
                // the stage area is idle from here to the next contraction loop: start the next
                // tile's first chunks now so they land while the triangular solve runs
                // (the tail of every other tile has already refilled the ring for its successor)
                if (it == 0 && G < ntile && tid == 0) {
#pragma unroll
                    for (int s0 = 0; s0 < NSTAGE; ++s0)
                        if (s0 < nk) tma_chunk((use_s + s0) % NSTAGE, s0, i0 + G * TM);
                }
                if (half) {
                    // through tile 1's contraction: warp 7 may use the ring again; then wait for L_jj
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_t1);
                    mbar_wait(bar_ljj, ljj_par);
                    finish_rows(wrow0, wrow0 <= n, wrow0 + fs, NMI == 2 && wrow0 + fs <= n);
                    // then the fragment of tile 0 this warp parked in global memory (MINUS C, where X goes)
                    const int prow1 = j0 + 8 * tbw + 64;
                    if (NMI == 2 && prow1 <= n && !*la_failword) {
                        const double* src = Kb + (size_t)(prow1 + g) * ld + j0 + 2 * t;
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const int col = j0 + 8 * ni + 2 * t;
                            double2 v = make_double2(0.0, 0.0);
                            if (prow1 + g <= n) {
                                if (col + 1 < n)
                                    v = *reinterpret_cast<const double2*>(src + 8 * ni);
                                else if (col < n)
                                    v.x = src[8 * ni];
                            }
                            acc[1][ni][0] = v.x;
                            acc[1][ni][1] = v.y;
                        }
                        finish_rows(0, false, prow1, true);
                    }
                } else {
                    // fragments that still need the solve: real rows, not part of the diagonal block
                    finish_rows(wrow0, it > 0 && wrow0 <= n, wrow0 + 64, NMI == 2 && wrow0 + 64 <= n);
                }
            }
            if (LA) {
                ljj_par ^= 1u;
                if (ntile > 1) t1_par ^= 1u;
            }
        }

        // ---- -1/2 z.z - sum log L_ii - n/2 log(2 pi) ----
        // z (row n) and the diagonal of L are read back in one fixed order, whoever wrote them:
        // the result is bit-identical for one CTA per matrix and for any cooperative group size.
        if (COOP) {
            __syncthreads();
            publish_pending();
            group_barrier();          // the other CTAs' rows are complete and visible
            if (!fail && ld_acquire(gsync + 2) == ordinal + 1) fail = (int)ld_acquire(gsync + 3);
            if (rank != 0) continue;  // CTA 0 of the group finishes alone
        }
        __syncthreads();
        if (LA) fail = *la_failword;
        double q = 0.0, lg = 0.0;
        if (!fail) {
            const double* z = Kb + (size_t)n * ld;
            for (int c = tid; c < n; c += THREADS) {
                const double zc = COOP ? __ldcg(z + c) : z[c];
                const double dc = COOP ? __ldcg(Kb + (size_t)c * ld + c) : Kb[(size_t)c * ld + c];
                q += zc * zc;
                lg += log(dc);
            }
        }
        q = warp_sum(q);
        lg = warp_sum(lg);
        if (lane == 0) {
            red[warp] = q;
            red[8 + warp] = lg;
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0, logdet = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) {
                s += red[w];
                logdet += red[8 + w];
            }
            p.ll[mat] = fail ? -INFINITY : (-0.5 * s - logdet - 0.5 * (double)n * log(2.0 * 3.141592653589793));
            p.info[mat] = (DBG & 16384) ? (int)blockIdx.x : fail;      // (experiment: who took which matrix)
            if (LA) *la_failword = 0;
        }
        __syncthreads();
    }
}

}  // namespace pf

// Row-tile passes (one pass = every warp contracts one 8-row fragment over a chunk; a tile of more
// than half its height takes two) times chunks, summed over the block columns, for the plain tiling
// (tiles of `tm` rows from the diagonal block down) and for the look-ahead tiling (tile 0 of tm rows,
// tile 1 of 7/8 tm rows, then tm-row tiles): where the ragged last tile of a block column falls
// differs between the two, by up to 7 % of the work at some orders.
static void mf_tiling_cost(int n, int tm, double* plain, double* lookahead) {
    const int t1 = tm / 8 * 7;
    auto tail = [&](long rows) -> long {            // passes of `rows` rows cut into tm-row tiles
        if (rows <= 0) return 0;
        const long r = rows % tm;
        return 2 * (rows / tm) + (r == 0 ? 0 : (r <= tm / 2 ? 1 : 2));
    };
    double a = 0.0, b = 0.0;
    for (int j0 = 0; j0 < n; j0 += 64) {
        const long R = (long)n + 1 - j0, nk = j0 / 16;
        a += (double)nk * tail(R);
        long p = R <= tm / 2 ? 1 : 2;
        const long below = R - tm;
        if (below > 0) {
            const long r1 = below < t1 ? below : t1;
            p += r1 <= t1 / 2 ? 1 : 2;
            p += tail(below - t1);
        }
        b += (double)nk * p;
    }
    *plain = a;
    *lookahead = b;
}

extern "C" int mf_potrf_loglike(mf_handle_t h, int batch, int n, double* K, int64_t ld,
                                int64_t stride, const double* resid, int64_t resid_stride,
                                double* ll, int32_t* info, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!K || !ll || !info) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0 || n <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch and n must be positive");
    if (ld < n || (ld % 16) != 0) MF_FAIL(h, MF_ERR_INVALID, "ld must be a multiple of 16 and >= n");
    if (stride < mf_rows_of(n) * ld)
        MF_FAIL(h, MF_ERR_INVALID, "stride must cover 8 * ceil((n + 1) / 8) rows (mf_matrix_stride)");
    if (((uintptr_t)K & 15) != 0) MF_FAIL(h, MF_ERR_INVALID, "K must be 16-byte aligned");
    pf::Params p;
    {
        // tensor map of the TMA ring, cached per (workspace, shape) in the handle
        mf_tmap_key key = {1, K, ld, stride, n, batch};
        mf_tmap_slot* slot = nullptr;
        const CUtensorMap* tm = mf_tmap_find(h, key, &slot);
        if (!tm) {
            mf_tmap_encode_fn encode = mf_tmap_encoder(h);
            if (!encode) return MF_ERR_CUDA;
            const cuuint64_t gdim[5] = {(cuuint64_t)ld, 4, 2, (cuuint64_t)(mf_rows_of(n) / 8), (cuuint64_t)batch};
            const cuuint64_t gstr[4] = {(cuuint64_t)(2 * ld * 8), (cuuint64_t)(ld * 8), (cuuint64_t)(8 * ld * 8),
                                        (cuuint64_t)(stride * 8)};
            const cuuint32_t box[5] = {16, 4, 2, 8, 1};
            const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            CUresult cr = encode(&slot->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, K, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) MF_FAIL(h, MF_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)cr);
            slot->key.kind = key.kind;
            tm = &slot->map;
        }
        p.tmap = *tm;
        mf_tmap_key key2 = {3, K, ld, stride, n, batch};
        slot = nullptr;
        const CUtensorMap* tm2 = mf_tmap_find(h, key2, &slot);
        if (!tm2) {
            mf_tmap_encode_fn encode = mf_tmap_encoder(h);
            if (!encode) return MF_ERR_CUDA;
            const cuuint64_t gdim[5] = {(cuuint64_t)ld, 4, 2, (cuuint64_t)(mf_rows_of(n) / 8), (cuuint64_t)batch};
            const cuuint64_t gstr[4] = {(cuuint64_t)(2 * ld * 8), (cuuint64_t)(ld * 8), (cuuint64_t)(8 * ld * 8),
                                        (cuuint64_t)(stride * 8)};
            const cuuint32_t box[5] = {16, 4, 2, 16, 1};
            const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            CUresult cr = encode(&slot->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, K, gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) MF_FAIL(h, MF_ERR_CUDA, "cuTensorMapEncodeTiled (128-row boxes) failed with %d", (int)cr);
            slot->key.kind = key2.kind;
            tm2 = &slot->map;
        }
        p.tmap128 = *tm2;
    }
    p.batch = batch;
    p.n = n;
    p.K = K;
    p.ld = ld;
    p.stride = stride;
    p.resid = resid;
    p.resid_stride = resid_stride;
    p.ll = ll;
    p.info = info;
    p.sm_count = h->sm_count;
    const int slots = 2 * h->sm_count;      // two resident CTAs per SM (registers and shared memory)
    const int grid = batch < slots ? batch : slots;
    cudaStream_t st = (cudaStream_t)stream;
    p.group = 1;
    p.sync = nullptr;
    p.rowdone = nullptr;
    p.rd_stride = 0;
    p.map_back = h->opt.potrf_map == 2 ? 0 : 1;
    p.counter = reinterpret_cast<int*>(h->coop_sync);

    // Tile variant.  Small matrices: 64-row tiles (one 8-row fragment per warp, 80 registers) and
    // THREE CTAs per SM.  The 64 sequential pivots per block column are a latency chain that only
    // other resident matrices can hide, and small matrices are mostly that chain; measured (2664
    // matrices): n = 300 +23 %, n = 600 +13 %, n = 1100 +1.5 %, n = 1200 -4.6 %, n = 1500 -9 %.  MF_NMI = 1 | 2
    // forces a variant (tests).
    // Larger matrices in batches too small to fill the GPU (cooperative regime) also take the
    // 64-row tiles: twice the tiles to spread over the CTAs of a group (measured: n = 6000, B = 1
    // 26.1 -> 18.3 ms, n = 30000 643 -> 564 ms; break-even near batch x 128-row-tiles = 2.5 x 296).
    const int experiment = MF_EXPERIMENT_OF(h);
    const long tiles128 = (long)batch * ((n + 1 + 127) / 128);
    const bool small_tiles = (h->opt.potrf_tile ? h->opt.potrf_tile == 1 : (n <= 1100 || tiles128 < 5L * h->sm_count)) &&
                             experiment == 0;
    const int tile_rows = small_tiles ? 64 : 128;
    const int cta_slots = (small_tiles ? 3 : 2) * h->sm_count;
    const int tile_smem = small_tiles ? pf::smem_bytes(3, 1) : pf::smem_bytes(3);
    // The two CTAs of an SM do not share it fairly: the older one runs at nearly the speed it would
    // have alone (90 % of the SM at n = 3300), the younger one crawls until the older one has nothing
    // left to take - and then finishes its matrix alone.  With a handful of waves that tail is a sixth
    // of the run (592 matrices of order 3300: 264 ms instead of 217).  So with few waves a CTA may not
    // take more than its share, ceil(batch / CTAs): the old CTA leaves early and the young one gets the
    // SM; with many waves the tail is short and taking more than one's share is what balances the CTAs
    // (+4 % at the headline batch).
    {
        const int ctas = batch < cta_slots ? batch : cta_slots;
        const int share = (batch + ctas - 1) / ctas;
        p.quota = share <= 4 ? share : 0;
    }
    // look-ahead or plain kernel (one CTA per matrix): look-ahead hides the diagonal block's pivot chain
    // (2 - 10 % at the orders measured) unless its tiling wastes more than that on ragged tiles
    bool lookahead = h->opt.potrf_lookahead != 1;
    if (h->opt.potrf_lookahead == 0) {
        double cp = 0.0, cl = 0.0;
        mf_tiling_cost(n, tile_rows, &cp, &cl);
        // what hiding the pivot chain is worth shrinks with the order (measured: 11 % at n = 600, 2 % at 1500)
        const double worth = 60.0 / n < 0.12 ? 60.0 / n : 0.12;
        lookahead = cl <= (1.0 + worth) * cp;
    }
    const void* const coop_fn = small_tiles ? (const void*)pf::mf_potrf_ll_kernel<3, 3, 0, 1, 1>
                                            : (const void*)pf::mf_potrf_ll_kernel<3, 2, 0, 1, 2>;
    {
        // Fewer matrices than CTA slots: several CTAs per matrix, at most one per row tile of the
        // first block column.  MF_COOP_GROUP overrides the choice (1 = never; tests use it).
        // (for a handful of large matrices at most two CTAs per SM, also on the small tiles: with
        // three, one n = 30000 matrix takes 785 ms instead of 500; many small groups do better
        // with all three: 128 matrices of order 600, 97.5 k -> 111 k evaluations/s)
        const int coop_slots = batch >= 16 ? cta_slots : 2 * h->sm_count;
        const int ntile0 = (n + 1 + tile_rows - 1) / tile_rows;
        int G = coop_slots / batch;
        if (G > ntile0) G = ntile0;
        const bool forced = h->opt.potrf_group > 0;
        if (forced) G = h->opt.potrf_group;
        if (G > coop_slots) G = coop_slots;
        if (G > h->coop_groups) G = h->coop_groups;
        // One CTA per SM when that still gives every tile of the first block column its own CTA:
        // a CTA that has its SM to itself runs a tile ~1.75x faster than two sharing one, and
        // which CTAs share an SM cannot be chosen.  (With more tiles than SMs the shared mode
        // wins: n = 30000, one matrix, 148 CTAs 567 ms, 296 or 444 CTAs 500 ms.)
        const int per_sm1 = h->sm_count / batch;          // group size with one CTA per SM
        const bool alone = !forced && per_sm1 >= 2 && per_sm1 >= ntile0;
        if (alone) G = per_sm1 < ntile0 ? per_sm1 : ntile0;
        // one rowdone word per 64-row block (the residual row included) and group
        const int rd_stride = (n + 1 + 63) / 64 + 2;
        if (G >= 2 && experiment == 0 && rd_stride <= h->coop_words) {
            int ngroups = (alone ? h->sm_count : coop_slots) / G;
            if (ngroups > batch) ngroups = batch;
            if (ngroups > h->coop_groups) ngroups = h->coop_groups;
            if (ngroups > h->coop_words / rd_stride) ngroups = h->coop_words / rd_stride;
            // a dynamic shared-memory request above half an SM keeps a second CTA off the SM
            const int smem = alone ? 120 * 1024 : tile_smem;
            p.group = G;
            p.sync = h->coop_sync;
            p.rowdone = h->coop_sync + 8 * h->coop_groups;
            p.rd_stride = rd_stride;
            MF_CUDA_CHECK(h, cudaMemsetAsync(h->coop_sync, 0,
                                             ((size_t)8 * h->coop_groups + (size_t)ngroups * rd_stride) * sizeof(unsigned), st));
            MF_CUDA_CHECK(h, cudaFuncSetAttribute(coop_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
            void* args[] = {(void*)&p};
            const cudaError_t ce = cudaLaunchCooperativeKernel(coop_fn, dim3(ngroups * G), dim3(pf::THREADS), args,
                                                               smem, st);
            if (ce == cudaSuccess) return MF_OK;
            if (ce != cudaErrorCooperativeLaunchTooLarge) {
                snprintf(h->err, sizeof(h->err), "cooperative launch failed: %s", cudaGetErrorString(ce));
                return MF_ERR_CUDA;
            }
            // the CTAs cannot all be resident (SMs shared with another context): one CTA per matrix
            (void)cudaGetLastError();
            p.group = 1;
        }
    }
    MF_CUDA_CHECK(h, cudaMemsetAsync(h->coop_sync, 0, sizeof(int), st));      // p.counter
    // Batches that fill the GPU with 128-row tiles: the look-ahead instantiation (the diagonal block
    // of every block column is factorised by one warp behind the next tile's contraction).
    // potrf_lookahead = 1 keeps the plain kernel (tests compare the two bit for bit).
    if (!small_tiles && experiment == 0 && lookahead) {
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<3, 2, 0, 0, 2, 1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, pf::smem_bytes(3)));
        pf::mf_potrf_ll_kernel<3, 2, 0, 0, 2, 1><<<grid, pf::THREADS, pf::smem_bytes(3), st>>>(p);
        MF_CUDA_CHECK(h, cudaGetLastError());
        return MF_OK;
    }
    if (small_tiles && experiment == 0 && lookahead) {
        const int grid1 = batch < cta_slots ? batch : cta_slots;
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<3, 3, 0, 0, 1, 1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem));
        pf::mf_potrf_ll_kernel<3, 3, 0, 0, 1, 1><<<grid1, pf::THREADS, tile_smem, st>>>(p);
        MF_CUDA_CHECK(h, cudaGetLastError());
        return MF_OK;
    }
    if (small_tiles) {
        const int grid1 = batch < cta_slots ? batch : cta_slots;
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<3, 3, 0, 0, 1>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, tile_smem));
        pf::mf_potrf_ll_kernel<3, 3, 0, 0, 1><<<grid1, pf::THREADS, tile_smem, st>>>(p);
        MF_CUDA_CHECK(h, cudaGetLastError());
        return MF_OK;
    }
#define MF_LAUNCH_S(DBG_, GRID_, SMEM_)                                                            \
    do {                                                                                           \
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<3, 2, DBG_, 0, 2>,               \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_)); \
        pf::mf_potrf_ll_kernel<3, 2, DBG_, 0, 2><<<GRID_, pf::THREADS, SMEM_, st>>>(p);               \
    } while (0)
#define MF_LAUNCH(DBG_) MF_LAUNCH_S(DBG_, grid, pf::smem_bytes(3))
#ifdef MF_EXPERIMENTS
    // timing-experiment instantiations (scripts/ab_kernels.py); their results are not likelihoods.
    // Bit 2048 on top of any code: ONE CTA per SM (a shared-memory request above half an SM keeps
    // the second one off).
    {
        if (experiment & 8192) {
            // ... and a ring of FOUR stages instead of three (needs the whole SM's shared memory)
            const int g1 = batch < h->sm_count ? batch : h->sm_count;
            const int code = experiment & ~(2048 | 8192);
#define MF_LAUNCH_6(DBG_)                                                                                       \
    do {                                                                                                        \
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<4, 1, DBG_, 0, 2>,                          \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, pf::smem_bytes(4))); \
        pf::mf_potrf_ll_kernel<4, 1, DBG_, 0, 2><<<g1, pf::THREADS, pf::smem_bytes(4), st>>>(p);                 \
    } while (0)
            if (code == 0) MF_LAUNCH_6(0);
            else if (code == 96) MF_LAUNCH_6(96);
            else if (code == 8) MF_LAUNCH_6(8);
            else MF_FAIL(h, MF_ERR_INVALID, "unknown experiment %d", experiment);
#undef MF_LAUNCH_6
            MF_CUDA_CHECK(h, cudaGetLastError());
            return MF_OK;
        }
        const bool alone = (experiment & 2048) != 0;
        const int xgrid = alone ? (batch < h->sm_count ? batch : h->sm_count) : grid;
        const int xsmem = alone ? 120 * 1024 : pf::smem_bytes(3);
        switch (experiment & ~2048) {
            case 0: MF_LAUNCH_S(0, xgrid, xsmem); break;
            case 8: MF_LAUNCH_S(8, xgrid, xsmem); break;
            case 32: MF_LAUNCH_S(32, xgrid, xsmem); break;
            case 64: MF_LAUNCH_S(64, xgrid, xsmem); break;
            case 96: MF_LAUNCH_S(96, xgrid, xsmem); break;
            case 128: MF_LAUNCH_S(128, xgrid, xsmem); break;
            case 224: MF_LAUNCH_S(224, xgrid, xsmem); break;
            case 256: MF_LAUNCH_S(256, xgrid, xsmem); break;
            case 352: MF_LAUNCH_S(352, xgrid, xsmem); break;
            case 736: MF_LAUNCH_S(736, xgrid, xsmem); break;
            case 1760: MF_LAUNCH_S(1760, xgrid, xsmem); break;
            case 4832: MF_LAUNCH_S(4832, xgrid, xsmem); break;
            case 16384: MF_LAUNCH_S(16384, xgrid, xsmem); break;
            default: MF_FAIL(h, MF_ERR_INVALID, "unknown experiment %d", experiment);
        }
    }
#else
    MF_LAUNCH(0);
#endif
#undef MF_LAUNCH
#undef MF_LAUNCH_S
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
