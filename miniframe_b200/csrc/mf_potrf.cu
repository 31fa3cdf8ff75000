// Batched blocked Cholesky fused with forward substitution, log-determinant and
// the quadratic form: one launch replaces cho_factor + cho_solve + sum(log(diag))
// of the reference's log_likelihood (BIGgp.py:305-309, BIGgpPlusJitt.py:311-315,
// SMALLgp.py:269-273) for a whole batch of hyper-parameter sets.
//
// Design (B200: 148 SMs, FP64 DMMA, 126 MB L2, HBM3e):
//   * one matrix per CTA, two CTAs resident per SM, persistent over the batch:
//     no inter-CTA synchronisation anywhere; the serial diagonal-block phases of
//     one CTA are hidden behind the other CTA's tensor-core work;
//   * LEFT-looking by 64-wide block columns: tile C(128 x 64) = K - L_left L_top^T
//     is a dense contraction over the already-final columns, accumulated in
//     registers with mma.sync m8n8k4 f64 (SASS DMMA.8x8x4) from a cp.async
//     double-buffered, XOR-swizzled shared-memory pipeline.  Every tile of K is
//     read once and every tile of L written once (HBM traffic ~ n^3/(6*64)*8 B
//     per matrix, arithmetic intensity 16 flop/B);
//   * each warp owns 16 complete rows (16 x 64 accumulators), so the triangular
//     solve against the diagonal block runs entirely in registers on the tensor
//     cores: the C-fragment layout (row g, columns 2t, 2t+1) is reused directly as
//     the A operand with the k index permuted, the matching B operand being one
//     16-byte shared-memory load;
//   * the residual vector rides along as row n of the matrix: its row of L is
//     z = L^-1 r, so the forward substitution costs one extra row in the last row
//     tile, and -1/2 z.z - sum(log L_ii) - n/2 log(2 pi) falls out at the end.
#include <math.h>
#include <stdlib.h>

#include "mf_common.cuh"

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
constexpr int SZ_COL = (2 * 136 + 128) * 8;          // double-buffered pivot column (v, v/d, status), pivots, 1/sqrt
constexpr int SZ_RED = 64 * 8;                       // reduction scratch + 2 x NSTAGE mbarriers
constexpr int smem_bytes(int nstage) { return nstage * STAGE_BYTES + SZ_LD + SZ_DINV + SZ_COL + SZ_RED; }
static_assert(NB * AW_LD * 8 <= 2 * STAGE_BYTES, "potf2 staging array must fit in the stage area");
// 3 stages: 97,792 B, two CTAs per SM

struct Params {
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
    long long* trace;     // DBG & 16: per-tile phase timestamps of the first matrix of each CTA
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
// arrive on `bar` once all cp.async copies this thread has issued so far have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// DBG (timing experiments only): 8 = skip the contraction loop (results are garbage),
// 16 = record per-tile phase timestamps of the first matrix of every CTA
template <int NSTAGE, int MINB, int DBG, int VAR>
__global__ void __launch_bounds__(THREADS, MINB) mf_potrf_ll_kernel(const Params p) {
    constexpr int OFF_LD = NSTAGE * STAGE_BYTES;
    constexpr int OFF_DINV = OFF_LD + SZ_LD;
    constexpr int OFF_COL = OFF_DINV + SZ_DINV;
    constexpr int OFF_RED = OFF_COL + SZ_COL;
    extern __shared__ __align__(128) unsigned char smem[];
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

    // cp.async mapping: 8 consecutive threads move one 128-byte row of a stage
    const int lrow = tid >> 3, lch = tid & 7;
    const uint32_t ldst = stage_u32 + lrow * 128 + ((lch ^ ((lrow & 1) << 2)) << 4);
    // fragment address inside a stage: 16-byte chunk (4h + t) of row r, swizzled by row parity
    const int fswz = (g & 1) << 2;
    // potf2 ownership: rows tr + 16a, columns tc + 16b (a, b = 0..3) live in registers
    const int tr = tid & 15, tc = tid >> 4;

    // producer/consumer ring over the stages, no CTA-wide barrier in the contraction loop:
    //   full[s]  - 256 arrivals, one per thread, raised by the copy engine when that thread's
    //              cp.async copies into stage s have landed;
    //   empty[s] - 8 arrivals, one per warp, after the warp's last fragment load from stage s.
    // Chunks are numbered over the whole kernel lifetime (gfill issued, guse consumed):
    // chunk c lives in stage c % NSTAGE and is that stage's (c / NSTAGE)-th use.
    const uint32_t bar_full = smem_u32(smem + OFF_RED + 256);
    const uint32_t bar_empty = bar_full + 8 * NSTAGE;
    if (tid == 0) {
        for (int s0 = 0; s0 < NSTAGE; ++s0) {
            mbar_init(bar_full + 8 * s0, THREADS);
            mbar_init(bar_empty + 8 * s0, THREADS / 32);
        }
    }
    __syncthreads();
    uint32_t fill_s = 0, fill_k = 0;      // stage / use count of the next chunk to issue
    uint32_t use_s = 0, use_k = 0;        // ... of the next chunk to consume

    for (int mat = blockIdx.x; mat < p.batch; mat += gridDim.x) {
        double* const Kb = p.K + (size_t)mat * p.stride;
        const double* const rb = p.resid ? p.resid + (size_t)mat * p.resid_stride : nullptr;
        double logdet = 0.0;          // meaningful in warp 0 only
        int fail = 0;
        const int ncb = (n + NB - 1) / NB;

        for (int jb = 0; jb < ncb && !fail; ++jb) {
            const int j0 = jb * NB;
            const int nv = min(NB, n - j0);
            const int ntile = (n + 1 - j0 + TM - 1) / TM;
            const int nk = j0 / KC;
            const int nicols = (nv + 7) >> 3;     // 8-column groups that hold real columns
            // rows written by the slowest warp in the previous block column must be in
            // memory before anyone prefetches them (the first chunks are issued ahead of
            // the first barrier of the contraction loop); also fences the stage area
            __syncthreads();

            for (int it = 0; it < ntile; ++it) {
                const int i0 = j0 + it * TM;
                const int wrow0 = i0 + 16 * warp;              // first row this warp owns
                long long* tr_ = nullptr;
                if ((DBG & 16) && mat == (int)blockIdx.x && tid == 0) {
                    int tix = 0;
                    for (int q = 0; q < jb; ++q) tix += (n + 1 - q * NB + TM - 1) / TM;
                    tr_ = p.trace + ((size_t)blockIdx.x * 200 + tix + it) * 8;
                    unsigned smid;
                    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                    tr_[0] = smid; tr_[1] = jb * 1000 + it; tr_[2] = clock64();
                }
                // 8-column groups this warp has to produce: none if its rows lie beyond the
                // residual row, only the lower triangle inside the diagonal block
                int nilim = (wrow0 > n) ? 0 : nicols;
                if (it == 0 && warp < 4) nilim = min(nilim, 2 * warp + 2);

                auto issue_chunk = [&](int kc, int r0) {
                    // wait until every warp has released the previous use of this stage
                    if (fill_k > 0) mbar_wait(bar_empty + 8 * fill_s, (fill_k - 1) & 1);
                    const double* src_col = Kb + (size_t)kc * KC + lch * 2;
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int grow = (q < 4) ? (r0 + lrow + 32 * q) : (j0 + lrow + 32 * (q - 4));
                        const bool ok = grow <= n;
                        const double* src = ok ? src_col + (size_t)grow * ld : Kb;
                        cp_async16(ldst + fill_s * STAGE_BYTES + q * (32 * 128), src, ok ? 16 : 0);
                    }
                    cp_async_mbar_arrive(bar_full + 8 * fill_s);
                    if (++fill_s == NSTAGE) {
                        fill_s = 0;
                        ++fill_k;
                    }
                };

                // pipeline prologue: up to NSTAGE - 1 chunks in flight; for every tile but the
                // first of a block column it was issued before the previous tile's triangular solve
                if (it == 0) {
#pragma unroll
                    for (int s0 = 0; s0 < NSTAGE - 1; ++s0)
                        if (s0 < nk) issue_chunk(s0, i0);
                }
                // pull the NEXT tile of K (same block column, else the head of the next one)
                // into L2 while this tile is being contracted: one 512-byte row per thread
                if (tid < TM) {
                    const bool same = it + 1 < ntile;
                    const int pj = same ? j0 : j0 + NB;
                    const int prow = (same ? i0 + TM : j0 + NB) + tid;
                    if (prow < n && pj < n) {
                        const unsigned bytes = (unsigned)min(NB, (int)ld - pj) * 8u;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(Kb + (size_t)prow * ld + pj), "r"(bytes) : "memory");
                    }
                }

                // accumulators start at -K(tile); the residual enters as row n
                double acc[2][8][2];
                const bool interior = (it > 0) && (i0 + TM <= n) && (j0 + NB <= n);
                if (interior) {
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        const double* src = Kb + (size_t)(wrow0 + 8 * mi + g) * ld + j0 + 2 * t;
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const double2 v = *reinterpret_cast<const double2*>(src + 8 * ni);
                            acc[mi][ni][0] = flip(v.x);
                            acc[mi][ni][1] = flip(v.y);
                        }
                    }
                } else {
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
                        const int row = wrow0 + 8 * mi + g;
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const int col = j0 + 8 * ni + 2 * t;
                            double v0 = 0.0, v1 = 0.0;
                            if (row < n) {
                                if (col <= row && col < n) v0 = Kb[(size_t)row * ld + col];
                                if (col + 1 <= row && col + 1 < n) v1 = Kb[(size_t)row * ld + col + 1];
                            } else if (row == n && rb != nullptr) {
                                if (col < n) v0 = rb[col];
                                if (col + 1 < n) v1 = rb[col + 1];
                            }
                            acc[mi][ni][0] = flip(v0);
                            acc[mi][ni][1] = flip(v1);
                        }
                    }
                }

                if ((DBG & 16) && tr_) tr_[3] = clock64();
                // ---- contraction over the finished columns k < j0 (tensor cores) ----
                for (int kc = 0; kc < nk; ++kc) {
                    if (DBG & 8) {          // timing experiment: consume the ring without computing
                        mbar_wait(bar_full + 8 * use_s, use_k & 1);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_empty + 8 * use_s);
                        if (kc + NSTAGE - 1 < nk) issue_chunk(kc + NSTAGE - 1, i0);
                        if (++use_s == NSTAGE) { use_s = 0; ++use_k; }
                        continue;
                    }
                    mbar_wait(bar_full + 8 * use_s, use_k & 1);      // chunk kc has landed
                    const unsigned char* sb = smem + use_s * STAGE_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1) {
                            // refill the ring two chunks ahead; issued between the two halves so
                            // the tensor pipe already has work queued while the copies go out
                            if (kc + NSTAGE - 1 < nk) issue_chunk(kc + NSTAGE - 1, i0);
                        }
                        const int choff = ((4 * h + t) ^ fswz) << 4;
                        double2 a[2];
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi)
                            a[mi] = *reinterpret_cast<const double2*>(sb + (16 * warp + 8 * mi + g) * 128 + choff);
#pragma unroll
                        for (int nq = 0; nq < 2; ++nq) {
                            double2 b[4];
                            if (4 * nq < nilim) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    b[u] = *reinterpret_cast<const double2*>(sb + (TM + 8 * (4 * nq + u) + g) * 128 + choff);
                            }
                            if (h == 1 && nq == 1) {
                                // last fragment load of this chunk: hand the stage back
                                __syncwarp();
                                if (lane == 0) mbar_arrive(bar_empty + 8 * use_s);
                            }
                            if (4 * nq < nilim) {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
#pragma unroll
                                    for (int mi = 0; mi < 2; ++mi)
                                        dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].x, b[u].x);
#pragma unroll
                                for (int u = 0; u < 4; ++u)
#pragma unroll
                                    for (int mi = 0; mi < 2; ++mi)
                                        dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].y, b[u].y);
                            }
                        }
                    }
                    if (++use_s == NSTAGE) {
                        use_s = 0;
                        ++use_k;
                    }
                }
                // acc now holds S - K = -C
                if ((DBG & 16) && tr_) tr_[4] = clock64();

                if (it == 0) {
                    // ---- diagonal block: Cholesky of C[0:64, 0:64], one barrier per column ----
                    __syncthreads();              // every warp is done with the stages Aw aliases
                    if (warp < 4) {
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 8; ++ni) {
                                double* d = Aw + (16 * warp + 8 * mi + g) * AW_LD + 8 * ni + 2 * t;
                                d[0] = flip(acc[mi][ni][0]);
                                d[1] = flip(acc[mi][ni][1]);
                            }
                    }
                    __syncthreads();
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
                            double* col = colbuf + (c & 1) * 136;
                            if (tc == cl) {
                                // the 16 threads that own column c sit in one half-warp; lane cl of
                                // it owns the pivot
                                const double d = __shfl_sync(0xffffu << (16 * (lane >> 4)), e[cb][cb], (lane & 16) + cl);
                                const bool bad = !DBG && !(d > 0.0);   // LAPACK dpotrf pivot test
                                if (VAR == 0) {
                                    const double rs = bad ? 0.0 : rsqrt(d);
                                    if (tr == cl) col[128] = bad ? -1.0 : 1.0;
#pragma unroll
                                    for (int a = 0; a < 4; ++a) {
                                        const int r = tr + 16 * a;
                                        const double v = (r > c) ? e[a][cb] * rs : (r == c ? d * rs : 0.0);
                                        col[r] = v;
                                        col[64 + r] = v;
                                        if ((r >> 3) >= (c >> 3)) Ld[ldblk(r >> 3, c >> 3) + (r & 7) * 8 + (c & 7)] = v;
                                    }
                                } else {
                                    // L D L^T recurrence: only a reciprocal sits on the pivot chain, the
                                    // 1/sqrt(d) scaling of the columns is applied after the last pivot
                                    const double inv = bad ? 0.0 : __drcp_rn(d);
                                    if (tr == cl) {
                                        col[128] = bad ? -1.0 : 1.0;
                                        colbuf[272 + c] = d;
                                    }
#pragma unroll
                                    for (int a = 0; a < 4; ++a) {
                                        const int r = tr + 16 * a;
                                        const double v = (r > c) ? e[a][cb] : (r == c ? d : 0.0);
                                        col[r] = v;
                                        col[64 + r] = v * inv;
                                        if ((r >> 3) >= (c >> 3)) Ld[ldblk(r >> 3, c >> 3) + (r & 7) * 8 + (c & 7)] = v;
                                    }
                                }
                            }
                            __syncthreads();
                            if (col[128] < 0.0) {
                                fail = j0 + c + 1;
                                break;
                            }
                            double lr[4], lk[4];
#pragma unroll
                            for (int a = 0; a < 4; ++a) lr[a] = col[64 + tr + 16 * a];
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
                    if (fail) break;
                    if (VAR == 1) {
                        // scale column c by 1/sqrt(d_c): L = V D^-1/2, L_cc = d_c / sqrt(d_c)
                        __syncthreads();
                        if (tid < 64) colbuf[336 + tid] = (tid < nv) ? rsqrt(colbuf[272 + tid]) : 1.0;
                        __syncthreads();
                        for (int idx = tid; idx < 36 * 64; idx += THREADS) {
                            const int blkid = idx >> 6, rr = (idx >> 3) & 7, cc = idx & 7;
                            int bi = 0;
                            while ((bi + 1) * (bi + 2) / 2 <= blkid) ++bi;
                            const int bj = blkid - bi * (bi + 1) / 2;
                            const int c = 8 * bj + cc;
                            if (c < nv && 8 * bi + rr >= c) Ld[idx] *= colbuf[336 + c];
                        }
                    }
                    // columns beyond the matrix order: identity, keeps the 8x8 inverses finite
                    if (tid < 64)
                        for (int c = nv; c < NB; ++c)
                            if ((tid >> 3) >= (c >> 3))
                                Ld[ldblk(tid >> 3, c >> 3) + (tid & 7) * 8 + (c & 7)] = (tid == c) ? 1.0 : 0.0;
                    __syncthreads();
                    if (tid < 64) {
                        // column cc of MINUS the inverse of the 8x8 diagonal block bb of L_jj
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
                    if (warp == 0) {              // sum of log L_ii of this block
                        double lg = 0.0;
                        if (lane < nv) lg += log(Ld[ldblk(lane >> 3, lane >> 3) + (lane & 7) * 9]);
                        if (lane + 32 < nv) lg += log(Ld[ldblk((lane + 32) >> 3, (lane + 32) >> 3) + (lane & 7) * 9]);
                        logdet += warp_sum(lg);
                    }
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
                    __syncthreads();
                }
                // the stage area is idle from here to the next contraction loop: start the next
                // tile's first chunks now so they land while the triangular solve runs
                if (it + 1 < ntile) {
#pragma unroll
                    for (int s0 = 0; s0 < NSTAGE - 1; ++s0)
                        if (s0 < nk) issue_chunk(s0, i0 + TM);
                }
                if ((DBG & 16) && tr_) tr_[5] = clock64();
                if (it == 0 && warp < 4) continue;   // rows of the diagonal block are done
                if (wrow0 > n) continue;             // nothing real in this warp's rows

                // ---- X = C L_jj^-T on the tensor cores, 8-column blocks, in registers.
                // acc holds -C; Dinv holds -inv(L_bb), so x = acc Dinv^T = +X, and the later
                // column groups (still negated) absorb +X L^T ----
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
                    if (DBG & 64) break;
                    const double2 dv = *reinterpret_cast<const double2*>(Dinv + nb * 64 + g * 8 + 2 * t);
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) {
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
                        for (int mi = 0; mi < 2; ++mi) {
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], acc[mi][nb][0], lv.x);
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], acc[mi][nb][1], lv.y);
                        }
                    }
                }

                // ---- rows of L (and z) are final: write them once ----
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    const int row = wrow0 + 8 * mi + g;
                    if (row <= n && !(DBG & 128)) {
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
            }
        }

        // ---- -1/2 z.z - sum log L_ii - n/2 log(2 pi) ----
        __syncthreads();
        double q = 0.0;
        if (!fail) {
            const double* z = Kb + (size_t)n * ld;
            for (int c = tid; c < n; c += THREADS) q += z[c] * z[c];
        }
        q = warp_sum(q);
        if (lane == 0) red[warp] = q;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) s += red[w];
            p.ll[mat] = fail ? -INFINITY : (-0.5 * s - logdet - 0.5 * (double)n * log(2.0 * 3.141592653589793));
            p.info[mat] = fail;
        }
        __syncthreads();
    }
}

}  // namespace pf

extern "C" int mf_potrf_loglike(mf_handle_t h, int batch, int n, double* K, int64_t ld,
                                int64_t stride, const double* resid, int64_t resid_stride,
                                double* ll, int32_t* info, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!K || !ll || !info) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0 || n <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch and n must be positive");
    if (ld < n || (ld % 16) != 0) MF_FAIL(h, MF_ERR_INVALID, "ld must be a multiple of 16 and >= n");
    if (stride < (int64_t)(n + 1) * ld) MF_FAIL(h, MF_ERR_INVALID, "stride must cover n + 1 rows");
    if (((uintptr_t)K & 15) != 0) MF_FAIL(h, MF_ERR_INVALID, "K must be 16-byte aligned");
    pf::Params p;
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
    const char* es = getenv("MF_SLOTS_PER_SM");
    const char* en = getenv("MF_NSTAGE");
    const int per_sm = es ? atoi(es) : 2;
    const int nstage = en ? atoi(en) : 3;
    const int slots = per_sm * h->sm_count;
    const int grid = batch < slots ? batch : slots;
    cudaStream_t st = (cudaStream_t)stream;
#define MF_LAUNCH(NS, MB)                                                                          \
    do {                                                                                           \
        MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel<NS, MB, DBG_, VAR_>,                \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                              pf::smem_bytes(NS)));                                \
        pf::mf_potrf_ll_kernel<NS, MB, DBG_, VAR_><<<grid, pf::THREADS, pf::smem_bytes(NS), st>>>(p);    \
    } while (0)
#define VAR_ 0
    const char* ed = getenv("MF_DBG");
    const int dbg = ed ? atoi(ed) : 0;
    p.trace = nullptr;
    if (dbg == 16) {
        const size_t words = (size_t)grid * 200 * 8;
        MF_CUDA_CHECK(h, cudaMalloc(&p.trace, words * 8));
        MF_CUDA_CHECK(h, cudaMemsetAsync(p.trace, 0, words * 8, st));
#define DBG_ 16
        MF_LAUNCH(3, 2);
#undef DBG_
        MF_CUDA_CHECK(h, cudaStreamSynchronize(st));
        long long* host = (long long*)malloc(words * 8);
        MF_CUDA_CHECK(h, cudaMemcpy(host, p.trace, words * 8, cudaMemcpyDeviceToHost));
        const char* path = getenv("MF_TRACE");
        FILE* f = fopen(path ? path : "gpurun_out/potrf_trace.bin", "wb");
        if (f) { fwrite(host, 8, words, f); fclose(f); }
        free(host);
        cudaFree(p.trace);
        return MF_OK;
    }
    if (dbg) {
#define DBG_ 8
        if (dbg == 8) MF_LAUNCH(3, 2);
#undef DBG_
#define DBG_ 32
        if (dbg == 32) MF_LAUNCH(3, 2);
#undef DBG_
#define DBG_ 64
        if (dbg == 64) MF_LAUNCH(3, 2);
#undef DBG_
#define DBG_ 96
        if (dbg == 96) MF_LAUNCH(3, 2);
#undef DBG_
#define DBG_ 128
        if (dbg == 128) MF_LAUNCH(3, 2);
#undef DBG_
#define DBG_ 224
        if (dbg == 224) MF_LAUNCH(3, 2);
#undef DBG_
        MF_CUDA_CHECK(h, cudaGetLastError());
        return MF_OK;
    }
#define DBG_ 0
    const char* ev = getenv("MF_VARIANT");
    if (ev && atoi(ev) == 1 && per_sm >= 2) {
#undef VAR_
#define VAR_ 1
        MF_LAUNCH(3, 2);
#undef VAR_
#define VAR_ 0
    } else if (per_sm >= 2) {
        MF_LAUNCH(3, 2);
    } else {
        switch (nstage) {
            case 4: MF_LAUNCH(4, 1); break;
            case 6: MF_LAUNCH(6, 1); break;
            case 8: MF_LAUNCH(8, 1); break;
            default: MF_LAUNCH(3, 1); break;
        }
    }
#undef MF_LAUNCH
#undef DBG_
#undef VAR_
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
