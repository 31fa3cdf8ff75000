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

#include "mf_common.cuh"

namespace pf {

constexpr int TM = 128;       // rows per tile (8 warps x 16)
constexpr int NB = 64;        // block-column width
constexpr int KC = 16;        // contraction chunk per pipeline stage
constexpr int NSTAGE = 2;
constexpr int THREADS = 256;
constexpr int STAGE_ROWS = TM + NB;
constexpr int STAGE_BYTES = STAGE_ROWS * KC * 8;     // 24576
constexpr int LD_LD = 72;     // doubles; 64 B mod 128 B -> conflict-free 16 B fragment loads
constexpr int AW_LD = 65;
constexpr int OFF_LD = NSTAGE * STAGE_BYTES;         // diagonal block L_jj
constexpr int OFF_DINV = OFF_LD + NB * LD_LD * 8;    // inverses of its eight 8x8 diagonal blocks
constexpr int OFF_RED = OFF_DINV + 8 * 64 * 8;
constexpr int SMEM_BYTES = OFF_RED + 64 * 8;
static_assert(NB * AW_LD * 8 <= NSTAGE * STAGE_BYTES, "potf2 work array must fit in the stage area");

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
};

__global__ void __launch_bounds__(THREADS, 2) mf_potrf_ll_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char smem[];
    double* const Aw = reinterpret_cast<double*>(smem);                 // aliases the stages
    double* const Ld = reinterpret_cast<double*>(smem + OFF_LD);
    double* const Dinv = reinterpret_cast<double*>(smem + OFF_DINV);
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

    for (int mat = blockIdx.x; mat < p.batch; mat += gridDim.x) {
        double* const Kb = p.K + (size_t)mat * p.stride;
        const double* const rb = p.resid ? p.resid + (size_t)mat * p.resid_stride : nullptr;
        double logdet = 0.0;
        int fail = 0;
        const int ncb = (n + NB - 1) / NB;

        for (int jb = 0; jb < ncb && !fail; ++jb) {
            const int j0 = jb * NB;
            const int nv = min(NB, n - j0);
            const int ntile = (n + 1 - j0 + TM - 1) / TM;
            const int nk = j0 / KC;
            // rows written by the slowest warp in the previous block column must be in
            // memory before anyone prefetches them (the first chunks are issued ahead of
            // the first barrier of the contraction loop)
            __syncthreads();

            for (int it = 0; it < ntile; ++it) {
                const int i0 = j0 + it * TM;

                auto load_stage = [&](int s, int kc) {
                    const double* src_col = Kb + (size_t)kc * KC + lch * 2;
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        const int grow = (q < 4) ? (i0 + lrow + 32 * q) : (j0 + lrow + 32 * (q - 4));
                        const bool ok = grow <= n;
                        const double* src = ok ? src_col + (size_t)grow * ld : Kb;
                        cp_async16(ldst + s * STAGE_BYTES + q * (32 * 128), src, ok ? 16 : 0);
                    }
                };

                if (nk > 0) {
                    load_stage(0, 0);
                    cp_async_commit();
                }

                // accumulators start at -K(tile); the residual enters as row n
                double acc[2][8][2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    const int row = i0 + 16 * warp + 8 * mi + g;
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) {
                        const int col = j0 + 8 * ni + 2 * t;
                        double v0 = 0.0, v1 = 0.0;
                        if (row < n) {
                            if (col + 1 <= row && col + 1 < n) {
                                const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)row * ld + col);
                                v0 = v.x;
                                v1 = v.y;
                            } else if (col <= row && col < n) {
                                v0 = Kb[(size_t)row * ld + col];
                            }
                        } else if (row == n && rb != nullptr) {
                            if (col < n) v0 = rb[col];
                            if (col + 1 < n) v1 = rb[col + 1];
                        }
                        acc[mi][ni][0] = -v0;
                        acc[mi][ni][1] = -v1;
                    }
                }

                // ---- contraction over the finished columns k < j0 (tensor cores) ----
                for (int kc = 0; kc < nk; ++kc) {
                    if (kc + 1 < nk) {
                        load_stage((kc + 1) & 1, kc + 1);
                        cp_async_commit();
                        cp_async_wait<1>();
                    } else {
                        cp_async_wait<0>();
                    }
                    __syncthreads();
                    const unsigned char* sb = smem + (kc & 1) * STAGE_BYTES;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int choff = ((4 * h + t) ^ fswz) << 4;
                        double2 a[2];
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi)
                            a[mi] = *reinterpret_cast<const double2*>(sb + (16 * warp + 8 * mi + g) * 128 + choff);
#pragma unroll
                        for (int ni = 0; ni < 8; ++ni) {
                            const double2 b = *reinterpret_cast<const double2*>(sb + (TM + 8 * ni + g) * 128 + choff);
#pragma unroll
                            for (int mi = 0; mi < 2; ++mi) {
                                dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].x, b.x);
                                dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi].y, b.y);
                            }
                        }
                    }
                    __syncthreads();
                }

                // acc = S - K  ->  C = K - S
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) {
                        acc[mi][ni][0] = -acc[mi][ni][0];
                        acc[mi][ni][1] = -acc[mi][ni][1];
                    }

                if (it == 0) {
                    // ---- diagonal block: unblocked Cholesky of C[0:64, 0:64] in shared memory ----
                    __syncthreads();
                    if (warp < 4) {
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 8; ++ni) {
                                double* d = Aw + (16 * warp + 8 * mi + g) * AW_LD + 8 * ni + 2 * t;
                                d[0] = acc[mi][ni][0];
                                d[1] = acc[mi][ni][1];
                            }
                    }
                    const int r = tid & 63, kg = tid >> 6;
                    for (int c = 0; c < nv; ++c) {
                        __syncthreads();
                        const double d = Aw[c * AW_LD + c];
                        if (!(d > 0.0)) {       // LAPACK dpotrf: info = index of the failing pivot
                            fail = j0 + c + 1;
                            break;
                        }
                        const double s = sqrt(d);
                        const double inv = 1.0 / s;
                        logdet += log(s);
                        const double lr = Aw[r * AW_LD + c] * inv;
                        if (kg == 0) Ld[r * LD_LD + c] = (r > c) ? lr : (r == c ? s : 0.0);
                        if (r > c)
                            for (int k = c + 1 + kg; k <= r; k += 4)
                                Aw[r * AW_LD + k] -= lr * (Aw[k * AW_LD + c] * inv);
                    }
                    if (fail) break;
                    for (int c = nv + kg; c < NB; c += 4) Ld[r * LD_LD + c] = (r == c) ? 1.0 : 0.0;
                    __syncthreads();
                    if (tid < 64) {
                        // column cc of the inverse of the 8x8 diagonal block bb of L_jj
                        const int bb = tid >> 3, cc = tid & 7;
                        const double* Lb = Ld + (8 * bb) * LD_LD + 8 * bb;
                        double x[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            double sum = (i == cc) ? 1.0 : 0.0;
#pragma unroll
                            for (int k = 0; k < i; ++k) sum -= Lb[i * LD_LD + k] * x[k];
                            x[i] = sum / Lb[i * LD_LD + i];
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) Dinv[bb * 64 + i * 8 + cc] = x[i];
                    }
                    // L_jj (and, when the block column is the last one, the z entries that
                    // live inside it) go back to global memory, one 512-byte row per warp pass
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int idx = tid + THREADS * q;
                        const int rr = idx >> 5, cp = (idx & 31) * 2;
                        const int row = j0 + rr, col = j0 + cp;
                        if (row <= n && col < n) {
                            double* dst = Kb + (size_t)row * ld + col;
                            if (col + 1 < n)
                                *reinterpret_cast<double2*>(dst) = make_double2(Ld[rr * LD_LD + cp], Ld[rr * LD_LD + cp + 1]);
                            else
                                dst[0] = Ld[rr * LD_LD + cp];
                        }
                    }
                    __syncthreads();
                    if (warp < 4) continue;     // rows of the diagonal block are done
                }

                // ---- X = C L_jj^-T on the tensor cores, 8-column blocks, in registers ----
#pragma unroll
                for (int nb = 0; nb < 8; ++nb) {
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
                        const double2 lv = *reinterpret_cast<const double2*>(Ld + (8 * nb2 + g) * LD_LD + 8 * nb + 2 * t);
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi) {
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], -acc[mi][nb][0], lv.x);
                            dmma884(acc[mi][nb2][0], acc[mi][nb2][1], -acc[mi][nb][1], lv.y);
                        }
                    }
                }

                // ---- rows of L (and z) are final: write them once ----
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) {
                    const int row = i0 + 16 * warp + 8 * mi + g;
                    if (row <= n) {
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
    MF_CUDA_CHECK(h, cudaFuncSetAttribute(pf::mf_potrf_ll_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          pf::SMEM_BYTES));
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
    const int slots = 2 * h->sm_count;
    const int grid = batch < slots ? batch : slots;
    pf::mf_potrf_ll_kernel<<<grid, pf::THREADS, pf::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
