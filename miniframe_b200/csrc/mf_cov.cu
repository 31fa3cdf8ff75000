// Fused covariance build (HBM-write bound): replaces compute_matrix() and the
// 17 (BIGgp.py:172-278) / 9-per-block (SMALLgp.py:139-249) separate N x N
// kernel-class evaluations of the reference.
//
//   * ONE sincos + ONE exp per UNORDERED pair of epochs {i, j}: the lag functions are
//     even or odd (or, for the quirky third derivative, kept for both signs), so the
//     value computed for (i, j) also gives (j, i).  A CTA owns 32 x 32 pair tiles with
//     tile_i >= tile_j; the base functions stay in registers for the direct entries
//     and cross one padded shared-memory tile for the transposed ones, so that every
//     global store is a fully coalesced 256-byte row segment;
//   * the base functions fan out to every block (P, Q) of the n x n matrix that needs
//     them (6 blocks + 3 transposed blocks for BIGgp in lower mode = 9 outputs per
//     transcendental evaluation; the reference does 17 evaluations for 9 blocks);
//   * the per-sample coefficient table (kernel constants, block coefficients) is
//     computed once per CTA and sample, in parallel over threads, and staged in shared
//     memory together with the time arrays; a CTA walks many tiles of one sample.
#include "mf_cov.cuh"

namespace {

constexpr int TT = 32;           // pair-tile edge
constexpr int THREADS = 256;     // 8 warps x 4 rows
constexpr int RPW = 4;
constexpr int TLD = 33;          // padded tile stride (doubles)

struct CovParams {
    mf_problem_t prob;
    int batch;
    int full;
    int tiles_per_dim;           // T = ceil(N / 32)
    int n_tiles;                 // T (T + 1) / 2
    const double* t;
    const double* diag_add;
    const double* hyper;
    const double* extra;
    double* K;
    int64_t ld;
    int64_t stride;
};

template <int KERNEL, bool HIGH>
__global__ void __launch_bounds__(THREADS)
mf_cov_build_kernel(const CovParams p) {
    extern __shared__ __align__(16) unsigned char cov_smem[];
    SampleCoefs& sc = *reinterpret_cast<SampleCoefs*>(cov_smem);
    constexpr int NBASE = HIGH ? 6 : 3;
    double* const tiles = reinterpret_cast<double*>(cov_smem + ((sizeof(SampleCoefs) + 15) / 16) * 16);
    double* const ti_s = tiles + NBASE * TT * TLD;
    double* const tj_s = ti_s + TT;

    const int N = p.prob.n_epochs;
    const int M = p.prob.n_series;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double nugget = p.prob.nugget;
    const bool has_extra = p.prob.framework == MF_FRAMEWORK_SMALLGP &&
                           p.prob.extra_kernel != MF_KERNEL_NONE && M > 1;
    const bool full = p.full != 0;

    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
        __syncthreads();                       // previous sample's readers are done with sc
        if (tid == 0)
            mf_make_coefs(p.prob, p.hyper + (size_t)b * p.prob.hyper_len,
                          p.extra ? p.extra + (size_t)b * p.prob.extra_len : nullptr, sc);
        double* const Kb = p.K + (size_t)b * p.stride;

        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            // triangular index -> (tile_i >= tile_j)
            int tile_i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((tile_i + 1) * (tile_i + 2) / 2 <= tile) ++tile_i;
            while (tile_i * (tile_i + 1) / 2 > tile) --tile_i;
            const int tile_j = tile - tile_i * (tile_i + 1) / 2;
            const int i0 = tile_i * TT, j0 = tile_j * TT;
            const bool diag_tile = tile_i == tile_j;

            __syncthreads();                   // tiles / times of the previous tile are consumed
            if (tid < TT) ti_s[tid] = (i0 + tid < N) ? p.t[i0 + tid] : 0.0;
            if (tid >= 32 && tid < 32 + TT) tj_s[tid - 32] = (j0 + tid - 32 < N) ? p.t[j0 + tid - 32] : 0.0;
            __syncthreads();

            // ---- direct pass: lane <-> j, rows 4 warp + rr <-> i; entries (P N + i, Q N + j)
            const int j = j0 + lane;
            const double tj = tj_s[lane];
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int il = warp * RPW + rr;
                const int i = i0 + il;
                Bases bs;
                const bool active = (i < N) && (j < N) && (!diag_tile || i >= j);
                if (active) {
                    const double r = ti_s[il] - tj;
                    mf_bases<KERNEL, HIGH>(r, sc.k, bs);
                    const double ez = has_extra ? mf_kernel_value(p.prob.extra_kernel, r, sc.e) : 0.0;
                    const bool same = (i == j);
                    for (int P = 0; P < M; ++P) {
                        const double dadd = same ? p.diag_add[P * N + i] : 0.0;
                        const int Qhi = full ? M : P + 1;
                        for (int Q = 0; Q < Qhi; ++Q)
                            Kb[(size_t)(P * N + i) * p.ld + (size_t)Q * N + j] =
                                mf_entry(sc, bs, P, Q, same, ez, dadd, nugget);
                    }
                } else {
                    bs.E = bs.AE = bs.Gdd = bs.Hf = bs.Hr = bs.G4 = 0.0;
                }
                double* tl = tiles + il * TLD + lane;
                tl[0] = bs.E;
                tl[TT * TLD] = bs.AE;
                tl[2 * TT * TLD] = bs.Gdd;
                if (HIGH) {
                    tl[3 * TT * TLD] = bs.Hf;
                    tl[4 * TT * TLD] = bs.Hr;
                    tl[5 * TT * TLD] = bs.G4;
                }
            }
            __syncthreads();

            // ---- transposed pass: lane <-> i (now the column), rows 4 warp + rr <-> j (now the
            // row); entries (P N + j, Q N + i) at lag t_j - t_i = -r: AE changes sign, H(r) and
            // H(-r) swap.  Only pairs with i > j (the diagonal was written above).
            const int ic = i0 + lane;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int jl = warp * RPW + rr;
                const int jr = j0 + jl;
                if (ic < N && jr < N && ic > jr) {
                    const double* tl = tiles + lane * TLD + jl;
                    Bases bs;
                    bs.E = tl[0];
                    bs.AE = -tl[TT * TLD];
                    bs.Gdd = tl[2 * TT * TLD];
                    if (HIGH) {
                        bs.Hf = tl[4 * TT * TLD];
                        bs.Hr = tl[3 * TT * TLD];
                        bs.G4 = tl[5 * TT * TLD];
                    } else {
                        bs.Hf = bs.Hr = bs.G4 = 0.0;
                    }
                    const double ez = has_extra ? mf_kernel_value(p.prob.extra_kernel, ti_s[lane] - tj_s[jl], sc.e) : 0.0;
                    for (int P = 0; P < M; ++P) {
                        // lower mode: the (j, i) entry lies above the diagonal of the n x n
                        // matrix unless the row block is below the column block
                        const int Qhi = full ? M : P;
                        for (int Q = 0; Q < Qhi; ++Q)
                            Kb[(size_t)(P * N + jr) * p.ld + (size_t)Q * N + ic] =
                                mf_entry(sc, bs, P, Q, false, ez, 0.0, nugget);
                    }
                }
            }
        }
    }
}


// ---------------------------------------------------------------------------------
// Rajpaul (BIGgp / BIGgpPlusJitt) likelihood path: three series, lower triangle only,
// no nugget (BIGgp.log_likelihood never applies one, BIGgp.py:300).  Same tiling and
// the same operation order as the generic kernel (terms with a zero coefficient are
// simply not issued, which leaves every bit unchanged), but the twelve non-zero block
// coefficients live in registers, the nine outputs per pair are produced in one go and
// only the three transposed VALUES cross shared memory.
// ---------------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(THREADS, 3)
mf_cov_build_big_lower_kernel(const CovParams p) {
    extern __shared__ __align__(16) unsigned char big_smem[];
    SampleCoefs& sc = *reinterpret_cast<SampleCoefs*>(big_smem);
    // two sets of three transposition tiles: one barrier per pair tile
    double* const tiles = reinterpret_cast<double*>(big_smem + ((sizeof(SampleCoefs) + 15) / 16) * 16);
    double* const t_s = tiles + 2 * 3 * TT * TLD;        // the whole time array, staged once

    const int N = p.prob.n_epochs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = p.ld;
    const int64_t o10 = (int64_t)N * ld, o20 = 2 * o10;          // row-block offsets
    const int64_t o11 = o10 + N, o21 = o20 + N, o22 = o20 + 2 * N;

    for (int k = tid; k < N; k += THREADS) t_s[k] = p.t[k];

    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
        __syncthreads();
        if (tid == 0)
            mf_make_coefs(p.prob, p.hyper + (size_t)b * p.prob.hyper_len,
                          p.extra ? p.extra + (size_t)b * p.prob.extra_len : nullptr, sc);
        __syncthreads();
        const double e00 = sc.c[0][0].cE, g00 = sc.c[0][0].cGdd;
        const double e11 = sc.c[1][1].cE;
        const double e22 = sc.c[2][2].cE, g22 = sc.c[2][2].cGdd;
        const double e10 = sc.c[1][0].cE, a10 = sc.c[1][0].cAE;
        const double e20 = sc.c[2][0].cE, g20 = sc.c[2][0].cGdd, a20 = sc.c[2][0].cAE;
        const double e21 = sc.c[2][1].cE, a21 = sc.c[2][1].cAE;
        double* const Kb = p.K + (size_t)b * p.stride;

        int prev_i0 = -1, prev_j0 = 0, par = 0;
        for (int tile = blockIdx.x;; tile += gridDim.x) {
            const bool have = tile < p.n_tiles;
            int i0 = 0, j0 = 0;
            bool diag_tile = false;
            if (have) {
                int tile_i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
                while ((tile_i + 1) * (tile_i + 2) / 2 <= tile) ++tile_i;
                while (tile_i * (tile_i + 1) / 2 > tile) --tile_i;
                const int tile_j = tile - tile_i * (tile_i + 1) / 2;
                i0 = tile_i * TT;
                j0 = tile_j * TT;
                diag_tile = tile_i == tile_j;
            }
            // ---- transposed stores of the PREVIOUS tile: lane <-> i (column), row <-> j
            if (prev_i0 >= 0) {
                const double* tb = tiles + (par ^ 1) * 3 * TT * TLD;
                const int ic = prev_i0 + lane;
#pragma unroll
                for (int rr = 0; rr < RPW; ++rr) {
                    const int jl = warp * RPW + rr;
                    const int jr = prev_j0 + jl;
                    if (ic < N && jr < N && ic > jr) {
                        const double* tl = tb + lane * TLD + jl;
                        double* dst = Kb + (size_t)jr * ld + ic;
                        dst[o10] = tl[0];
                        dst[o20] = tl[TT * TLD];
                        dst[o21] = tl[2 * TT * TLD];
                    }
                }
            }
            if (!have) break;
            // ---- direct pass of this tile: lane <-> j, rows 4 warp + rr <-> i
            double* const tw = tiles + par * 3 * TT * TLD;
            const int j = j0 + lane;
            const double tj = (j < N) ? t_s[j] : 0.0;
#pragma unroll
            for (int rr = 0; rr < RPW; ++rr) {
                const int il = warp * RPW + rr;
                const int i = i0 + il;
                double t10 = 0.0, t20 = 0.0, t21 = 0.0;
                if ((i < N) && (j < N) && (!diag_tile || i >= j)) {
                    Bases bs;
                    mf_bases<KERNEL, false>(t_s[i] - tj, sc.k, bs);
                    double* dst = Kb + (size_t)i * ld + j;
                    if (i != j) {
                        dst[0] = df(g00, bs.Gdd, dm(e00, bs.E));
                        dst[o11] = dm(e11, bs.E);
                        dst[o22] = df(g22, bs.Gdd, dm(e22, bs.E));
                        const double u10 = dm(e10, bs.E);
                        const double u20 = df(g20, bs.Gdd, dm(e20, bs.E));
                        const double u21 = dm(e21, bs.E);
                        dst[o10] = df(a10, bs.AE, u10);
                        dst[o20] = df(a20, bs.AE, u20);
                        dst[o21] = df(a21, bs.AE, u21);
                        t10 = df(a10, -bs.AE, u10);      // the same blocks at lag -r
                        t20 = df(a20, -bs.AE, u20);
                        t21 = df(a21, -bs.AE, u21);
                    } else {
                        // same epoch: white noise on every block, errors and jitter on the diagonal
                        for (int P = 0; P < 3; ++P)
                            for (int Q = 0; Q <= P; ++Q)
                                Kb[(size_t)(P * N + i) * ld + (size_t)Q * N + j] =
                                    mf_entry(sc, bs, P, Q, true, 0.0, p.diag_add[P * N + i], 0.0);
                    }
                }
                double* tl = tw + il * TLD + lane;
                tl[0] = t10;
                tl[TT * TLD] = t20;
                tl[2 * TT * TLD] = t21;
            }
            __syncthreads();          // tile buffer `par` complete; buffer `par ^ 1` fully read
            prev_i0 = i0;
            prev_j0 = j0;
            par ^= 1;
        }
    }
}


// ---------------------------------------------------------------------------------
// Rajpaul likelihood path in EPOCH-INTERLEAVED order (mode MF_COV_LOWER_INTERLEAVED): row and
// column r = 3 i + P instead of P N + i, a symmetric permutation of K that leaves the
// log-likelihood unchanged (the residual is permuted with it, mf_interleave_resid).  The nine
// values a pair of epochs (i > j) feeds - six blocks at lag t_i - t_j, three at t_j - t_i - then
// form one 3 x 3 micro-block, a 32 x 32 pair tile is a dense 96 x 96 tile of the matrix, and the
// nine output streams of the standard order (block offsets N * 8 bytes, not line-aligned, six
// megabytes apart) become ONE stream of 768-byte, 128-byte-aligned row segments.  Half a pair tile
// (16 x 32 pairs = 48 x 96 doubles) is staged in shared memory and written by a TMA bulk-tensor
// store (SASS UTMASTG) that overlaps the arithmetic of the next half; tiles on the diagonal are
// written row by row so that only the lower triangle is touched (algorithmic bytes = bytes
// written).  Same per-entry arithmetic as the standard-order kernels: the matrices are equal bit
// for bit up to the permutation.  The coefficient table of a sample is computed by 26 threads
// in parallel, one field each (a single thread needs ~3.6 us for its six divisions in sequence).
// ---------------------------------------------------------------------------------
constexpr int IL_TI = 16;                    // epochs i per half tile
constexpr int IL_ROWS = 3 * IL_TI;           // 48 matrix rows
constexpr int IL_COLS = 3 * TT;              // 96 matrix columns
constexpr int IL_BUF = IL_ROWS * IL_COLS;    // doubles per staging buffer (36,864 bytes)

struct CovILParams {
    alignas(64) CUtensorMap tmap;            // {column, row, matrix} over the chunk, box 96 x 48 x 1
    CovParams c;
};

__device__ __forceinline__ void tma_store_tile(const CUtensorMap* tm, uint32_t src, int c0, int r0, int mat) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(r0), "r"(mat), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int KERNEL>
__global__ void __launch_bounds__(THREADS, 3)
mf_cov_build_big_il_kernel(const __grid_constant__ CovILParams q) {
    const CovParams& p = q.c;
    extern __shared__ __align__(128) unsigned char il_smem_raw[];
    // the source of a bulk-tensor store must be 128-byte aligned: align by hand (128 spare bytes)
    unsigned char* const il_smem = il_smem_raw + ((128u - (smem_u32(il_smem_raw) & 127u)) & 127u);
    double* const stage = reinterpret_cast<double*>(il_smem);                    // 2 x [48][96]
    BigCoefs& sc = *reinterpret_cast<BigCoefs*>(il_smem + 2 * IL_BUF * sizeof(double));

    const int N = p.prob.n_epochs;
    const int n = 3 * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t ld = p.ld;
    const int npar = KERNEL == MF_KERNEL_QP ? 4 : 2;
    const bool jitt = p.prob.framework == MF_FRAMEWORK_BIGGP_JITT && p.extra != nullptr;
    unsigned half_count = 0;                 // halves this CTA has staged: buffer = half_count & 1

    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
        __syncthreads();                     // the previous sample's readers are done with sc
        {
            const double* hy = p.hyper + (size_t)b * p.prob.hyper_len;
            if (tid < MF_KC_FIELDS) {
                reinterpret_cast<double*>(&sc.k)[tid] = mf_kc_field(KERNEL, hy, tid);
                reinterpret_cast<double*>(&sc.e)[tid] = 0.0;
            } else if (tid >= 16 && tid < 22) {
                int P, Q;
                mf_big_pair_index(tid - 16, P, Q);
                const BlockCoef u = mf_big_pair_coef(tid - 16, hy + npar);
                mf_set_pair(sc, P, Q, u.cE, u.cAE, u.cGdd, u.cH, u.cG4, u.cW);
            } else if (tid >= 24 && tid < 27) {
                const int k = tid - 24;
                sc.amp2[k] = 0.0;
                sc.jit2[k] = jitt ? dm(p.extra[(size_t)b * p.prob.extra_len + k], p.extra[(size_t)b * p.prob.extra_len + k]) : 0.0;
            }
        }
        __syncthreads();
        const double e00 = sc.c[0][0].cE, g00 = sc.c[0][0].cGdd;
        const double e11 = sc.c[1][1].cE;
        const double e22 = sc.c[2][2].cE, g22 = sc.c[2][2].cGdd;
        const double e10 = sc.c[1][0].cE, a10 = sc.c[1][0].cAE;
        const double e20 = sc.c[2][0].cE, g20 = sc.c[2][0].cGdd, a20 = sc.c[2][0].cAE;
        const double e21 = sc.c[2][1].cE, a21 = sc.c[2][1].cAE;
        double* const Kb = p.K + (size_t)b * p.stride;

        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            int tile_i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((tile_i + 1) * (tile_i + 2) / 2 <= tile) ++tile_i;
            while (tile_i * (tile_i + 1) / 2 > tile) --tile_i;
            const int tile_j = tile - tile_i * (tile_i + 1) / 2;
            const int j0 = tile_j * TT;
            const bool diag_tile = tile_i == tile_j;
            const int j = j0 + lane;
            const double tj = (j < N) ? p.t[j] : 0.0;

            for (int half = 0; half < 2; ++half) {
                const int i0 = tile_i * TT + half * IL_TI;
                if (i0 >= N) break;
                double* const buf = stage + (half_count & 1u) * IL_BUF;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int il = warp + 8 * rr;
                    const int i = i0 + il;
                    if (i >= N || j >= N || (diag_tile && i < j)) continue;
                    Bases bs;
                    mf_bases<KERNEL, false>(p.t[i] - tj, sc.k, bs);
                    double* m = buf + (3 * il) * IL_COLS + 3 * lane;      // micro-block (3 i + P, 3 j + Q)
                    if (i != j) {
                        const double u10 = dm(e10, bs.E);
                        const double u20 = df(g20, bs.Gdd, dm(e20, bs.E));
                        const double u21 = dm(e21, bs.E);
                        m[0] = df(g00, bs.Gdd, dm(e00, bs.E));
                        m[1] = df(a10, -bs.AE, u10);                      // (0,1) = block (1,0) at lag -r
                        m[2] = df(a20, -bs.AE, u20);
                        m[IL_COLS] = df(a10, bs.AE, u10);
                        m[IL_COLS + 1] = dm(e11, bs.E);
                        m[IL_COLS + 2] = df(a21, -bs.AE, u21);
                        m[2 * IL_COLS] = df(a20, bs.AE, u20);
                        m[2 * IL_COLS + 1] = df(a21, bs.AE, u21);
                        m[2 * IL_COLS + 2] = df(g22, bs.Gdd, dm(e22, bs.E));
                    } else {
                        // same epoch: white noise on every block, errors and jitter on the diagonal
                        for (int P = 0; P < 3; ++P)
                            for (int Q = 0; Q <= P; ++Q)
                                m[P * IL_COLS + Q] = mf_entry(sc, bs, P, Q, true, 0.0, p.diag_add[P * N + i], 0.0);
                    }
                }
                // staged values -> async proxy; the store of the previous half (other buffer) must
                // have read its buffer before anyone refills it after the next barrier
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                if (!diag_tile) {
                    if (tid == 0) tma_store_tile(&q.tmap, smem_u32(buf), 3 * j0, 3 * i0, b);
                } else {
                    // rows of the lower triangle only: row R holds columns 3 j0 .. R
#pragma unroll 1
                    for (int rl = warp; rl < IL_ROWS; rl += THREADS / 32) {
                        const int R = 3 * i0 + rl;
                        if (R >= n) break;
                        const int ncols = R - 3 * j0 + 1;
                        const double* src = buf + rl * IL_COLS;
                        double* dst = Kb + (size_t)R * ld + 3 * j0;
                        for (int c = 2 * lane; c < ncols; c += 64) {
                            if (c + 1 < ncols)
                                *reinterpret_cast<double2*>(dst + c) = *reinterpret_cast<const double2*>(src + c);
                            else
                                dst[c] = src[c];
                        }
                    }
                }
                ++half_count;
            }
        }
    }
    // the last store still reads shared memory: it must have completed before the CTA retires
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// residual rows into the same order: dst[M i + P] = src[P N + i]
__global__ void mf_interleave_resid_kernel(int batch, int M, int N, const double* src, int64_t src_stride,
                                           double* dst, int64_t dst_stride) {
    const int n = M * N;
    for (int b = blockIdx.y; b < batch; b += gridDim.y) {
        const double* s = src + (size_t)b * src_stride;
        double* d = dst + (size_t)b * dst_stride;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
            d[r] = s[(r % M) * N + r / M];
    }
}

}  // namespace

// CTAs per sample (grid.x; the samples are grid.y, so the CTAs of one sample are scheduled
// together).  Enough CTAs to fill the GPU when there are few samples - and, with many samples,
// still ~1/32 of the resident CTAs per sample: the nine output blocks of a matrix lie megabytes
// apart, and with one CTA per sample 1184 matrices (10^4 pages of 2 MB) are written at once;
// keeping only ~32 matrices in flight makes the same stores 17 % faster (measured: 1 CTA per
// sample 3.99 TB/s, 34-46 CTAs 4.69 TB/s, 136 CTAs - one tile each - 4.04 TB/s).
static int cov_ctas_per_sample(mf_handle_t h, int samples, int n_tiles) {
    if (h->opt.cov_ctas > 0) return h->opt.cov_ctas < n_tiles ? h->opt.cov_ctas : n_tiles;
    const int resident = 8 * h->sm_count;
    int bx = (resident + samples - 1) / samples;
    const int spread = (resident + 31) / 32;
    if (bx < spread) bx = spread;
    if (bx > (n_tiles + 1) / 2) bx = (n_tiles + 1) / 2;      // at least two tiles per CTA
    if (bx < 1) bx = 1;
    return bx;
}

template <int KERNEL, bool HIGH>
static int launch_cov(mf_handle_t h, const CovParams& p, cudaStream_t st) {
    const size_t smem = ((sizeof(SampleCoefs) + 15) / 16) * 16 +
                        (size_t)((HIGH ? 6 : 3) * TT * TLD + 2 * TT) * sizeof(double);
    MF_CUDA_CHECK(h, cudaFuncSetAttribute(mf_cov_build_kernel<KERNEL, HIGH>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int by = p.batch < 65535 ? p.batch : 65535;
    const int bx = cov_ctas_per_sample(h, by, p.n_tiles);
    dim3 grid(bx, by);
    mf_cov_build_kernel<KERNEL, HIGH><<<grid, THREADS, smem, st>>>(p);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}

// interleaved order, BIGgp family: TMA-stored dense tiles
template <int KERNEL>
static int launch_cov_il(mf_handle_t h, const CovParams& p, cudaStream_t st) {
    CovILParams q;
    q.c = p;
    const int n = 3 * p.prob.n_epochs;
    mf_tmap_key key = {2, p.K, p.ld, p.stride, n, p.batch};
    mf_tmap_slot* slot = nullptr;
    const CUtensorMap* tm = mf_tmap_find(h, key, &slot);
    if (!tm) {
        mf_tmap_encode_fn encode = mf_tmap_encoder(h);
        if (!encode) return MF_ERR_CUDA;
        const cuuint64_t gdim[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)p.batch};
        const cuuint64_t gstr[2] = {(cuuint64_t)(p.ld * 8), (cuuint64_t)(p.stride * 8)};
        const cuuint32_t box[3] = {IL_COLS, IL_ROWS, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult cr = encode(&slot->map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p.K, gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) MF_FAIL(h, MF_ERR_CUDA, "cuTensorMapEncodeTiled (covariance tiles) failed with %d", (int)cr);
        slot->key.kind = key.kind;
        tm = &slot->map;
    }
    q.tmap = *tm;
    const size_t smem = 2 * IL_BUF * sizeof(double) + sizeof(BigCoefs) + 128;
    MF_CUDA_CHECK(h, cudaFuncSetAttribute(mf_cov_build_big_il_kernel<KERNEL>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int by = p.batch < 65535 ? p.batch : 65535;
    // CTAs per sample: about sixteen pair tiles per CTA, so that a CTA's coefficient set-up is
    // spread over many tiles while every sample still has enough CTAs in flight (measured, n = 1500,
    // 136 tiles: 4 / 8 / 17 / 34 / 68 / 136 CTAs per sample -> 5.32 / 5.37 / 5.29 / 5.13 / 4.75 /
    // 4.09 TB/s; n = 6000, 2016 tiles: 8 / 34 / 136 -> 5.04 / 5.21 / 5.51 TB/s); more when the
    // batch alone cannot fill the GPU
    int bx = h->opt.cov_ctas > 0 ? h->opt.cov_ctas : (p.n_tiles + 15) / 16;
    if (h->opt.cov_ctas == 0 && (long)bx * by < 3L * h->sm_count) bx = (3 * h->sm_count + by - 1) / by;
    if (bx > p.n_tiles) bx = p.n_tiles;
    mf_cov_build_big_il_kernel<KERNEL><<<dim3(bx, by), THREADS, smem, st>>>(q);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_interleave_resid(mf_handle_t h, int batch, int n_series, int n_epochs,
                                   const double* src, int64_t src_stride, double* dst,
                                   int64_t dst_stride, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!src || !dst) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0 || n_series < 1 || n_series > MF_MAX_SERIES || n_epochs < 1)
        MF_FAIL(h, MF_ERR_INVALID, "bad batch/n_series/n_epochs");
    const int n = n_series * n_epochs;
    const int bx = (n + 255) / 256 < 64 ? (n + 255) / 256 : 64;
    mf_interleave_resid_kernel<<<dim3(bx, batch < 65535 ? batch : 65535), 256, 0, (cudaStream_t)stream>>>(
        batch, n_series, n_epochs, src, src_stride, dst, dst_stride);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}

extern "C" int mf_build_cov(mf_handle_t h, const mf_problem_t* prob, int batch, const double* t,
                            const double* diag_add, const double* hyper, const double* extra,
                            int mode, double* K, int64_t ld, int64_t stride, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!prob || !t || !diag_add || !hyper || !K) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch must be positive");
    if (mode != MF_COV_LOWER && mode != MF_COV_FULL && mode != MF_COV_LOWER_INTERLEAVED)
        MF_FAIL(h, MF_ERR_INVALID, "unknown covariance mode %d", mode);
    const int full = mode == MF_COV_FULL;
    const int M = prob->n_series, N = prob->n_epochs;
    if (M < 1 || M > MF_MAX_SERIES || N < 1) MF_FAIL(h, MF_ERR_INVALID, "bad n_series/n_epochs");
    if (prob->kernel != MF_KERNEL_SE && prob->kernel != MF_KERNEL_QP)
        MF_FAIL(h, MF_ERR_INVALID, "unknown kernel id %d", prob->kernel);
    const int npar = prob->kernel == MF_KERNEL_QP ? 4 : 2;
    const int64_t n = (int64_t)M * N;
    if (ld < n || stride < n * ld) MF_FAIL(h, MF_ERR_INVALID, "ld/stride too small");
    if (prob->framework == MF_FRAMEWORK_SMALLGP) {
        if (prob->hyper_len != npar + 3 * M) MF_FAIL(h, MF_ERR_INVALID, "SMALLgp hyper_len");
        if (prob->extra_kernel != MF_KERNEL_NONE) {
            const int ne = prob->extra_kernel == MF_KERNEL_QP ? 4 : 2;
            if (prob->extra_kernel != MF_KERNEL_SE && prob->extra_kernel != MF_KERNEL_QP)
                MF_FAIL(h, MF_ERR_INVALID, "unknown extra kernel id");
            if (!extra || prob->extra_len < ne + M) MF_FAIL(h, MF_ERR_INVALID, "SMALLgp extra_len");
        }
    } else if (prob->framework == MF_FRAMEWORK_BIGGP ||
               prob->framework == MF_FRAMEWORK_BIGGP_JITT) {
        if (M != 3) MF_FAIL(h, MF_ERR_INVALID, "BIGgp has exactly 3 series");
        if (prob->hyper_len != npar + 5) MF_FAIL(h, MF_ERR_INVALID, "BIGgp hyper_len");
        if (prob->framework == MF_FRAMEWORK_BIGGP_JITT && (!extra || prob->extra_len < 3))
            MF_FAIL(h, MF_ERR_INVALID, "BIGgpPlusJitt needs 3 jitters");
    } else {
        MF_FAIL(h, MF_ERR_INVALID, "unknown framework id %d", prob->framework);
    }
    CovParams p;
    p.prob = *prob;
    p.batch = batch;
    p.full = full ? 1 : 0;
    p.tiles_per_dim = (N + TT - 1) / TT;
    p.n_tiles = p.tiles_per_dim * (p.tiles_per_dim + 1) / 2;
    p.t = t;
    p.diag_add = diag_add;
    p.hyper = hyper;
    p.extra = prob->extra_len > 0 ? extra : nullptr;
    p.K = K;
    p.ld = ld;
    p.stride = stride;
    cudaStream_t st = (cudaStream_t)stream;
    const bool high = prob->framework == MF_FRAMEWORK_SMALLGP;
    if (mode == MF_COV_LOWER_INTERLEAVED) {
        if (high || prob->nugget != 0.0)
            MF_FAIL(h, MF_ERR_UNSUPPORTED, "the interleaved order is built for BIGgp / BIGgpPlusJitt without nugget");
        if (((uintptr_t)K & 15) != 0 || (ld & 1) != 0 || (stride & 1) != 0)
            MF_FAIL(h, MF_ERR_INVALID, "K must be 16-byte aligned, ld and stride even");
        return prob->kernel == MF_KERNEL_QP ? launch_cov_il<MF_KERNEL_QP>(h, p, st)
                                             : launch_cov_il<MF_KERNEL_SE>(h, p, st);
    }
    if (!high && !p.full && prob->nugget == 0.0) {
        const int by = batch < 65535 ? batch : 65535;
        const int bx = cov_ctas_per_sample(h, by, p.n_tiles);
        dim3 grid(bx, by);
        const size_t smem = ((sizeof(SampleCoefs) + 15) / 16) * 16 +
                            (size_t)(2 * 3 * TT * TLD + N) * sizeof(double);
        if (smem > h->smem_optin) MF_FAIL(h, MF_ERR_UNSUPPORTED, "n_epochs %d too large for the staged time array", N);
        if (prob->kernel == MF_KERNEL_QP) {
            MF_CUDA_CHECK(h, cudaFuncSetAttribute(mf_cov_build_big_lower_kernel<MF_KERNEL_QP>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mf_cov_build_big_lower_kernel<MF_KERNEL_QP><<<grid, THREADS, smem, st>>>(p);
        } else {
            MF_CUDA_CHECK(h, cudaFuncSetAttribute(mf_cov_build_big_lower_kernel<MF_KERNEL_SE>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            mf_cov_build_big_lower_kernel<MF_KERNEL_SE><<<grid, THREADS, smem, st>>>(p);
        }
        MF_CUDA_CHECK(h, cudaGetLastError());
        return MF_OK;
    }
    if (prob->kernel == MF_KERNEL_QP)
        return high ? launch_cov<MF_KERNEL_QP, true>(h, p, st) : launch_cov<MF_KERNEL_QP, false>(h, p, st);
    return high ? launch_cov<MF_KERNEL_SE, true>(h, p, st) : launch_cov<MF_KERNEL_SE, false>(h, p, st);
}
