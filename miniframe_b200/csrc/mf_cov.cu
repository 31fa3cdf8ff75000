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

}  // namespace

// CTAs per sample (grid.x; the samples are grid.y, so the CTAs of one sample are scheduled
// together).  Enough CTAs to fill the GPU when there are few samples - and, with many samples,
// still ~1/32 of the resident CTAs per sample: the nine output blocks of a matrix lie megabytes
// apart, and with one CTA per sample 1184 matrices (10^4 pages of 2 MB) are written at once;
// keeping only ~32 matrices in flight makes the same stores 17 % faster (measured: 1 CTA per
// sample 3.99 TB/s, 34-46 CTAs 4.69 TB/s, 136 CTAs - one tile each - 4.04 TB/s).
static int cov_ctas_per_sample(mf_handle_t h, int samples, int n_tiles) {
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

extern "C" int mf_build_cov(mf_handle_t h, const mf_problem_t* prob, int batch, const double* t,
                            const double* diag_add, const double* hyper, const double* extra,
                            int full, double* K, int64_t ld, int64_t stride, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!prob || !t || !diag_add || !hyper || !K) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch must be positive");
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
