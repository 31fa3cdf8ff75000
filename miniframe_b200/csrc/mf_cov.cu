// Fused covariance build (HBM-write bound): replaces compute_matrix() and the
// 17 (BIGgp.py:172-278) / 9-per-block (SMALLgp.py:139-249) separate N x N
// kernel-class evaluations of the reference with ONE sincos + ONE exp per
// (i, j) lag, kept in registers and fanned out to every block (P, Q) that needs
// it, written with coalesced 16-byte stores straight into the layout the
// factorisation reads.  Times and the per-sample coefficient table are staged in
// shared memory.
#include "mf_cov.cuh"

namespace {

constexpr int TI = 32;          // epoch rows per CTA tile
constexpr int TJ = 64;          // epoch columns per CTA tile (2 per lane)
constexpr int ROWS_PER_WARP = 4;
constexpr int THREADS = 256;

struct CovParams {
    mf_problem_t prob;
    int batch;
    int full;
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
    __shared__ SampleCoefs sc;
    __shared__ double ti_s[TI];
    __shared__ double tj_s[TJ];

    const int N = p.prob.n_epochs;
    const int M = p.prob.n_series;
    const int tiles_j = (N + TJ - 1) / TJ;
    const int tile_i = blockIdx.x / tiles_j;
    const int tile_j = blockIdx.x % tiles_j;
    const int i0 = tile_i * TI, j0 = tile_j * TJ;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double nugget = p.prob.nugget;
    const bool has_extra = p.prob.framework == MF_FRAMEWORK_SMALLGP &&
                           p.prob.extra_kernel != MF_KERNEL_NONE && M > 1;
    const bool vec_base = ((N & 1) == 0) && ((p.ld & 1) == 0);

    if (tid < TI) ti_s[tid] = (i0 + tid < N) ? p.t[i0 + tid] : 0.0;
    if (tid >= 64 && tid < 64 + TJ) tj_s[tid - 64] = (j0 + tid - 64 < N) ? p.t[j0 + tid - 64] : 0.0;

    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
        __syncthreads();   // previous sample's readers are done with sc
        if (tid == 0)
            mf_make_coefs(p.prob, p.hyper + (size_t)b * p.prob.hyper_len,
                          p.extra ? p.extra + (size_t)b * p.prob.extra_len : nullptr, sc);
        __syncthreads();
        double* Kb = p.K + (size_t)b * p.stride;

        const int jl = 2 * lane;
        const int j = j0 + jl;
        const bool j1ok = (j + 1 < N);
        const double tj0 = tj_s[jl], tj1 = tj_s[jl + 1];

#pragma unroll
        for (int rr = 0; rr < ROWS_PER_WARP; ++rr) {
            const int il = warp * ROWS_PER_WARP + rr;
            const int i = i0 + il;
            if (i >= N || j >= N) break;
            const double ti = ti_s[il];
            Bases b0, b1;
            mf_bases<KERNEL, HIGH>(ti - tj0, sc.k, b0);
            mf_bases<KERNEL, HIGH>(ti - tj1, sc.k, b1);
            double ez0 = 0.0, ez1 = 0.0;
            if (has_extra) {
                ez0 = mf_kernel_value(p.prob.extra_kernel, ti - tj0, sc.e);
                ez1 = mf_kernel_value(p.prob.extra_kernel, ti - tj1, sc.e);
            }
            for (int P = 0; P < M; ++P) {
                const int Qend = p.full ? M : P + 1;
                const double dadd = p.diag_add[P * N + i];
                for (int Q = 0; Q < Qend; ++Q) {
                    // lower mode keeps (P > Q) blocks whole and the i >= j half of (P == P)
                    const bool keep0 = p.full || P != Q || i >= j;
                    const bool keep1 = j1ok && (p.full || P != Q || i >= j + 1);
                    if (!keep0 && !keep1) continue;
                    const double v0 = mf_entry(sc, b0, P, Q, i == j, ez0, dadd, nugget);
                    const double v1 = mf_entry(sc, b1, P, Q, i == j + 1, ez1, dadd, nugget);
                    double* dst = Kb + (size_t)(P * N + i) * p.ld + (size_t)Q * N + j;
                    if (keep0 && keep1 && vec_base) {
                        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                    } else {
                        if (keep0) dst[0] = v0;
                        if (keep1) dst[1] = v1;
                    }
                }
            }
        }
    }
}

}  // namespace

template <int KERNEL, bool HIGH>
static int launch_cov(mf_handle_t h, const CovParams& p, cudaStream_t st) {
    const int N = p.prob.n_epochs;
    const int tiles = ((N + TI - 1) / TI) * ((N + TJ - 1) / TJ);
    dim3 grid(tiles, p.batch < 65535 ? p.batch : 65535);
    mf_cov_build_kernel<KERNEL, HIGH><<<grid, THREADS, 0, st>>>(p);
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
    p.t = t;
    p.diag_add = diag_add;
    p.hyper = hyper;
    p.extra = prob->extra_len > 0 ? extra : nullptr;
    p.K = K;
    p.ld = ld;
    p.stride = stride;
    cudaStream_t st = (cudaStream_t)stream;
    const bool high = prob->framework == MF_FRAMEWORK_SMALLGP;
    if (prob->kernel == MF_KERNEL_QP)
        return high ? launch_cov<MF_KERNEL_QP, true>(h, p, st) : launch_cov<MF_KERNEL_QP, false>(h, p, st);
    return high ? launch_cov<MF_KERNEL_SE, true>(h, p, st) : launch_cov<MF_KERNEL_SE, false>(h, p, st);
}
