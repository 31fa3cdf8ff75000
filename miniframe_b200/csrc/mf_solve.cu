// Stand-alone fused forward substitution + log-determinant on an existing lower
// factor: z = L^-1 r, ll = -1/2 z.z - sum(log L_ii) - n/2 log(2 pi).  This is the
// cho_solve / log-diag tail of the reference's log_likelihood (BIGgp.py:307-309)
// for callers that keep one factor and change the residual; the reference's
// dpotrs does two triangular solves, one is enough for the quadratic form.
// HBM-read bound: the lower triangle of L (4 n^2 bytes) is streamed once.
#include <math.h>

#include "mf_common.cuh"

namespace sv {

constexpr int NB = 64;
constexpr int THREADS = 256;
constexpr int LDS = 65;

struct Params {
    int batch;
    int n;
    const double* L;
    int64_t ld;
    int64_t stride;
    const double* resid;
    int64_t resid_stride;
    double* z;
    double* ll;
    const int32_t* info;
};

__global__ void __launch_bounds__(THREADS) mf_solve_logdet_kernel(const Params p) {
    __shared__ double Ls[NB * LDS];
    __shared__ double s_s[NB];
    __shared__ double red[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;

    for (int mat = blockIdx.x; mat < p.batch; mat += gridDim.x) {
        if (p.info != nullptr && p.info[mat] != 0) {
            if (tid == 0) p.ll[mat] = -INFINITY;
            continue;
        }
        const double* Lb = p.L + (size_t)mat * p.stride;
        const double* rb = p.resid + (size_t)mat * p.resid_stride;
        double* zb = p.z + (size_t)mat * n;
        double logdet = 0.0;   // accumulated by warp 0, lane-uniform

        for (int j0 = 0; j0 < n; j0 += NB) {
            const int nv = min(NB, n - j0);
            __syncthreads();   // z[0:j0) written, Ls / s_s free
            // diagonal block to shared memory (coalesced rows)
            for (int idx = tid; idx < NB * NB; idx += THREADS) {
                const int r = idx >> 6, c = idx & 63;
                Ls[r * LDS + c] = (r < nv && c <= r) ? Lb[(size_t)(j0 + r) * p.ld + j0 + c] : 0.0;
            }
            // s = r_block - L[block, 0:j0) z[0:j0): one warp per row, lanes stride the columns
            for (int r = warp; r < nv; r += THREADS / 32) {
                const double* Lr = Lb + (size_t)(j0 + r) * p.ld;
                double acc = 0.0;
                for (int k = lane; k < j0; k += 32) acc += Lr[k] * zb[k];
                acc = warp_sum(acc);
                if (lane == 0) s_s[r] = rb[j0 + r] - acc;
            }
            __syncthreads();
            if (warp == 0) {
                // 64 x 64 triangular solve, rows lane and lane + 32
                double s0 = (lane < nv) ? s_s[lane] : 0.0;
                double s1 = (lane + 32 < nv) ? s_s[lane + 32] : 0.0;
                for (int c = 0; c < nv; ++c) {
                    const double d = Ls[c * LDS + c];
                    const double sc = __shfl_sync(0xffffffffu, (c < 32) ? s0 : s1, c & 31);
                    const double zc = sc / d;
                    logdet += log(d);
                    if (lane == (c & 31)) zb[j0 + c] = zc;
                    if (lane > c) s0 -= Ls[lane * LDS + c] * zc;
                    if (lane + 32 > c) s1 -= Ls[(lane + 32) * LDS + c] * zc;
                }
            }
        }
        __syncthreads();
        double q = 0.0;
        for (int c = tid; c < n; c += THREADS) q += zb[c] * zb[c];
        q = warp_sum(q);
        if (lane == 0) red[warp] = q;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < THREADS / 32; ++w) s += red[w];
            p.ll[mat] = -0.5 * s - logdet - 0.5 * (double)n * log(2.0 * 3.141592653589793);
        }
    }
}

}  // namespace sv

extern "C" int mf_solve_logdet(mf_handle_t h, int batch, int n, const double* L, int64_t ld,
                               int64_t stride, const double* resid, int64_t resid_stride,
                               double* z, double* ll, const int32_t* info, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!L || !resid || !z || !ll) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0 || n <= 0 || ld < n) MF_FAIL(h, MF_ERR_INVALID, "bad batch/n/ld");
    sv::Params p;
    p.batch = batch;
    p.n = n;
    p.L = L;
    p.ld = ld;
    p.stride = stride;
    p.resid = resid;
    p.resid_stride = resid_stride;
    p.z = z;
    p.ll = ll;
    p.info = info;
    const int slots = 4 * h->sm_count;
    const int grid = batch < slots ? batch : slots;
    sv::mf_solve_logdet_kernel<<<grid, sv::THREADS, 0, (cudaStream_t)stream>>>(p);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
