// Gradient of the log-likelihood with respect to the hyper-parameters (SURVEY.md section 8, row f3;
// not in the reference, whose optimisers would difference minus_log_likelihood, BIGgp.py:315-317):
//
//     d ll / d theta_k = -1/2 sum_{R,C} W[R][C] dK[R][C]/d theta_k,    W = K^-1 - alpha alpha^T,
//
// with alpha = K^-1 r.  This file is the contraction: W (one n x n matrix per sample, in the
// EPOCH-INTERLEAVED order the likelihood path factorises in, r = 3 i + P) is streamed once
// (HBM-read bound, 4 n^2 bytes per sample: only the lower triangle is read), and dK/dtheta is never
// formed: like the covariance build, one sincos + one exp per unordered pair of epochs give the
// base functions E, AE, Gdd AND their derivatives with respect to the kernel's shape parameters;
// the pair's 3 x 3 micro-block of W is folded into 24 running sums per thread
//     U_E, U_G [6 block pairs]   sum of W E, W Gdd over each block (P <= Q)
//     U_A      [3 off-diagonal]  sum of W AE
//     U_W      [6]               sum of W over same-epoch entries (white noise, errors, jitter)
//     g        [3]               the shape-parameter gradients themselves
// from which a second, tiny kernel assembles the gradient with the derivatives of the block
// coefficients (BIGgp.py:167-234).  Partial sums are combined in a fixed order: the result is
// reproducible bit for bit.
#include "mf_cov.cuh"

namespace {

constexpr int GT = 32;            // pair-tile edge
constexpr int GTHREADS = 256;
constexpr int NACC = 24;

struct GradParams {
    mf_problem_t prob;
    int batch;
    int n_tiles;
    const double* t;
    const double* hyper;
    const double* extra;
    const double* W;
    int64_t ld, stride;
    double* partial;              // [batch][gridDim.x][NACC]
    double* grad;                 // [batch][hyper_len + extra_len]
};

struct GBases {
    double E, AE, Gdd;
    double dE[3], dAE[3], dGdd[3];
};

// base functions and their derivatives with respect to (ell_e, ell_p, P) / (ell) at the signed lag r
template <int KERNEL>
__device__ __forceinline__ void grad_bases(double r, const double* kp, GBases& o) {
    double A, q0, dA[3], dq[3];
    if (KERNEL == MF_KERNEL_QP) {
        const double le = kp[0], lp = kp[1], P = kp[2];
        const double f2 = lp * lp, f3 = le * le;
        const double x = (MF_PI * r) / P;
        double s, c;
        sincos(x, &s, &c);
        const double c2s2 = c * c - s * s, sc = s * c;
        o.E = exp(-(2.0 * s * s / f2) - 0.5 * r * r / f3);
        A = MF_4PI * sc / (f2 * P) + r / f3;
        q0 = 1.0 / f3 + MF_4PI2 * c2s2 / (f2 * P * P);
        const double dx = -x / P;
        const double le3 = le * le * le, lp3 = lp * lp * lp;
        o.dE[0] = o.E * r * r / le3;
        o.dE[1] = o.E * 4.0 * s * s / lp3;
        o.dE[2] = o.E * (-4.0 * sc * dx / f2);
        dA[0] = -2.0 * r / le3;
        dA[1] = -2.0 * MF_4PI * sc / (lp3 * P);
        dA[2] = MF_4PI / f2 * (c2s2 * dx / P - sc / (P * P));
        dq[0] = -2.0 / le3;
        dq[1] = -2.0 * MF_4PI2 * c2s2 / (lp3 * P * P);
        dq[2] = MF_4PI2 / f2 * (-4.0 * sc * dx / (P * P) - 2.0 * c2s2 / (P * P * P));
    } else {
        const double ell = kp[0];
        const double f2 = ell * ell, l3 = f2 * ell;
        o.E = exp(-0.5 * r * r / f2);
        A = r / f2;
        q0 = 1.0 / f2;
        o.dE[0] = o.E * r * r / l3;
        dA[0] = -2.0 * r / l3;
        dq[0] = -2.0 / l3;
        o.dE[1] = o.dE[2] = dA[1] = dA[2] = dq[1] = dq[2] = 0.0;
    }
    o.AE = A * o.E;
    const double q = q0 - A * A;
    o.Gdd = q * o.E;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.dAE[k] = dA[k] * o.E + A * o.dE[k];
        o.dGdd[k] = (dq[k] - 2.0 * A * dA[k]) * o.E + q * o.dE[k];
    }
}

// index of the block pair P <= Q in the 6-vectors: (0,0) (1,1) (2,2) (0,1) (0,2) (1,2)
__device__ __forceinline__ int pair_index(int P, int Q) { return P == Q ? P : (P + Q + 2); }

template <int KERNEL>
__global__ void __launch_bounds__(GTHREADS)
mf_grad_contract_kernel(const GradParams p) {
    __shared__ double coef[15];            // cE[6], cGdd[6], cAE[3] of the upper-form block pairs
    __shared__ double kpar[4];
    __shared__ double red[GTHREADS / 32][NACC];
    const int N = p.prob.n_epochs;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npar = KERNEL == MF_KERNEL_QP ? 4 : 2;
    const int64_t ld = p.ld;

    for (int b = blockIdx.y; b < p.batch; b += gridDim.y) {
        const double* hy = p.hyper + (size_t)b * p.prob.hyper_len;
        __syncthreads();
        if (tid < 6) {
            const BlockCoef u = mf_big_pair_coef(tid, hy + npar);
            coef[tid] = u.cE;
            coef[6 + tid] = u.cGdd;
            if (tid >= 3) coef[12 + tid - 3] = u.cAE;
        }
        if (tid >= 8 && tid < 8 + npar) kpar[tid - 8] = hy[tid - 8];
        __syncthreads();
        const double* Wb = p.W + (size_t)b * p.stride;
        double acc[NACC];
#pragma unroll
        for (int k = 0; k < NACC; ++k) acc[k] = 0.0;
        double* const uE = acc;            // [6]
        double* const uG = acc + 6;        // [6]
        double* const uA = acc + 12;       // [3]: pairs (0,1) (0,2) (1,2)
        double* const uW = acc + 15;       // [6]
        double* const gk = acc + 21;       // [3]

        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            int tile_i = (int)((sqrtf(8.0f * (float)tile + 1.0f) - 1.0f) * 0.5f);
            while ((tile_i + 1) * (tile_i + 2) / 2 <= tile) ++tile_i;
            while (tile_i * (tile_i + 1) / 2 > tile) --tile_i;
            const int tile_j = tile - tile_i * (tile_i + 1) / 2;
            const int j = tile_j * GT + lane;
            const double tj = (j < N) ? p.t[j] : 0.0;
#pragma unroll 1
            for (int rr = 0; rr < GT / 8; ++rr) {
                const int i = tile_i * GT + warp + 8 * rr;
                if (i >= N || j >= N || i < j) continue;
                GBases bs;
                grad_bases<KERNEL>(p.t[i] - tj, kpar, bs);
                const double* wp = Wb + (size_t)(3 * i) * ld + 3 * j;
                double mE = 0.0, mG = 0.0, mA = 0.0;
                if (i != j) {
                    double w[3][3];
#pragma unroll
                    for (int P = 0; P < 3; ++P)
#pragma unroll
                        for (int Q = 0; Q < 3; ++Q) w[P][Q] = wp[(size_t)P * ld + Q];
#pragma unroll
                    for (int P = 0; P < 3; ++P) {
                        const double w2 = 2.0 * w[P][P];      // entries (i, j) and (j, i) of block (P, P)
                        uE[P] += w2 * bs.E;
                        uG[P] += w2 * bs.Gdd;
                        mE += w2 * coef[P];
                        mG += w2 * coef[6 + P];
                    }
#pragma unroll
                    for (int P = 0; P < 2; ++P)
#pragma unroll
                        for (int Q = P + 1; Q < 3; ++Q) {
                            // block (P, Q): entry (i, j) at lag r and entry (j, i) at lag -r, which by
                            // symmetry of W is W[(Q, i), (P, j)]
                            const int k = pair_index(P, Q);
                            const double ws = w[P][Q] + w[Q][P], wd = w[P][Q] - w[Q][P];
                            uE[k] += ws * bs.E;
                            uG[k] += ws * bs.Gdd;
                            uA[k - 3] += wd * bs.AE;
                            mE += 2.0 * ws * coef[k];             // and the transposed block (Q, P)
                            mG += 2.0 * ws * coef[6 + k];
                            mA += 2.0 * wd * coef[12 + k - 3];
                        }
                } else {
#pragma unroll
                    for (int P = 0; P < 3; ++P)
#pragma unroll
                        for (int Q = 0; Q <= P; ++Q) {
                            const int k = pair_index(Q, P);
                            const double wl = wp[(size_t)P * ld + Q];
                            uE[k] += wl * bs.E;
                            uG[k] += wl * bs.Gdd;
                            uW[k] += wl;
                            const double m = (P == Q) ? wl : 2.0 * wl;
                            mE += m * coef[k];
                            mG += m * coef[6 + k];
                        }
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) gk[k] += mE * bs.dE[k] + mG * bs.dGdd[k] + mA * bs.dAE[k];
            }
        }
        // fixed-order reduction: lanes (butterfly), then warps 0..7
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            const double v = warp_sum(acc[k]);
            if (lane == 0) red[warp][k] = v;
        }
        __syncthreads();
        if (tid < NACC) {
            double sum = 0.0;
            for (int w = 0; w < GTHREADS / 32; ++w) sum += red[w][tid];
            p.partial[((size_t)b * gridDim.x + blockIdx.x) * NACC + tid] = sum;
        }
    }
}

// gradient from the running sums: one warp per sample
__global__ void mf_grad_finish_kernel(const GradParams p, int ctas) {
    const int b = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= p.batch) return;
    double u = 0.0;
    if (lane < NACC)
        for (int c = 0; c < ctas; ++c) u += p.partial[((size_t)b * ctas + c) * NACC + lane];
    // every lane gets all 24 sums
    double U[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) U[k] = __shfl_sync(0xffffffffu, u, k);
    if (lane != 0) return;
    const double* uE = U;
    const double* uG = U + 6;
    const double* uA = U + 12;
    const double* uW = U + 15;
    const double* gk = U + 21;
    const int qp = p.prob.kernel == MF_KERNEL_QP;
    const int npar = qp ? 4 : 2, nshape = npar - 1;
    const double* hy = p.hyper + (size_t)b * p.prob.hyper_len;
    const double wn = hy[npar - 1];
    const double vc = hy[npar], vr = hy[npar + 1], lc = hy[npar + 2], bc = hy[npar + 3], br = hy[npar + 4];
    double* g = p.grad + (size_t)b * (p.prob.hyper_len + p.prob.extra_len);
    for (int k = 0; k < nshape; ++k) g[k] = -0.5 * gk[k];
    // white noise sits on the diagonal of EVERY block (quirk Q1): sum of mult cW U_W
    double cW[6];
    for (int k = 0; k < 6; ++k) cW[k] = mf_big_pair_coef(k, hy + npar).cW;
    double sw = 0.0;
    for (int k = 0; k < 6; ++k) sw += (k < 3 ? 1.0 : 2.0) * cW[k] * uW[k];
    g[npar - 1] = -0.5 * 2.0 * wn * sw;
    // d(cE, cAE, cGdd, cW)/d(vc, vr, lc, bc, br) per block pair, (0,0) (1,1) (2,2) (0,1) (0,2) (1,2)
    const double wn2 = wn * wn;
    double d[5];
    // vc
    d[0] = (2 * vc * uE[0] + 2 * (vc + vr) * wn2 * uW[0]) +
           2.0 * (lc * uE[3] + lc * wn2 * uW[3]) +
           2.0 * (bc * uE[4] + br * uA[1] + (bc + br) * wn2 * uW[4]);
    // vr
    d[1] = (2 * vr * uG[0] + 2 * (vc + vr) * wn2 * uW[0]) +
           2.0 * (-lc * uA[0] + lc * wn2 * uW[3]) +
           2.0 * (-bc * uA[1] + br * uG[4] + (br + bc) * wn2 * uW[4]);
    // lc
    d[2] = (2 * lc * uE[1] + 2 * lc * wn2 * uW[1]) +
           2.0 * (vc * uE[3] - vr * uA[0] + (vc + vr) * wn2 * uW[3]) +
           2.0 * (bc * uE[5] + br * uA[2] + (bc + br) * wn2 * uW[5]);
    // bc
    d[3] = (2 * bc * uE[2] + 2 * (bc + br) * wn2 * uW[2]) +
           2.0 * (vc * uE[4] - vr * uA[1] + (vc + vr) * wn2 * uW[4]) +
           2.0 * (lc * uE[5] + lc * wn2 * uW[5]);
    // br
    d[4] = (2 * br * uG[2] + 2 * (bc + br) * wn2 * uW[2]) +
           2.0 * (vc * uA[1] + vr * uG[4] + (vr + vc) * wn2 * uW[4]) +
           2.0 * (lc * uA[2] + lc * wn2 * uW[5]);
    for (int m = 0; m < 5; ++m) g[npar + m] = -0.5 * d[m];
    if (p.prob.framework == MF_FRAMEWORK_BIGGP_JITT && p.extra != nullptr)
        for (int q = 0; q < 3; ++q)
            g[p.prob.hyper_len + q] = -0.5 * 2.0 * p.extra[(size_t)b * p.prob.extra_len + q] * uW[q];
}

}  // namespace

extern "C" int mf_grad_ctas(int n_epochs) {
    const int T = (n_epochs + GT - 1) / GT;
    const int tiles = T * (T + 1) / 2;
    return tiles < 16 ? tiles : 16;
}

extern "C" int mf_grad_contract(mf_handle_t h, const mf_problem_t* prob, int batch, const double* t,
                                const double* hyper, const double* extra, const double* W,
                                int64_t ld, int64_t stride, double* partial, double* grad,
                                void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!prob || !t || !hyper || !W || !partial || !grad) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch must be positive");
    if (prob->framework != MF_FRAMEWORK_BIGGP && prob->framework != MF_FRAMEWORK_BIGGP_JITT)
        MF_FAIL(h, MF_ERR_UNSUPPORTED, "the gradient is built for BIGgp / BIGgpPlusJitt");
    if (prob->n_series != 3 || prob->n_epochs < 1) MF_FAIL(h, MF_ERR_INVALID, "bad n_series/n_epochs");
    if (prob->kernel != MF_KERNEL_SE && prob->kernel != MF_KERNEL_QP) MF_FAIL(h, MF_ERR_INVALID, "unknown kernel id");
    if (prob->hyper_len != (prob->kernel == MF_KERNEL_QP ? 9 : 7)) MF_FAIL(h, MF_ERR_INVALID, "BIGgp hyper_len");
    if (prob->framework == MF_FRAMEWORK_BIGGP_JITT && (!extra || prob->extra_len < 3))
        MF_FAIL(h, MF_ERR_INVALID, "BIGgpPlusJitt needs 3 jitters");
    const int64_t n = 3 * (int64_t)prob->n_epochs;
    if (ld < n || (stride != 0 && stride < n * ld)) MF_FAIL(h, MF_ERR_INVALID, "ld/stride too small");
    GradParams p;
    p.prob = *prob;
    p.batch = batch;
    const int T = (prob->n_epochs + GT - 1) / GT;
    p.n_tiles = T * (T + 1) / 2;
    p.t = t;
    p.hyper = hyper;
    p.extra = prob->extra_len > 0 ? extra : nullptr;
    p.W = W;
    p.ld = ld;
    p.stride = stride;
    p.partial = partial;
    p.grad = grad;
    cudaStream_t st = (cudaStream_t)stream;
    const int ctas = mf_grad_ctas(prob->n_epochs);
    const dim3 grid(ctas, batch < 65535 ? batch : 65535);
    if (prob->kernel == MF_KERNEL_QP)
        mf_grad_contract_kernel<MF_KERNEL_QP><<<grid, GTHREADS, 0, st>>>(p);
    else
        mf_grad_contract_kernel<MF_KERNEL_SE><<<grid, GTHREADS, 0, st>>>(p);
    MF_CUDA_CHECK(h, cudaGetLastError());
    mf_grad_finish_kernel<<<(batch + 3) / 4, 128, 0, st>>>(p, ctas);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
