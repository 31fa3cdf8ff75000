// Keplerian mean function for a whole batch of parameter rows (the residual r = y - mean(b) of
// the likelihood, reference BIGgp.py:303 with means.py:143-216): one thread per (sample, epoch)
// runs the reference's fixed-count iteration for the eccentric anomaly in registers - its third
// order starting guess and `iterations` steps E += (M - M_k) / (1 - e cos E) with M_k lagging one
// step behind (means.py:160-168, 204-212) - in the reference's operation order.  Replaces
// ~8 element-wise launches per iteration (800 for the reference's 100 iterations) by one.
#include <math.h>

#include "mf_common.cuh"

namespace {

__global__ void mf_kepler_rv_kernel(int batch, int N, const double* __restrict__ t,
                                    const double* __restrict__ pars, int iterations, double* out,
                                    int64_t out_stride, int accumulate) {
    const double two_pi = 2.0 * 3.141592653589793;
    for (int b = blockIdx.y; b < batch; b += gridDim.y) {
        const double P = pars[5 * b], K = pars[5 * b + 1], e = pars[5 * b + 2], w = pars[5 * b + 3],
                     T0 = pars[5 * b + 4];
        const double root = sqrt((1.0 + e) / (1.0 - e));
        const double ecw = e * cos(w);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
            const double M = two_pi * (t[i] - T0) / P;
            double E0 = M + e * sin(M) + 0.5 * (e * e) * sin(2.0 * M);
            double M0 = E0 - e * sin(E0);
            for (int k = 0; k < iterations; ++k) {
                double s, c;
                sincos(E0, &s, &c);
                const double E1 = E0 + (M - M0) / (1.0 - e * c);
                M0 = E0 - e * s;
                E0 = E1;
            }
            const double nu = 2.0 * atan(root * tan(E0 / 2.0));
            const double v = K * (ecw + cos(w + nu));
            double* o = out + (size_t)b * out_stride + i;
            *o = accumulate ? *o + v : v;
        }
    }
}

}  // namespace

extern "C" int mf_kepler_rv(mf_handle_t h, int batch, int n_epochs, const double* t, const double* pars,
                            int iterations, double* out, int64_t out_stride, int accumulate, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!t || !pars || !out) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0 || n_epochs <= 0 || iterations < 0 || out_stride < n_epochs)
        MF_FAIL(h, MF_ERR_INVALID, "bad batch/n_epochs/iterations/out_stride");
    const int bx = (n_epochs + 127) / 128;
    mf_kepler_rv_kernel<<<dim3(bx, batch < 65535 ? batch : 65535), 128, 0, (cudaStream_t)stream>>>(
        batch, n_epochs, t, pars, iterations, out, out_stride, accumulate ? 1 : 0);
    MF_CUDA_CHECK(h, cudaGetLastError());
    return MF_OK;
}
