// Register-resident FP64 throughput probes (roofline denominators).  mode 0
// issues mma.sync m8n8k4 f64 (DMMA.8x8x4) back to back on 16 independent
// accumulator pairs per warp; mode 1 issues DFMA on 16 independent chains.
#include "mf_common.cuh"

namespace {

__global__ void __launch_bounds__(256) probe_dmma_kernel(int iters, double* sink) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}

__global__ void __launch_bounds__(256) probe_dfma_kernel(int iters, double* sink) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) sink[0] = s;
}

}  // namespace

extern "C" int mf_probe_fp64(mf_handle_t h, int mode, int iters, float* ms_host,
                             double* flops_host, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!ms_host || !flops_host || iters <= 0) MF_FAIL(h, MF_ERR_INVALID, "bad probe arguments");
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    MF_CUDA_CHECK(h, cudaEventCreate(&e0));
    MF_CUDA_CHECK(h, cudaEventCreate(&e1));
    const int ctas = h->sm_count * 4, threads = 256;
    for (int rep = 0; rep < 2; ++rep) {      // first pass warms up
        MF_CUDA_CHECK(h, cudaEventRecord(e0, st));
        if (mode == 0)
            probe_dmma_kernel<<<ctas, threads, 0, st>>>(iters, nullptr);
        else
            probe_dfma_kernel<<<ctas, threads, 0, st>>>(iters, nullptr);
        MF_CUDA_CHECK(h, cudaGetLastError());
        MF_CUDA_CHECK(h, cudaEventRecord(e1, st));
        MF_CUDA_CHECK(h, cudaEventSynchronize(e1));
    }
    MF_CUDA_CHECK(h, cudaEventElapsedTime(ms_host, e0, e1));
    const double warps = (double)ctas * threads / 32.0;
    if (mode == 0)
        *flops_host = warps * (double)iters * 16.0 * (2.0 * 8 * 8 * 4);
    else
        *flops_host = warps * 32.0 * (double)iters * 16.0 * 2.0;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MF_OK;
}
