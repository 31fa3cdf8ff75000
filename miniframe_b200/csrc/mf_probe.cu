// Register-resident FP64 throughput probes (roofline denominators).
//   mode % 10: 0 = mma.sync m8n8k4 f64 (DMMA.8x8x4), 1 = DFMA, 2 = m16n8k4, 3 = m16n8k8,
//              4 = m16n8k16 (all f64 mma.sync shapes sm_100a accepts)
//              5..8 = m8n8k4 with 1, 2, 4, 8 independent accumulator tiles per warp
//   mode / 10: occupancy: 0 = 4 CTAs x 256 threads per SM, 1 = 1 x 128, 2 = 1 x 256, 3 = 2 x 256
// Each warp keeps 16 (8x8x4) / 8 (m16) independent accumulator tiles in flight.
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

// dependent-issue behaviour: ILP independent accumulator tiles per warp
template <int ILP>
__global__ void __launch_bounds__(256) probe_dmma_ilp_kernel(int iters, double* sink) {
    double acc[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 16 / ILP; ++rep)
#pragma unroll
            for (int i = 0; i < ILP; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) sink[0] = s;
}

template <int K>
__global__ void __launch_bounds__(256) probe_m16_kernel(int iters, double* sink) {
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0;
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + (threadIdx.x + i) * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 - (threadIdx.x + i) * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (K == 4)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (K == 8)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
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
    const int shape = mode % 10, occ = mode / 10;
    if (shape > 8 || occ > 3 || mode < 0) MF_FAIL(h, MF_ERR_INVALID, "unknown probe mode %d", mode);
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    MF_CUDA_CHECK(h, cudaEventCreate(&e0));
    MF_CUDA_CHECK(h, cudaEventCreate(&e1));
    const int per_sm[4] = {4, 1, 1, 2};
    const int thr[4] = {256, 128, 256, 256};
    const int ctas = h->sm_count * per_sm[occ], threads = thr[occ];
    for (int rep = 0; rep < 2; ++rep) {      // first pass warms up
        MF_CUDA_CHECK(h, cudaEventRecord(e0, st));
        switch (shape) {
            case 0: probe_dmma_kernel<<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 1: probe_dfma_kernel<<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 2: probe_m16_kernel<4><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 3: probe_m16_kernel<8><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 4: probe_m16_kernel<16><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 5: probe_dmma_ilp_kernel<1><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 6: probe_dmma_ilp_kernel<2><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            case 7: probe_dmma_ilp_kernel<4><<<ctas, threads, 0, st>>>(iters, nullptr); break;
            default: probe_dmma_ilp_kernel<8><<<ctas, threads, 0, st>>>(iters, nullptr); break;
        }
        MF_CUDA_CHECK(h, cudaGetLastError());
        MF_CUDA_CHECK(h, cudaEventRecord(e1, st));
        MF_CUDA_CHECK(h, cudaEventSynchronize(e1));
    }
    MF_CUDA_CHECK(h, cudaEventElapsedTime(ms_host, e0, e1));
    const double warps = (double)ctas * threads / 32.0;
    const double kk[9] = {4, 0, 4, 8, 16, 0, 0, 0, 0};
    if (shape == 1)
        *flops_host = warps * 32.0 * (double)iters * 16.0 * 2.0;
    else if (shape == 0 || shape >= 5)
        *flops_host = warps * (double)iters * 16.0 * (2.0 * 8 * 8 * 4);
    else
        *flops_host = warps * (double)iters * 8.0 * (2.0 * 16 * 8 * kk[shape]);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MF_OK;
}
