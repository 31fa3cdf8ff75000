// Dev probe, continued: does the ORDER in which one CTA visits the nine output blocks matter?
//   mode 0: per pair tile, every row writes its 6 direct segments back to back (the kernel)
//   mode 1: block outermost: the CTA sweeps the whole lower triangle once per output block
//   mode 2: super tiles of 32 x (32 S) pairs: per super tile, block by block, S segments per row
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE, int S>
__global__ void pat(double* base, long stride, int N, long ld, int batch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long off[6] = {0, (long)N * ld + N, 2L * N * ld + 2 * N, (long)N * ld, 2L * N * ld, 2L * N * ld + N};
    const int T = (N + 31) / 32;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        double* K = base + (long)b * stride;
        if (MODE == 0) {
            for (int ti = 0; ti < T; ++ti)
                for (int tj = 0; tj <= ti; ++tj)
                    for (int rr = 0; rr < 4; ++rr) {
                        const int i = ti * 32 + warp * 4 + rr, j = tj * 32 + lane;
                        if (i < N && j < N && (ti != tj || i >= j))
                            for (int k = 0; k < 6; ++k) K[(long)i * ld + j + off[k]] = 1.0;
                    }
        } else if (MODE == 1) {
            for (int k = 0; k < 6; ++k)
                for (int ti = 0; ti < T; ++ti)
                    for (int tj = 0; tj <= ti; ++tj)
                        for (int rr = 0; rr < 4; ++rr) {
                            const int i = ti * 32 + warp * 4 + rr, j = tj * 32 + lane;
                            if (i < N && j < N && (ti != tj || i >= j)) K[(long)i * ld + j + off[k]] = 1.0;
                        }
        } else {
            for (int ti = 0; ti < T; ++ti)
                for (int tj0 = 0; tj0 <= ti; tj0 += S)
                    for (int k = 0; k < 6; ++k)
                        for (int rr = 0; rr < 4; ++rr)
                            for (int s = 0; s < S && tj0 + s <= ti; ++s) {
                                const int tj = tj0 + s;
                                const int i = ti * 32 + warp * 4 + rr, j = tj * 32 + lane;
                                if (i < N && j < N && (ti != tj || i >= j)) K[(long)i * ld + j + off[k]] = 1.0;
                            }
        }
    }
}
int main() {
    const int batch = 1184;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int N : {500, 512}) {
        const int n = 3 * N;
        const long ld = (n + 15) / 16 * 16, stride = ld * (n + 8);
        double* d; cudaMalloc(&d, stride * batch * 8);
        const double bytes = 8.0 * 6.0 * N * (N + 1) / 2 * batch;
#define RUN(MODE, S, LABEL) do { \
        pat<MODE, S><<<1184, 256>>>(d, stride, N, ld, batch); cudaDeviceSynchronize(); \
        cudaEventRecord(a); pat<MODE, S><<<1184, 256>>>(d, stride, N, ld, batch); cudaEventRecord(b); cudaEventSynchronize(b); \
        float ms; cudaEventElapsedTime(&ms, a, b); \
        printf("N %d  %-44s: %.3f ms  %.0f GB/s  %s\n", N, LABEL, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError())); } while (0)
        RUN(0, 1, "6 direct, row by row (kernel order)");
        RUN(1, 1, "6 direct, block outermost");
        RUN(2, 2, "6 direct, super tile 32 x 64, block by block");
        RUN(2, 4, "6 direct, super tile 32 x 128, block by block");
        RUN(2, 16, "6 direct, whole tile row, block by block");
        cudaFree(d);
    }
    return 0;
}
