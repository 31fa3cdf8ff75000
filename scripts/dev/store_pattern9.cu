// Dev probe: the covariance build's store pattern without any arithmetic or shared memory:
// 3 x 3 blocks of order N, lower triangle; per 32 x 32 pair tile (ti >= tj) 6 direct row segments
// (rows of the i tile) and 3 "transposed" ones (rows of the j tile), 256 bytes each.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int MODE>   // 0: as the kernel; 1: direct only; 2: all nine as direct-style rows (no transposed geometry)
__global__ void pat(double* base, long stride, int N, long ld, int batch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long o10 = (long)N * ld, o20 = 2 * o10, o11 = o10 + N, o21 = o20 + N, o22 = o20 + 2 * N;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        double* K = base + (long)b * stride;
        const int T = (N + 31) / 32;
        for (int ti = 0; ti < T; ++ti)
            for (int tj = 0; tj <= ti; ++tj) {
                for (int rr = 0; rr < 4; ++rr) {
                    const int i = ti * 32 + warp * 4 + rr, j = tj * 32 + lane;
                    if (i < N && j < N && (ti != tj || i >= j)) {
                        double* d = K + (long)i * ld + j;
                        d[0] = 1.0; d[o11] = 1.0; d[o22] = 1.0; d[o10] = 1.0; d[o20] = 1.0; d[o21] = 1.0;
                    }
                }
                if (MODE == 0)
                    for (int rr = 0; rr < 4; ++rr) {
                        const int jr = tj * 32 + warp * 4 + rr, ic = ti * 32 + lane;
                        if (ic < N && jr < N && ic > jr) {
                            double* d = K + (long)jr * ld + ic;
                            d[o10] = 2.0; d[o20] = 2.0; d[o21] = 2.0;
                        }
                    }
            }
    }
}
int main(int argc, char** argv) {
    const int batch = 2368;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grids[] = {74, 148, 296, 444, 592, 888, 1184, 2368};
    for (int N : {500, 512})
    for (int grid : grids) { const int pad = 0;
        const int n = 3 * N;
        const long ld = (n + 15) / 16 * 16 + pad, stride = ld * (n + 8);
        double* d; cudaMalloc(&d, stride * batch * 8);
        const double full = 4.0 * n * (double)(n + 1) * batch;
        const double direct_only = full * (6.0 * N * (N + 1) / 2) / (n * (double)(n + 1) / 2);
#define RUN(MODE, BYTES, LABEL) do { \
        pat<MODE><<<grid, 256>>>(d, stride, N, ld, batch); cudaDeviceSynchronize(); \
        cudaEventRecord(a); pat<MODE><<<grid, 256>>>(d, stride, N, ld, batch); cudaEventRecord(b); cudaEventSynchronize(b); \
        float ms; cudaEventElapsedTime(&ms, a, b); \
        printf("N %d  %-34s: %.3f ms  %.0f GB/s  %s\n", N, LABEL, ms, (BYTES) / ms / 1e6, cudaGetErrorString(cudaGetLastError())); } while (0)
        char lab[64]; snprintf(lab, 64, "grid %d  6 direct + 3 transp", grid);
        RUN(0, full, lab);
        cudaFree(d);
    }
    return 0;
}
