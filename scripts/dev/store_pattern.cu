// Dev probe: write bandwidth of "lower-triangle tile" store patterns as a function of the
// contiguous granule.  One CTA per matrix (n x n doubles, ld padded), 8 warps; the CTA walks
// square tiles of edge W (tile_i >= tile_j) and every warp writes W-double row segments.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/dev/store_pattern scripts/dev/store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
template <int W, int VEC>
__global__ void pat(double* base, long stride, int n, long ld, int batch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = blockIdx.x; b < batch; b += gridDim.x) {
        double* K = base + (long)b * stride;
        const int T = (n + W - 1) / W;
        for (int ti = 0; ti < T; ++ti)
            for (int tj = 0; tj <= ti; ++tj)
                for (int r = warp; r < W; r += 8) {
                    const int row = ti * W + r;
                    if (row >= n) continue;
                    double* dst = K + (long)row * ld + (long)tj * W;
                    if (VEC == 1) {
                        for (int c = lane; c < W; c += 32) if (tj * W + c < n) dst[c] = 1.0;
                    } else {
                        for (int c = 2 * lane; c < W; c += 64) if (tj * W + c + 1 < n) *reinterpret_cast<double2*>(dst + c) = make_double2(1.0, 2.0);
                    }
                }
    }
}
int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1500;
    const int batch = 1184;
    const long ld = (n + 15) / 16 * 16, stride = ld * (n + 8);
    double* d; cudaMalloc(&d, stride * batch * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const double bytes = 4.0 * n * (double)(n + 1) * batch;   // lower triangle
#define RUN(W, VEC, CTAS) do { \
    pat<W, VEC><<<CTAS, 256>>>(d, stride, n, ld, batch); cudaDeviceSynchronize(); \
    cudaEventRecord(a); pat<W, VEC><<<CTAS, 256>>>(d, stride, n, ld, batch); cudaEventRecord(b); cudaEventSynchronize(b); \
    float ms; cudaEventElapsedTime(&ms, a, b); \
    printf("n %d  granule %4d B  vec %d  ctas %4d : %.3f ms  %.0f GB/s (lower-triangle bytes)  %s\n", n, W * 8, VEC, CTAS, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError())); } while (0)
    RUN(32, 1, 1184); RUN(64, 1, 1184); RUN(128, 1, 1184); RUN(256, 1, 1184);
    RUN(64, 2, 1184); RUN(128, 2, 1184); RUN(256, 2, 1184);
    RUN(32, 1, 444); RUN(128, 2, 444); RUN(32, 1, 296); RUN(128, 2, 296);
    return 0;
}
