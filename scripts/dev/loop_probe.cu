// Ceiling of the factorisation's contraction loop: the same fragment loads (16-byte, swizzled
// addresses) and DMMA groups as mf_potrf.cu, from static shared-memory data, without the TMA
// ring.  MODE bits: 1 = mbarrier try_wait per chunk (always complete), 2 = release atomic per
// chunk, 4 = warp-uniform branches around the groups as in the product loop, 8 = software-
// pipelined loads (next group's B fragments before this group's DMMAs; 96 fragment+acc registers)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o loop_probe loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
constexpr int STAGE = 192 * 128;
template <int MODE>
__global__ void __launch_bounds__(256, 2) loop_probe(int chunks, int nil0, int nil1, double* sink) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int gp = (g >> 1) | ((g & 1) << 2);
    const int tbw = warp < 4 ? warp : 11 - warp;
    double* f = reinterpret_cast<double*>(sm);
    for (int i = tid; i < 3 * STAGE / 8; i += 256) f[i] = 1e-3 * (i % 97);
    const uint32_t bar = smem_u32(sm + 3 * STAGE);
    int* released = reinterpret_cast<int*>(sm + 3 * STAGE + 64);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        released[0] = released[1] = released[2] = 0;
    }
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    int use_s = 0;
    for (int kc = 0; kc < chunks; ++kc) {
        if (MODE & 1) mbar_wait(bar, 0);
        const unsigned char* sb = sm + use_s * STAGE;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int choff = ((4 * h + t) ^ gp) << 4;
            double2 a[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
                a[mi] = *reinterpret_cast<const double2*>(sb + (8 * tbw + 64 * mi + gp) * 128 + choff);
#pragma unroll
            for (int nq = 0; nq < 2; ++nq) {
                double2 b[4];
                const bool d0 = (MODE & 4) ? 4 * nq < nil0 : true, d1 = (MODE & 4) ? 4 * nq < nil1 : true;
                if (d0 || d1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        b[u] = *reinterpret_cast<const double2*>(sb + (128 + 8 * (4 * nq + u) + gp) * 128 + choff);
                }
                if (d0 && d1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].x, b[u].x);
                } else if (d0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[0][4 * nq + u][0], acc[0][4 * nq + u][1], a[0].x, b[u].x);
                } else if (d1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[1][4 * nq + u][0], acc[1][4 * nq + u][1], a[1].x, b[u].x);
                }
                if ((MODE & 2) && h == 1 && nq == 1) {
                    __syncwarp();
                    if (lane == 0) {
                        const int old = atomicAdd(released + use_s, 1);
                        if ((old & 7) == 7 && chunks < 0) sink[1] = old;       // never: stands for the refill
                    }
                }
                if (d0 && d1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int mi = 0; mi < 2; ++mi) dmma884(acc[mi][4 * nq + u][0], acc[mi][4 * nq + u][1], a[mi].y, b[u].y);
                } else if (d0) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[0][4 * nq + u][0], acc[0][4 * nq + u][1], a[0].y, b[u].y);
                } else if (d1) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) dmma884(acc[1][4 * nq + u][0], acc[1][4 * nq + u][1], a[1].y, b[u].y);
                }
            }
        }
        if (++use_s == 3) use_s = 0;
    }
    double s = 0.0;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) s += acc[mi][ni][0] + acc[mi][ni][1];
    if (s == 123.456) sink[0] = s;
}

template <int MODE>
static void run(const char* name, int per_sm, int chunks) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int smem = per_sm == 1 ? 120 * 1024 : 3 * STAGE + 1024;
    cudaFuncSetAttribute(loop_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    double* sink;
    cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        loop_probe<MODE><<<sms * per_sm, 256, smem>>>(chunks, 8, 8, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double flops = (double)sms * per_sm * 8 * chunks * 64.0 * 512.0;
    printf("%-44s %d CTA/SM: %8.3f ms  %6.2f TFLOP/s %s\n", name, per_sm, best, flops / best / 1e9,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(sink);
}

int main() {
    const int chunks = 20000;
    for (int per_sm = 1; per_sm <= 2; ++per_sm) {
        run<0>("loads + DMMA groups", per_sm, chunks);
        run<1>("+ mbarrier try_wait per chunk", per_sm, chunks);
        run<3>("+ try_wait + release atomic", per_sm, chunks);
        run<4>("loads + DMMA, warp-uniform branches", per_sm, chunks);
        run<7>("branches + try_wait + release (product)", per_sm, chunks);
    }
    return 0;
}
