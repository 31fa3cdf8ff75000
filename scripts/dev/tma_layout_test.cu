// Dumps the shared-memory image of one 5-D TMA box (the factorisation kernel's tensor map)
// so the (row, column) -> byte mapping can be checked on the host.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>

__global__ void k(const __grid_constant__ CUtensorMap tm, double* out, int c0, int blk8, int mat) {
    extern __shared__ __align__(16) unsigned char raw[];
    unsigned char* smem = raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(raw) & 1023u)) & 1023u);
    uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t bar = base + 8192;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(8192) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(base), "l"(&tm), "r"(c0), "r"(0), "r"(0), "r"(blk8), "r"(mat), "r"(bar) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra W;\n\t}" ::"r"(bar) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<double*>(smem)[i];
}

int main() {
    const int n = 100, ld = 112, rows = 104, batch = 3;
    const size_t stride = (size_t)rows * ld;
    double* h = (double*)malloc(batch * stride * 8);
    for (int b = 0; b < batch; ++b)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < ld; ++c) h[b * stride + (size_t)r * ld + c] = b * 1e6 + r * 1000 + c;
    double *d, *o;
    cudaMalloc(&d, batch * stride * 8);
    cudaMalloc(&o, 1024 * 8);
    cudaMemcpy(d, h, batch * stride * 8, cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t gdim[5] = {(cuuint64_t)ld, 4, 2, (cuuint64_t)(rows / 8), (cuuint64_t)batch};
    const cuuint64_t gstr[4] = {(cuuint64_t)(2 * ld * 8), (cuuint64_t)(ld * 8), (cuuint64_t)(8 * ld * 8), (cuuint64_t)(stride * 8)};
    const cuuint32_t box[5] = {16, 4, 2, 8, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult cr = ((encode_fn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)cr);
    const int c0 = 32, blk8 = 8, mat = 1;      // rows 64.., columns 32..47 of matrix 1
    k<<<1, 128, 8192 + 1024 + 64>>>(tm, o, c0, blk8, mat);
    printf("launch: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    double* r = (double*)malloc(1024 * 8);
    cudaMemcpy(r, o, 1024 * 8, cudaMemcpyDeviceToHost);
    // print, for each 128-byte smem row position p, which matrix row it holds and the column order
    for (int p = 0; p < 64; ++p) {
        printf("pos %2d:", p);
        for (int e = 0; e < 16; e += 2) {
            double v = r[p * 16 + e];
            long iv = (long)v;
            printf(" [m%ld r%ld c%ld]", iv / 1000000, (iv / 1000) % 1000, iv % 1000);
        }
        printf("\n");
        if (p == 9) p = 53;
    }
    return 0;
}
