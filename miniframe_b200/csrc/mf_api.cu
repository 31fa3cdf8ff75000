// Handle management, layout helpers and the fused batch driver of the C ABI
// declared in include/miniframe_b200.h.
#include <new>

#include "mf_common.cuh"

extern "C" int mf_version(void) { return MF_VERSION; }

extern "C" const char* mf_status_string(int status) {
    switch (status) {
        case MF_OK: return "ok";
        case MF_ERR_INVALID: return "invalid argument";
        case MF_ERR_CUDA: return "CUDA error";
        case MF_ERR_WORKSPACE: return "workspace too small";
        case MF_ERR_UNSUPPORTED: return "unsupported";
        default: return "unknown status";
    }
}

extern "C" int mf_create(mf_handle_t* out, int device) {
    if (!out) return MF_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        return MF_ERR_CUDA;
    mf_handle_s* h = new (std::nothrow) mf_handle_s();
    if (!h) return MF_ERR_INVALID;
    h->device = device;
    h->err[0] = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete h;
        return MF_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    h->cc_major = prop.major;
    h->cc_minor = prop.minor;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    if (prop.major != 10) {     // built for sm_100a only: no other code path exists
        delete h;
        return MF_ERR_UNSUPPORTED;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        delete h;
        return MF_ERR_CUDA;
    }
    h->coop_groups = 3 * h->sm_count;
    h->coop_sync = nullptr;
    h->coop_dinv = nullptr;
    if (cudaMalloc(&h->coop_sync, (size_t)h->coop_groups * 8 * sizeof(unsigned)) != cudaSuccess ||
        cudaMalloc(&h->coop_dinv, (size_t)h->coop_groups * 512 * sizeof(double)) != cudaSuccess) {
        cudaFree(h->coop_sync);
        delete h;
        return MF_ERR_CUDA;
    }
    *out = h;
    return MF_OK;
}

extern "C" int mf_destroy(mf_handle_t h) {
    if (h) {
        cudaFree(h->coop_sync);
        cudaFree(h->coop_dinv);
    }
    delete h;
    return MF_OK;
}

extern "C" const char* mf_last_error(mf_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int mf_device_info(mf_handle_t h, int* sm_count, int* cc_major, int* cc_minor) {
    if (!h) return MF_ERR_INVALID;
    if (sm_count) *sm_count = h->sm_count;
    if (cc_major) *cc_major = h->cc_major;
    if (cc_minor) *cc_minor = h->cc_minor;
    return MF_OK;
}

extern "C" int64_t mf_ld(int n) { return mf_ld_of(n); }
extern "C" int64_t mf_matrix_stride(int n) { return mf_rows_of(n) * mf_ld_of(n); }
extern "C" size_t mf_workspace_bytes(int batch, int n) {
    return (size_t)batch * (size_t)mf_matrix_stride(n) * sizeof(double);
}

extern "C" int mf_loglike(mf_handle_t h, const mf_problem_t* prob, int batch, const double* t,
                          const double* diag_add, const double* hyper, const double* extra,
                          const double* resid, int64_t resid_stride, void* workspace,
                          size_t ws_bytes, double* ll, int32_t* info, void* stream) {
    if (!h) return MF_ERR_INVALID;
    if (!prob || !workspace || !ll || !info) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch must be positive");
    const int n = prob->n_series * prob->n_epochs;
    const int64_t ld = mf_ld_of(n);
    const int64_t stride = mf_matrix_stride(n);
    const size_t per = (size_t)stride * sizeof(double);
    size_t cap = ws_bytes / per;
    if (cap == 0) MF_FAIL(h, MF_ERR_WORKSPACE, "workspace holds no matrix: need %zu bytes per sample", per);
    // the factorisation keeps two matrices in flight per SM (three for n <= 1100): cut chunks at
    // whole waves
    const size_t wave = (n <= 1100 ? 3 : 2) * (size_t)h->sm_count;
    if (cap > wave && (size_t)batch > cap) cap = cap / wave * wave;
    if (((uintptr_t)workspace & 15) != 0) MF_FAIL(h, MF_ERR_INVALID, "workspace must be 16-byte aligned");
    double* K = static_cast<double*>(workspace);
    for (int b0 = 0; b0 < batch;) {
        const int nb = (int)((size_t)(batch - b0) < cap ? (size_t)(batch - b0) : cap);
        int rc = mf_build_cov(h, prob, nb, t, diag_add, hyper + (size_t)b0 * prob->hyper_len,
                              (extra && prob->extra_len > 0) ? extra + (size_t)b0 * prob->extra_len : nullptr,
                              0, K, ld, stride, stream);
        if (rc != MF_OK) return rc;
        rc = mf_potrf_loglike(h, nb, n, K, ld, stride,
                              resid ? resid + (size_t)b0 * resid_stride : nullptr, resid_stride,
                              ll + b0, info + b0, stream);
        if (rc != MF_OK) return rc;
        b0 += nb;
    }
    return MF_OK;
}
