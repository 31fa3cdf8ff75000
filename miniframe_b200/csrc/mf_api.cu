// Handle management, layout helpers and the fused batch driver of the C ABI
// declared in include/miniframe_b200.h.
#include <new>

#include "mf_common.cuh"

extern "C" int mf_version(void) { return MF_VERSION; }

mf_tmap_encode_fn mf_tmap_encoder(mf_handle_s* h) {
    // C++11 guarantees that concurrent callers initialise this exactly once
    static const mf_tmap_encode_fn fn = []() -> mf_tmap_encode_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return (mf_tmap_encode_fn)p;
    }();
    if (!fn) snprintf(h->err, sizeof(h->err), "cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

const CUtensorMap* mf_tmap_find(mf_handle_s* h, const mf_tmap_key& key, mf_tmap_slot** fresh) {
    for (int i = 0; i < MF_TMAP_SLOTS; ++i) {
        const mf_tmap_key& k = h->tmaps[i].key;
        if (k.kind == key.kind && k.base == key.base && k.ld == key.ld && k.stride == key.stride &&
            k.n == key.n && k.batch == key.batch)
            return &h->tmaps[i].map;
    }
    mf_tmap_slot* s = &h->tmaps[h->tmap_next];
    h->tmap_next = (h->tmap_next + 1) % MF_TMAP_SLOTS;
    s->key = key;
    s->key.kind = 0;            // becomes valid once the caller has encoded it
    *fresh = s;
    return nullptr;
}

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
    memset(&h->opt, 0, sizeof(h->opt));
    memset(h->tmaps, 0, sizeof(h->tmaps));
    h->tmap_next = 0;
    h->coop_groups = 3 * h->sm_count;
    h->coop_sync = nullptr;
    h->coop_words = 512 * 1024 - 8 * h->coop_groups;
    if (cudaMalloc(&h->coop_sync, (size_t)(8 * h->coop_groups + h->coop_words) * sizeof(unsigned)) != cudaSuccess) {
        delete h;
        return MF_ERR_CUDA;
    }
    *out = h;
    return MF_OK;
}

extern "C" int mf_destroy(mf_handle_t h) {
    if (h) {
        cudaFree(h->coop_sync);
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

extern "C" int mf_set_option(mf_handle_t h, const char* name, int value) {
    if (!h) return MF_ERR_INVALID;
    if (!name) MF_FAIL(h, MF_ERR_INVALID, "null option name");
    if (!strcmp(name, "potrf_tile")) {
        if (value < 0 || value > 2) MF_FAIL(h, MF_ERR_INVALID, "potrf_tile is 0 (choose), 1 (64 rows) or 2 (128 rows)");
        h->opt.potrf_tile = value;
    } else if (!strcmp(name, "potrf_group")) {
        if (value < 0) MF_FAIL(h, MF_ERR_INVALID, "potrf_group must be >= 0");
        h->opt.potrf_group = value;
    } else if (!strcmp(name, "potrf_map")) {
        if (value < 0 || value > 2) MF_FAIL(h, MF_ERR_INVALID, "potrf_map is 0 (choose), 1 or 2");
        h->opt.potrf_map = value;
    } else if (!strcmp(name, "potrf_lookahead")) {
        if (value < 0 || value > 2) MF_FAIL(h, MF_ERR_INVALID, "potrf_lookahead is 0 (choose), 1 (off) or 2 (on)");
        h->opt.potrf_lookahead = value;
    } else if (!strcmp(name, "l2_persist_mb")) {
        // L2 set-aside for the lines the factorisation's TMA loads mark evict_last (the block column's own
        // rows, re-read by every tile); -1 = as much as the device allows.  Off by default: it is a DEVICE
        // limit, and 64 MB of it make the factorisation 1 % faster under the board's power cap but halve
        // the covariance build's write bandwidth (6.9 -> 12.9 ms per 4096 matrices).
        int maxb = 0;
        MF_CUDA_CHECK(h, cudaDeviceGetAttribute(&maxb, cudaDevAttrMaxPersistingL2CacheSize, h->device));
        size_t want = value < 0 ? (size_t)maxb : (size_t)value << 20;
        if (want > (size_t)maxb) want = (size_t)maxb;
        MF_CUDA_CHECK(h, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    } else if (!strcmp(name, "cov_ctas")) {
        if (value < 0) MF_FAIL(h, MF_ERR_INVALID, "cov_ctas must be >= 0");
        h->opt.cov_ctas = value;
    } else if (!strcmp(name, "experiment")) {
#ifdef MF_EXPERIMENTS
        h->opt.experiment = value;
#else
        if (value != 0) MF_FAIL(h, MF_ERR_UNSUPPORTED, "timing experiments are not compiled into this library");
#endif
    } else {
        MF_FAIL(h, MF_ERR_INVALID, "unknown option '%s'", name);
    }
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
    if (!prob || !workspace || !ll || !info || !t || !diag_add || !hyper) MF_FAIL(h, MF_ERR_INVALID, "null pointer");
    if (batch <= 0) MF_FAIL(h, MF_ERR_INVALID, "batch must be positive");
    if (prob->n_series < 1 || prob->n_series > MF_MAX_SERIES || prob->n_epochs < 1)
        MF_FAIL(h, MF_ERR_INVALID, "bad n_series/n_epochs");
    if ((int64_t)prob->n_series * prob->n_epochs > INT32_MAX / 2) MF_FAIL(h, MF_ERR_INVALID, "matrix order too large");
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
    // BIGgp family without nugget: epoch-interleaved order (one dense, TMA-stored output stream);
    // the residual goes into row n of every matrix in the same order and is solved in place
    const bool interleaved = (prob->framework == MF_FRAMEWORK_BIGGP || prob->framework == MF_FRAMEWORK_BIGGP_JITT) &&
                             prob->nugget == 0.0;
    for (int b0 = 0; b0 < batch;) {
        const int nb = (int)((size_t)(batch - b0) < cap ? (size_t)(batch - b0) : cap);
        int rc = mf_build_cov(h, prob, nb, t, diag_add, hyper + (size_t)b0 * prob->hyper_len,
                              (extra && prob->extra_len > 0) ? extra + (size_t)b0 * prob->extra_len : nullptr,
                              interleaved ? MF_COV_LOWER_INTERLEAVED : MF_COV_LOWER, K, ld, stride, stream);
        if (rc != MF_OK) return rc;
        const double* rb = resid ? resid + (size_t)b0 * resid_stride : nullptr;
        int64_t rs = resid_stride;
        if (interleaved && resid) {
            rc = mf_interleave_resid(h, nb, prob->n_series, prob->n_epochs, rb, resid_stride,
                                     K + (size_t)n * ld, stride, stream);
            if (rc != MF_OK) return rc;
            rb = K + (size_t)n * ld;
            rs = stride;
        }
        rc = mf_potrf_loglike(h, nb, n, K, ld, stride, rb, rs, ll + b0, info + b0, stream);
        if (rc != MF_OK) return rc;
        b0 += nb;
    }
    return MF_OK;
}
