/*
 * miniframe_b200 -- C ABI of the B200-native GP log-likelihood hot path.
 *
 * The reference (jdavidrcamacho/mini-frame) is pure Python and has no FFI of
 * its own; the drop-in boundary is its class API.  Every entry point below
 * names the reference interface it replaces (file:line under
 * /root/reference/miniframe/).  The Python facade in miniframe_b200/ binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     its name ends in _host; the caller owns every data buffer and the
 *     workspace.  mf_create allocates the handle's own 2 MB of device scratch
 *     (hand-out counter and group barriers of the factorisation); no call after
 *     mf_create allocates, frees or synchronises;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*;
 *     NULL = the default stream) and is asynchronous with respect to the host;
 *   - every function returns an mf_status (0 = ok), never throws;
 *   - numerical failure is per sample: info[b] = k > 0 means the k-th pivot of
 *     sample b was not positive (LAPACK dpotrf convention) and ll[b] = -inf,
 *     which is what the reference's `except LinAlgError: return -np.inf`
 *     produces (BIGgp.py:310-311, SMALLgp.py:274-276).
 *
 * Workspace matrix layout ("packed lower + residual row")
 *   One matrix of order n = n_series * n_epochs occupies n + 1 rows (allocated as
 *   8 * ceil((n + 1) / 8) rows: the factorisation moves 8-row groups) of
 *   ld = mf_ld(n) doubles, row-major.  Rows 0..n-1 hold the covariance matrix
 *   (the lower triangle is what the factorisation reads and overwrites with
 *   L, K = L L^T); row n receives z = L^-1 r, the forward-substituted
 *   residual.  Consecutive matrices are mf_matrix_stride(n) doubles apart.
 *   The reference factors the UPPER triangle of a C-ordered K
 *   (cho_factor(..., lower=False), BIGgp.py:306); K is symmetric so the
 *   row-major lower factor here is the transpose of that U.
 */
#ifndef MINIFRAME_B200_H
#define MINIFRAME_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MF_VERSION 200

typedef enum {
    MF_OK = 0,
    MF_ERR_INVALID = 1,     /* bad argument */
    MF_ERR_CUDA = 2,        /* a CUDA call failed; see mf_last_error */
    MF_ERR_WORKSPACE = 3,   /* workspace too small */
    MF_ERR_UNSUPPORTED = 4  /* valid in the reference, not built here */
} mf_status;

/* which reference class builds the covariance */
#define MF_FRAMEWORK_BIGGP      0   /* BIGgp.py:14            */
#define MF_FRAMEWORK_BIGGP_JITT 1   /* BIGgpPlusJitt.py:14    */
#define MF_FRAMEWORK_SMALLGP    2   /* SMALLgp.py:12          */

/* kernel family, selected in the reference by class __name__
 * (BIGgp.py:112-117, SMALLgp.py:63-68) */
#define MF_KERNEL_NONE (-1)
#define MF_KERNEL_SE     0          /* kernels.py:80-259,  pars (ell, wn)              */
#define MF_KERNEL_QP     1          /* kernels.py:263-617, pars (ell_e, ell_p, P, wn)  */

#define MF_MAX_SERIES 8
#define MF_MAX_HYPER  32            /* >= 4 + 3 * MF_MAX_SERIES */
#define MF_MAX_EXTRA  16            /* >= 4 + MF_MAX_SERIES     */

/*
 * Problem descriptor (host struct, passed by pointer, copied at call time).
 *   hyper row (hyper_len doubles per sample) is the reference's `a`:
 *     BIGgp / BIGgpPlusJitt: [kernel pars | vc vr lc bc br]   (BIGgp.py:110-119, 167-169)
 *     SMALLgp:               [kernel pars | 3 * n_series scaling pars] (SMALLgp.py:61-68, 83-87)
 *   extra row (extra_len doubles per sample, may be 0):
 *     BIGgpPlusJitt: the three jitters `j`                     (BIGgpPlusJitt.py:263-269)
 *     SMALLgp:       `c` = [extra-kernel pars | n_series amplitudes] (SMALLgp.py:71-80, 152-155)
 *   nugget: 0 = off, else K <- (1-nugget) K + nugget diag(diag K)
 *           (0.01 in the reference: BIGgp.py:274-277, SMALLgp.py:245-248)
 */
typedef struct {
    int32_t framework;
    int32_t kernel;
    int32_t extra_kernel;
    int32_t n_series;
    int32_t n_epochs;
    int32_t hyper_len;
    int32_t extra_len;
    int32_t reserved;
    double  nugget;
} mf_problem_t;

typedef struct mf_handle_s* mf_handle_t;

int  mf_version(void);
const char* mf_status_string(int status);

/* One handle per (device, stream): it owns 2 MB of device scratch (the hand-out
 * counter and the group barriers of mf_potrf_loglike), a small cache of TMA
 * descriptors and the last error text, all of which concurrent launches would
 * share.  Calls on one handle must therefore be issued to one stream at a time;
 * different handles are independent (thread-safe across handles). */
int  mf_create(mf_handle_t* out, int device);
int  mf_destroy(mf_handle_t h);
const char* mf_last_error(mf_handle_t h);
int  mf_device_info(mf_handle_t h, int* sm_count, int* cc_major, int* cc_minor);

/* Tuning / test knobs of one handle (0 restores the library's own choice):
 *   "potrf_tile"   1: 64-row tiles, 2: 128-row tiles of the factorisation
 *   "potrf_group"  1: one CTA per matrix, G > 1: G CTAs per matrix (cooperative)
 *   "potrf_map"    cooperative factorisation, which CTA takes which row tile:
 *                  1: (tile - block column) mod G, 2: (tile + block column) mod G
 *   "potrf_lookahead"  one CTA per matrix: 1: the plain kernel, 2: the look-ahead kernel (the
 *                  diagonal block of a block column factorised by one warp behind the next tile's
 *                  contraction; bit-identical results); 0: whichever tiling wastes less on the
 *                  ragged tiles of this matrix order
 *   "l2_persist_mb"  L2 set-aside (MB, -1: the device maximum) for the rows the factorisation
 *                  re-reads; a DEVICE limit (cudaLimitPersistingL2CacheSize), left alone by default:
 *                  it costs the covariance build half its write bandwidth
 *   "cov_ctas"     CTAs per sample of the covariance build
 *   "experiment"   timing-experiment variants; rejected with MF_ERR_UNSUPPORTED
 *                  unless the library was built with -DMF_EXPERIMENTS (the
 *                  shipped library is not) */
int  mf_set_option(mf_handle_t h, const char* name, int value);

/* layout helpers (host arithmetic only) */
int64_t mf_ld(int n);               /* leading dimension in doubles, multiple of 16 */
int64_t mf_matrix_stride(int n);    /* 8 * ceil((n + 1) / 8) * mf_ld(n) doubles */
size_t  mf_workspace_bytes(int batch, int n);

/*
 * Covariance build: replaces compute_matrix() and everything under it
 * (BIGgp.py:103-119, 167-278; BIGgpPlusJitt.py:237-282; SMALLgp.py:54-91,
 * 139-249) together with the kernel classes it instantiates
 * (kernels.py:80-617).  One sin/cos/exp per (i, j); every block is written
 * straight into the workspace layout.
 *   t         [n_epochs]            observation times
 *   diag_add  [n]                   added to the diagonal (rverr^2 ... or yerr^2 or 1e-12)
 *   hyper     [batch x hyper_len]
 *   extra     [batch x extra_len]   or NULL when extra_len == 0
 *   mode      MF_COV_LOWER: lower triangle only, the reference's row order
 *               r = P * n_epochs + i (series P, epoch i);
 *             MF_COV_FULL: the full symmetric square (what compute_matrix returns);
 *             MF_COV_LOWER_INTERLEAVED: lower triangle in EPOCH-INTERLEAVED order
 *               r = n_series * i + P, a symmetric permutation of K that leaves the
 *               log-likelihood unchanged.  It turns the nine scattered output blocks
 *               into one dense stream of 128-byte-aligned tiles written with TMA
 *               stores, which is why mf_loglike uses it; BIGgp / BIGgpPlusJitt with
 *               nugget == 0 only.  A residual used with such a matrix must be in the
 *               same order (mf_interleave_resid).
 *   K         [batch] matrices, leading dimension ld, `stride` doubles apart
 */
#define MF_COV_LOWER              0
#define MF_COV_FULL               1
#define MF_COV_LOWER_INTERLEAVED  2
int mf_build_cov(mf_handle_t h, const mf_problem_t* prob, int batch,
                 const double* t, const double* diag_add,
                 const double* hyper, const double* extra, int mode,
                 double* K, int64_t ld, int64_t stride, void* stream);

/* dst[b][n_series * i + P] = src[b][P * n_epochs + i]: the residual y - mean
 * (BIGgp.py:303) in the order of MF_COV_LOWER_INTERLEAVED.  src_stride may be 0
 * (one residual for the whole batch); dst may be row n of the workspace matrices
 * (dst_stride = mf_matrix_stride), from where mf_potrf_loglike takes it in place. */
int mf_interleave_resid(mf_handle_t h, int batch, int n_series, int n_epochs,
                        const double* src, int64_t src_stride,
                        double* dst, int64_t dst_stride, void* stream);

/*
 * Blocked left-looking Cholesky with the residual carried as row n, fused with
 * the log-determinant and the quadratic form: replaces cho_factor + cho_solve +
 * sum(log(diag)) of log_likelihood (BIGgp.py:305-309, BIGgpPlusJitt.py:311-315,
 * SMALLgp.py:269-273).
 *   K      in: lower triangle of each matrix; out: L (lower) and, in row n, z = L^-1 r
 *   resid  [batch x n] with row stride resid_stride (0 = one residual shared by
 *          the whole batch), or NULL (z = 0, ll = -logdet - n/2 log 2pi); may be
 *          row n of the matrices themselves (resid = K + n * ld, resid_stride =
 *          stride): it is then replaced by z in place
 *   ll     [batch] log-likelihood, -inf where info != 0
 *   info   [batch]
 * One CTA per matrix when batch >= 2 x SM count; with fewer matrices several CTAs
 * share each one (cooperative launch).  The handle's scratch carries the hand-out
 * counter and the groups' synchronisation words: one handle per stream.  The
 * factor is the same, bit for bit, either way.
 */
int mf_potrf_loglike(mf_handle_t h, int batch, int n,
                     double* K, int64_t ld, int64_t stride,
                     const double* resid, int64_t resid_stride,
                     double* ll, int32_t* info, void* stream);

/*
 * Whole hot path for a batch: replaces log_likelihood() called once per
 * hyper-parameter set (BIGgp.py:281-312, BIGgpPlusJitt.py:285-318,
 * SMALLgp.py:252-277; the caller being examples/BIGgp_mcmc.py:47-64).
 * Processes the batch in chunks of floor(ws_bytes / matrix bytes) matrices.
 * resid is given in the reference's order; for BIGgp / BIGgpPlusJitt the
 * covariance is built and factorised in the epoch-interleaved order (the
 * log-likelihood does not depend on it) and info[b] then counts pivots in that
 * order.  The reference only ever tests info for zero (LinAlgError -> -inf).
 */
int mf_loglike(mf_handle_t h, const mf_problem_t* prob, int batch,
               const double* t, const double* diag_add,
               const double* hyper, const double* extra,
               const double* resid, int64_t resid_stride,
               void* workspace, size_t ws_bytes,
               double* ll, int32_t* info, void* stream);

/*
 * Stand-alone fused forward substitution + log-determinant on an existing
 * factor (the cho_solve / log-diag tail, BIGgp.py:307-309), for callers that
 * reuse one factor with several residuals (e.g. sampling mean parameters only).
 *   L      lower factor in the workspace layout (row n is not read); `stride` may be 0:
 *          ONE factor for all `batch` right-hand sides (used by predict_gp for L^-1 K*^T,
 *          BIGgp.py:635-637)
 *   z      [batch x n] out (row stride n), required (it is the kernel's working vector)
 *   info   factorisation status per sample or NULL; ll = -inf where info != 0
 */
int mf_solve_logdet(mf_handle_t h, int batch, int n,
                    const double* L, int64_t ld, int64_t stride,
                    const double* resid, int64_t resid_stride,
                    double* z, double* ll, const int32_t* info, void* stream);

/*
 * Keplerian mean function for a batch of parameter rows, the expensive part of the residual
 * y - mean(b) (BIGgp.py:303; means.py:143-216: the reference's fixed-count iteration for the
 * eccentric anomaly, `iterations` = 100, 500 for planet b of TOI175keplerians).
 *   pars  [batch x 5]  (P, K, e, w, T0) per row (T0 = t[0] - P phi / 2pi for the phase form)
 *   out   [batch x n_epochs], rows out_stride doubles apart; accumulate != 0 adds to it
 */
int mf_kepler_rv(mf_handle_t h, int batch, int n_epochs, const double* t, const double* pars,
                 int iterations, double* out, int64_t out_stride, int accumulate, void* stream);

/*
 * Gradient of the log-likelihood with respect to the hyper-parameters (not in the
 * reference; the consumer would be an optimiser or HMC on minus_log_likelihood,
 * BIGgp.py:315-317):  grad[b][k] = -1/2 sum_{R,C} W_b[R][C] dK_b[R][C]/dtheta_k  with
 * W = K^-1 - alpha alpha^T, alpha = K^-1 r.  The caller supplies W (e.g. from the
 * factor mf_loglike leaves in the workspace); dK/dtheta is evaluated on the fly from t
 * and the hyper-parameters and never stored.  BIGgp / BIGgpPlusJitt.
 *   W        [batch] symmetric matrices in the EPOCH-INTERLEAVED order (row 3 i + P),
 *            leading dimension ld, `stride` doubles apart; only the lower triangle is read
 *   partial  scratch, batch x mf_grad_ctas(n_epochs) x 24 doubles
 *   grad     [batch x (hyper_len + extra_len)]: d ll / d(kernel pars, vc vr lc bc br, jitters)
 */
int mf_grad_ctas(int n_epochs);
int mf_grad_contract(mf_handle_t h, const mf_problem_t* prob, int batch,
                     const double* t, const double* hyper, const double* extra,
                     const double* W, int64_t ld, int64_t stride,
                     double* partial, double* grad, void* stream);

/*
 * Diagnostics: register-resident FP64 throughput probes used as roofline
 * denominators (mode 0: mma.sync m8n8k4 f64 "DMMA", mode 1: DFMA).  Runs
 * `iters` inner iterations on every SM, returns the elapsed milliseconds of the
 * launch and the floating-point operations it executed.
 */
int mf_probe_fp64(mf_handle_t h, int mode, int iters, float* ms_host, double* flops_host,
                  void* stream);   /* times itself: creates events and synchronises the stream */

#ifdef __cplusplus
}
#endif
#endif /* MINIFRAME_B200_H */
