/*
 * jolideco_b200 — C ABI of the B200-native MAP-deconvolution hot path.
 *
 * The reference (jolideco/jolideco) is pure Python on PyTorch and has no FFI: the functions
 * below are what a binding for its hot path (jolideco/core.py:214-230) would call instead of the
 * ATen op stream.  Each entry cites the reference code it replaces (paths relative to the
 * reference root).  See INTEGRATION.md for the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (fp32 unless stated) borrowed for the call; images are
 *     dense row-major 2-D arrays (the reference's (1,1,H,W) tensors);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs;
 *   - return 0 on success, a negative jd_status otherwise; jd_last_error() gives the message of
 *     the last failure on the calling thread;
 *   - no hidden global state, no allocation: workspaces are caller-provided;
 *   - sm_100a only.  There is no CPU path: without a Blackwell device every call fails.
 *
 * Geometry: flux grid fH x fW (upsampled by f), counts grid H x W = fH/f x fW/f, PSF kh x kw
 * (already upsampled), patches size 8x8 (D = 64), stride s, ny = (fH-8)/s+1, nx = (fW-8)/s+1,
 * patch p = iy*nx + ix, patch element d = 8*u + v.
 */
#ifndef JOLIDECO_B200_H
#define JOLIDECO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JD_ABI_VERSION 1

typedef enum {
  JD_OK = 0,
  JD_ERR_INVALID = -1,   /* bad argument (shape, null pointer, unsupported size) */
  JD_ERR_CUDA = -2,      /* CUDA runtime error (message has cudaGetErrorString) */
  JD_ERR_NO_DEVICE = -3, /* no sm_100 device */
  JD_ERR_UNSUPPORTED = -4
} jd_status;

typedef void* jd_stream_t;

int jd_abi_version(void);
const char* jd_last_error(void);
/* 1 if `device` is a compute-capability 10.x GPU, 0 if not, <0 on error. */
int jd_device_supported(int device);

/* ---- a1: flux parameterisation (models/core.py:583-594) --------------------------------
 * flux = exp(theta) if use_log_flux else theta;  flux *= mask if mask != NULL (uint8 0/1). */
int jd_flux_forward(const float* theta, const uint8_t* mask, float* flux, int64_t n, int use_log_flux,
                    jd_stream_t stream);

/* ---- a3: PSF convolution of flux*exposure (models/npred.py:175-178, utils/torch.py:337-370)
 * conv[i,j] = sum_{a,b} psf[a,b] * g[i+sy-a, j+sx-b],  g = flux*exposure (exposure may be NULL),
 * s = (k-1)/2, zero outside: identical to rfft2*rfft2 -> irfft2 -> centred crop.
 * Direct shared-memory-tiled form, any PSF size. */
int jd_conv_forward_direct(const float* flux, const float* exposure, const float* psf, float* conv, int fH,
                           int fW, int kh, int kw, jd_stream_t stream);

/* Adjoint of the above w.r.t. flux, fused with the sum-pool adjoint (replicate f x f) and the
 * exposure product:  dflux[m,n] (+)= E[m,n] * sum_{a,b} psf[a,b] * dc[m-sy+a, n-sx+b],
 * dc[y,x] = dpool[y/f, x/f] (0 outside H x W).  accumulate != 0 adds into dflux. */
int jd_conv_backward_direct(const float* dpool, const float* exposure, const float* psf, float* dflux,
                            int accumulate, int fH, int fW, int kh, int kw, int f, int H, int W,
                            jd_stream_t stream);

/* Tuning / test hook of the direct kernels (process-wide; 0 = automatic): v3 = 0 selects the chunked,
 * synchronously staged kernel instead of the cp.async one; tx = 8 | 16 | 28 forces 32- | 64-column tiles | 64-column tiles with two column groups per thread;
 * split = 1 | 2 | 4 forces the number of thread groups the PSF rows are divided over. */
int jd_conv_tuning(int v3, int tx, int split);

/* ---- a3, large PSFs: shared-memory FFT convolution (same arithmetic contract as the direct entries) ----
 * jd_fftconv_sizes: element counts (floats) of the cached PSF spectrum and of the scratch workspace for an
 * fH x fW image and a kh x kw PSF (both axes padded to the next r * 2^L >= n + k - 1, r in {1, 3, 5}).
 * jd_fftconv_prepare_psf: PSF spectrum, computed once per dataset (the reference recomputes rfft2(psf) on every
 * call, utils/torch.py:368).  jd_conv_forward_fft / jd_conv_backward_fft: drop-in for the *_direct entries. */
int jd_fftconv_sizes(int fH, int fW, int kh, int kw, int64_t* psf_hat_elems, int64_t* workspace_elems);
int jd_fftconv_prepare_psf(const float* psf, int kh, int kw, int fH, int fW, float* psf_hat, float* workspace,
                           jd_stream_t stream);
int jd_conv_forward_fft(const float* flux, const float* exposure, const float* psf_hat, float* workspace,
                        float* conv, int fH, int fW, int kh, int kw, jd_stream_t stream);
int jd_conv_backward_fft(const float* dpool, const float* exposure, const float* psf_hat, float* workspace,
                         float* dflux, int accumulate, int fH, int fW, int kh, int kw, int f, int H, int W,
                         jd_stream_t stream);

/* Pre-clip sum-pool alone (F.avg_pool2d(kernel_size=f, divisor_override=1), models/npred.py:181-184):
 * pool[I,J] = sum_{u,v<f} conv[f I+u, f J+v]; conv has row stride fW.  Used by the autograd binding of
 * NPredModel.forward; the fused step uses jd_poisson_forward_backward instead. */
int jd_pool_sum(const float* conv, float* pool, int H, int W, int f, int fW, jd_stream_t stream);

/* ---- a3/a4/a6: sum-pool + clip + background + Poisson cash statistic and its gradient ------
 * (models/npred.py:181-191, 234-261; loss.py:35-37 = nn.PoissonNLLLoss(log_input=False,
 * reduction="mean", eps=1e-25, full=True))
 *   pool  = sum over f x f blocks of conv;  npred = max(pool,0) + B * exp(*bkg_log_norm)
 *   loss_sum[0] += sum_pix [ npred - c log(npred+eps) + 1{c>1}(c log c - c + .5 log(2 pi c)) ]
 *   dpool = grad_scale * (1 - c/(npred+eps)) * 1{pool >= 0}        (grad_scale = 1/(H W) for the mean)
 *   dlogb[0] += sum_pix grad_scale * (1 - c/(npred+eps)) * B * exp(*bkg_log_norm)
 * npred, dpool, dlogb, bkg_log_norm may be NULL.  loss_sum / dlogb are double accumulators. */
int jd_poisson_forward_backward(const float* conv, const float* background, const float* bkg_log_norm,
                                const float* counts, float* npred, float* dpool, double* loss_sum,
                                double* dlogb, int H, int W, int f, int fW, float eps, float grad_scale,
                                jd_stream_t stream);

/* ---- a1..a7 batched: the likelihood of EVERY dataset of a joint iteration in one launch per direction ----
 * (models/npred.py:160-191, 210-261; loss.py:35-37, 257-261).  One table entry per dataset, resident in device
 * memory; all datasets of a call share the geometry (flux grid fH x fW, PSF kh x kw, f in {1, 2}).
 * forward : conv = PSF (*) (flux . exposure) -> f x f sum-pool -> clip -> + B exp(*bkg_log_norm) -> Poisson cash
 *           statistic and its gradient, all in the convolution epilogue (conv / npred never reach memory):
 *             *loss_sum += sum_pix [npred - c log(npred + eps)] + loss_const
 *             dpool      = grad_scale (1 - c / (npred + eps)) 1{pool >= 0}         (skipped when dpool == NULL)
 *             *dlogb    += sum_pix grad_scale (1 - c / (npred + eps)) B exp(*bkg_log_norm)
 *           loss_const = sum_pix 1{c > 1} (c log c - c + 1/2 log(2 pi c)), the counts-only Stirling term of
 *           nn.PoissonNLLLoss(full=True), computed once per dataset by the caller.
 * backward: dflux (+)= exposure . (PSF (*)^T up_f(dpool))   (+= when accumulate != 0).
 * Same direct correlation as jd_conv_*_direct (asymmetric crop of even PSFs included). */
typedef struct {
  const float* flux;         /* NPred input: the flux, or this dataset's shifted flux; fH x fW */
  const float* exposure;     /* fH x fW, may be NULL */
  const float* psf;          /* kh x kw */
  const float* background;   /* H x W */
  const float* counts;       /* H x W */
  const float* bkg_log_norm; /* 1 float or NULL (NPredCalibration._background_norm) */
  float* dpool;              /* H x W: forward output / backward input; NULL = loss only */
  double* loss_sum;          /* accumulator or NULL */
  double* dlogb;             /* accumulator or NULL */
  float* dflux;              /* fH x fW: backward output */
  double loss_const;
  int32_t accumulate;
  int32_t reserved;
} jd_lik_dataset;

/* 1 if (kh, kw, f) is covered by the batched direct kernels (PSF rows of <= 37..40 taps, f in {1, 2}) */
int jd_likelihood_supported(int kh, int kw, int f);
int jd_likelihood_forward(const jd_lik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                          int H, int W, float eps, float grad_scale, jd_stream_t stream);
int jd_likelihood_backward(const jd_lik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                           int H, int W, jd_stream_t stream);

/* The same two entry points for large PSFs, on the shared-memory FFT path (csrc/jd_fft.cu) - every dataset of the step
 * per launch (grid.y): row FFTs -> column FFT . PSF^ . inverse column FFT -> inverse row FFTs with the sum-pool (f = 1, 2)
 * and the Poisson statistic + gradient fused into the last pass (the convolution is never written); the adjoint the same
 * three passes with conj(PSF^).  Record = jd_lik_dataset (its `psf` field is not read) + the dataset's cached PSF spectrum
 * (jd_fftconv_prepare_psf) and its own spectrum workspace (jd_fftconv_sizes). */
typedef struct {
  jd_lik_dataset lik;
  float* workspace;     /* jd_fftconv_sizes(...).workspace_elems floats, private to the dataset */
  const float* psf_hat; /* jd_fftconv_prepare_psf */
} jd_fftlik_dataset;
int jd_likelihood_forward_fft(const jd_fftlik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                              int H, int W, float eps, float grad_scale, jd_stream_t stream);
int jd_likelihood_backward_fft(const jd_fftlik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                               int H, int W, jd_stream_t stream);

/* Measurement helper (bench.py): launches a pure FP32-FMA kernel (8 independent chains per thread, 8 x 256 threads
 * per SM) and returns the number of floating-point operations it executes, or < 0 on a launch error.  Timed by the
 * caller with CUDA events, it gives the FP32 pipe peak the direct convolution is compared with. */
int64_t jd_probe_fp32_fma(int iters, float* out, jd_stream_t stream);

/* ---- a9: GMM log-probabilities of explicit feature vectors (priors/patches/gmm.py:262-281) --
 * logp[p,k] = -0.5 * sum_j (x_p . Lw[k][:,j] - mw[k][j])^2 + ck[k]
 * with the packed constants  Lw[k] = L_k diag(sqrt(w)),  mw[k] = (mu_k L_k) sqrt(w),
 * ck[k] = -D/2 log 2pi + sum_j log L_k[j,j] + log pi_k.  Any D <= 1024. */
int jd_gmm_log_prob(const float* x, int64_t P, int D, int K, const float* Lw, const float* mw,
                    const float* ck, float* logp, jd_stream_t stream);

/* ---- a8: cycle-spin roll + overlapping 8x8 patches (utils/torch.py:91-119, 226-275) ----------
 * X[p', 8u+v] = flux[(s*iy+u-sy) mod fH, (s*ix+v-sx) mod fW] for patch rows iy in [row_begin,row_end).
 * Integer-exact; shares its index function with every other patch kernel. */
int jd_extract_patches(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                       int row_end, float* X, jd_stream_t stream);

/* ---- a8+a9+a10: GMM patch prior on the flux image (priors/patches/core.py:189-246) ---------
 * Gathers 8x8 patches of the rolled image r[y,x] = flux[(y-sy) mod fH, (x-sx) mod fW] for patch
 * rows iy in [row_begin,row_end), subtracts the patch mean, evaluates all K components and
 * reduces v_p = max_k (marginalize=0) or logsumexp_k (marginalize=1).  shift_yx: 2 int32 on the
 * device (sy, sx).  Outputs (local patch index p' = (iy-row_begin)*nx + ix):
 *   value[p'] (v_p), argmax[p'] (int32), sum[0] += sum_p v_p (double), logp (optional, P' x K,
 *   needed by the marginalize=1 backward).  Patches containing NaN or values <= -1e5 are skipped
 *   (value 0, argmax -1), as the reference filters them (core.py:215-216).
 * backend: 0 = FP32 CUDA cores (check path; the tensor-core path is jd_gmm_prior_forward_tc). */
int jd_gmm_prior_forward(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                         int row_begin, int row_end, const float* Lw, const float* mw, const float* ck,
                         int K, int marginalize, float* value, int32_t* argmax, float* logp, double* sum,
                         int backend, jd_stream_t stream);

/* tcgen05 (5th-gen tensor core) forward, split-TF32 with FP32 accumulation in TMEM.  Same contract
 * and outputs as jd_gmm_prior_forward; the component matrices are passed as the pre-packed operand
 * image Bt built once by jd_gmm_tc_pack from Lw (K x 64 x 64): per component 32 KB holding Lw_k^T
 * split into TF32 hi/lo halves in the 128B-swizzled K-major shared-memory layout the MMA reads.
 * upper_tri != 0 asserts that every Lw_k is upper triangular (true for precision Cholesky factors,
 * utils/numpy.py:16-34): the kernel then skips the structurally-zero part of the product.
 * zero_mean != 0 asserts mw == 0 (zero-mean mixtures such as zoran-weiss): the epilogue skips the
 * mean subtraction.
 * NOTE: the tensor-core forwards write logp COMPONENT-MAJOR (K x P', coalesced over patch rows), unlike the
 * CUDA-core forward (P' x K); jd_gmm_prior_backward_lse_tc consumes that layout. */
size_t jd_gmm_tc_packed_bytes(int K);
int jd_gmm_tc_pack(const float* Lw, int K, void* Bt, jd_stream_t stream);
int jd_gmm_prior_forward_tc(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                            int row_begin, int row_end, const void* Bt, const float* mw, const float* ck,
                            int K, int upper_tri, int zero_mean, int marginalize, float* value,
                            int32_t* argmax, float* logp, double* sum, jd_stream_t stream);

/* Stream-K variant of jd_gmm_prior_forward_tc: the (patch tile, component) space is cut into equal chunks, one
 * per CTA pair, so that all SMs stream the same number of components whatever the patch count is (row-block
 * shards, image sizes that do not fill the last wave).  `workspace`: jd_gmm_tc_sk_workspace_bytes(P', K) bytes,
 * 256-byte aligned, ZERO-INITIALISED ONCE by the caller (arrival counters; the kernel leaves them at zero),
 * not shared between launches that may run concurrently.  Same outputs as jd_gmm_prior_forward_tc. */
int64_t jd_gmm_tc_sk_workspace_bytes(int64_t n_patches, int K);
/* The decomposition the launch will use (host arithmetic only): CTA pairs, components per chunk, partial slots per tile. */
int jd_gmm_tc_sk_plan(int64_t n_patches, int K, int* n_cta_pairs, int* chunk, int* max_segments_per_tile);
int jd_gmm_prior_forward_tc_sk(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                               int row_end, const void* Bt, const float* mw, const float* ck, int K, int upper_tri,
                               int zero_mean, int marginalize, void* workspace, float* value, int32_t* argmax,
                               float* logp, double* sum, jd_stream_t stream);

/* logsumexp (marginalize=1) backward on the tensor cores: G[p',:] = scale * sum_k r[p',k] (xc_p Lam_k - bk_k) minus
 * its row mean, r = exp(logpT[k,p'] - lse[p']).  Bt_lam = jd_gmm_tc_pack(Lam) (any 64x64 matrices pack), logpT and
 * lse (= value) come from a tensor-core forward with marginalize=1. */
int jd_gmm_prior_backward_lse_tc(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                                 int row_begin, int row_end, const void* Bt_lam, const float* bk, int K,
                                 const float* logpT, const float* lse, float scale, float* G, jd_stream_t stream);

/* Same forward with SPLIT-FP16 operands (kind::f16, FP32 accumulation): hi/lo FP16 pairs with power-of-two
 * row / component scales carry the same 22 significand bits as the TF32 pairs at half the operand bytes and
 * twice the tensor rate.  Bt: K x 16 KB image from jd_gmm_tc16_pack, binv[k] = 1 / component scale. */
size_t jd_gmm_tc16_packed_bytes(int K);
int jd_gmm_tc16_pack(const float* Lw, int K, void* Bt, float* binv, jd_stream_t stream);
int jd_gmm_prior_forward_tc16(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                              int row_begin, int row_end, const void* Bt, const float* binv, const float* mw,
                              const float* ck, int K, int upper_tri, int zero_mean, int marginalize,
                              float* value, int32_t* argmax, float* logp, double* sum, jd_stream_t stream);

/* ---- a8..a10 forward, third tensor-core kernel (csrc/jd_gmm_tcm.cu): split-TF32 with the two correction products on
 * the FP16 pipe (same 2^-21 accuracy as 3 x TF32, two thirds of the tensor work), persistent stream-K decomposition,
 * gather of the next tile overlapped with the MMAs of the current one.  Same contract as jd_gmm_prior_forward_tc_sk;
 * Bt / binv come from jd_gmm_tcm_pack (scaled operand image + 1 / component scale), `workspace` holds
 * jd_gmm_tcm_workspace_bytes(n_patches, K) zero-initialised bytes (256-byte aligned; self-resetting counters). */
size_t jd_gmm_tcm_packed_bytes(int K);
int jd_gmm_tcm_pack(const float* Lw, int K, void* Bt, float* binv, jd_stream_t stream);
int64_t jd_gmm_tcm_workspace_bytes(int64_t n_patches, int K);
int jd_gmm_prior_forward_tcm(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                             int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                             int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                             int32_t* argmax, float* logp, double* sum, jd_stream_t stream);

/* ---- a8..a10 forward, fourth tensor-core kernel (csrc/jd_gmm_tcm2.cu): two patch tiles per CTA against every staged
 * operand image, one issuer warp + epilogue group + private accumulator slots per tile, a stripped issue loop.  Same
 * contract as jd_gmm_prior_forward_tcm, in two precision recipes of equal accuracy (22 significand bits per operand):
 *   jd_gmm_prior_forward_tcm2   - mixed TF32 / FP16 split, Bt / binv from jd_gmm_tcm_pack;
 *   jd_gmm_prior_forward_tc16x2 - split FP16, Bt / binv from jd_gmm_tc16_pack (12 instead of 16 MMAs per tile and
 *                                 component, three accumulator slots per tile, half the operand bytes).
 * `workspace`: jd_gmm_tcm2_workspace_bytes(n_patches, K) zero-initialised bytes, 256-byte aligned (both recipes). */
int64_t jd_gmm_tcm2_workspace_bytes(int64_t n_patches, int K);
int jd_gmm_prior_forward_tcm2(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                              int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                              int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                              int32_t* argmax, float* logp, double* sum, jd_stream_t stream);
int jd_gmm_prior_forward_tc16x2(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                                int row_end, const void* Bt16, const float* binv, const float* mw, const float* ck,
                                int K, int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                                int32_t* argmax, float* logp, double* sum, jd_stream_t stream);

/* The two-tile forward (recipe 0 = jd_gmm_prior_forward_tcm2, 1 = jd_gmm_prior_forward_tc16x2, same arguments) on at most
 * `clusters` CTA pairs (0 = every SM pair): the kernel owns the SMs it runs on, so a step whose likelihood chain is
 * short (one dataset) runs that chain on the remaining SMs from a second stream.  Max / argmax results do not depend on
 * `clusters`; logsumexp values can differ in the last bits (the partial sums of a split tile are merged per chunk).  The
 * workspace of jd_gmm_tcm2_workspace_bytes is large enough for every cluster count. */
int jd_gmm_prior_forward_tcx2_on(int recipe, int clusters, const float* flux, int fH, int fW, const int32_t* shift_yx,
                                 int stride, int row_begin, int row_end, const void* Bt, const float* binv,
                                 const float* mw, const float* ck, int K, int upper_tri, int zero_mean, int marginalize,
                                 void* workspace, float* value, int32_t* argmax, float* logp, double* sum,
                                 jd_stream_t stream);

/* Per-patch gradient  G[p',:] = scale * sum_k R[p',k] (xc_p Lam_k - bk_k),  minus its row mean,
 * R = one-hot(argmax) or softmax_k(logp) (marginalize=1; needs logp and value from the forward);
 * Lam_k = Lw_k Lw_k^T, bk_k = mw_k Lw_k^T.  (Autograd mirror of gmm.py:270-272 + norms.py:97-103.)
 * For d prior/d flux pass scale = -stride^2/64/(fH fW).
 * workspace (optional, marginalize=0): jd_gmm_backward_workspace_elems(P', K) int32, ZERO-INITIALISED before the first
 * use (the kernels leave it ready for the next launch); when given, patches are bucketed by winning component and
 * each Lam_k is staged in shared memory once per <= 64 patches (a register-tiled 64x64x64 product) instead of being
 * re-read from L2 for every patch - the faster path from a few ten thousand patches on.  K <= 4096. */
int64_t jd_gmm_backward_workspace_elems(int64_t P, int K);
int jd_gmm_prior_backward(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                          int row_begin, int row_end, const float* Lam, const float* bk, int K,
                          int marginalize, const int32_t* argmax, const float* logp, const float* value,
                          float scale, float* G, int32_t* workspace, jd_stream_t stream);

/* Max-mode backward for upper-triangular factors (precision Cholesky factors are): same G as
 * jd_gmm_prior_backward(marginalize = 0), computed as (xc Lw_k* - mw_k*) Lw_k*^T from the triangular factor
 * (12 KB of L2 reads per patch instead of the 16 KB of Lam_k*).  Lw must be upper triangular. */
int jd_gmm_prior_backward_max_tri(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                                  int row_end, const float* Lw, const float* mw, int K, const int32_t* argmax,
                                  float scale, float* G, jd_stream_t stream);

/* col2im of the per-patch gradients, gather form (deterministic, no atomics): for every pixel
 * sum the <= (8/s)^2 patch entries covering it at the rolled coordinates
 * (unfold_backward x2 + roll backward).  dflux (+)= fold(G) if accumulate else = fold(G). */
int jd_patch_fold(const float* G, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                  int row_end, float* dflux, int accumulate, jd_stream_t stream);

/* ---- a12: fused gradient sum + chain rule + Adam (torch.optim.Adam defaults, core.py:39-42,229)
 * g = (dflux_a + scale_b * dflux_b) * dflux/dtheta,  dflux/dtheta = flux (log param) or mask;
 * m,v,theta updated in place with bias correction for step number `step` (1-based).
 * dflux_b may be NULL. */
int jd_adam_step(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                 const float* dflux_a, const float* dflux_b, float scale_b, int use_log_flux, int64_t n,
                 int step, float lr, float beta1, float beta2, float eps, jd_stream_t stream);

/* Same update with the bias-correction scalars read from device memory (adam_scalars[0] = lr/(1-b1^t),
 * adam_scalars[1] = sqrt(1-b2^t), written by jd_step_begin) so that the step can live in a CUDA graph. */
int jd_adam_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                     const float* dflux_a, const float* dflux_b, float scale_b, int use_log_flux, int64_t n,
                     const float* adam_scalars, float beta1, float beta2, float eps, jd_stream_t stream);

/* ---- device-side step bookkeeping (replaces the host-side state of core.py:214-229: the two
 * torch.randint draws of utils/torch.py:108-116 and torch.optim.Adam's step counter) --------------
 * counters[0] = cycle-spin draws consumed, counters[1] = Adam step t (both int32 on the device).
 * If shift_table != NULL: shift_out[0..1] = shift_table[counters[0]] (clamped to n_shifts-1), counters[0]++.
 * If advance_adam: counters[1]++ and adam_scalars = {lr/(1-b1^t), sqrt(1-b2^t)}.
 * zero_acc[0..n_acc) (double loss accumulators) are cleared. */
int jd_step_begin(int32_t* counters, const int32_t* shift_table, int n_shifts, int32_t* shift_out,
                  int advance_adam, float lr, float beta1, float beta2, float* adam_scalars, double* zero_acc,
                  int n_acc, jd_stream_t stream);

/* ---- a5 / 8f-2: sub-pixel shift of the flux by an NPredCalibration (utils/torch.py:196-223 `shift_image_torch` =
 * affine_grid + grid_sample, bilinear, zeros padding; npred.py:226-230 applies it with scale = upsampling factor).
 * shift_xy: device pair (shift_x, shift_y) - the storage of NPredCalibration.shift_xy.  shifted[i,j] samples flux at
 * (i + scale shift_y, j + scale shift_x).  jd_shift_backward: dflux (+)= shift^T dshifted and, when dshift_xy != NULL,
 * dshift_xy[0..1] += (dL/dshift_x, dL/dshift_y) (double accumulators, zeroed by the caller).  The reference skips the
 * operator (and never trains the shift) when both shifts are ~0 (utils/torch.py:211): callers do the same. */
int jd_shift_forward(const float* flux, const float* shift_xy, int scale, int fH, int fW, float* shifted,
                     jd_stream_t stream);
int jd_shift_backward(const float* dshifted, const float* flux, const float* shift_xy, int scale, int fH, int fW,
                      float* dflux, int accumulate, double* dshift_xy, jd_stream_t stream);

/* Adam on the scalar calibration parameters of one dataset (NPredCalibration._background_norm, models/npred.py:
 * 298-333): grad[i] are the double accumulators written by jd_poisson_forward_backward (dlogb), counter is the
 * parameter's own step count (torch.optim.Adam only advances it when the parameter received a gradient). */
int jd_adam_scalar_step_dev(float* param, float* m, float* v, const double* grad, int32_t* counter, int n,
                            float lr, float beta1, float beta2, float eps, jd_stream_t stream);

/* ---- multi-GPU joint step: gradient all-reduce fused with Adam over NVLink peer memory ------------------
 * grad_ptrs_dev / theta_ptrs_dev: device arrays of `world` pointers to every rank's partial-gradient and theta
 * buffers (symmetric memory, peer-addressable).  Rank r sums the partial gradients of its pixel slice
 * [n r/world, n (r+1)/world) from all peers (fixed order), applies Adam there (chain rule as jd_adam_step_dev) and
 * stores the new theta slice into every replica.  The caller provides the cross-rank barriers before (all
 * partial gradients written) and after (all slices stored).  n must be a multiple of 4.
 * Replaces ncclAllReduce + jd_adam_step_dev of the NCCL variant. */
int jd_adam_allreduce_peer(const void* grad_ptrs_dev, const void* theta_ptrs_dev, int rank, int world, float* m,
                           float* v, const float* flux, const uint8_t* mask, int use_log_flux, int64_t n,
                           const float* adam_scalars, float beta1, float beta2, float eps, jd_stream_t stream);

/* The same reduce + Adam + broadcast with both cross-rank barriers inside the kernel (flag words in symmetric
 * memory, release/acquire at system scope), so that a multi-rank joint step is one CUDA graph per rank.
 * sig_ptrs_dev: device array of `world` pointers to every rank's flag block (64 uint32, zero-initialised, symmetric
 * memory); sync_state: 12 uint32 of THIS rank's device memory, 8-byte aligned, zero-initialised and never reset by the
 * caller: [0] epoch, [1] finished-CTA count, [4..11] four int64 %globaltimer stamps of the last launch (entry, entry
 * barrier passed, theta stores fenced, exit barrier passed).  Every rank must launch it the same number of times.
 * world <= 32; the peer loads are unrolled for world = 2, 4, 8. */
int jd_adam_allreduce_peer_sync(const void* grad_ptrs_dev, const void* theta_ptrs_dev, const void* sig_ptrs_dev,
                                uint32_t* sync_state, int rank, int world, float* m, float* v, const float* flux,
                                const uint8_t* mask, int use_log_flux, int64_t n, const float* adam_scalars,
                                float beta1, float beta2, float eps, jd_stream_t stream);

/* Joint iteration (TotalLoss.__call__, loss.py:257-261), gradient assembly:
 * jd_adam_joint_step_dev: g = sum_q parts[q] (n_parts images, part_stride floats apart) + scale_b * fold(G)
 *   (G may be NULL), then chain rule + Adam as jd_adam_step_dev - the whole update of a one-GPU joint step;
 * jd_grad_reduce_local : out = the same g, no update - this rank's partial gradient of a multi-GPU joint step
 *   (out may alias parts[0]). */
int jd_adam_joint_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                           const float* parts, int n_parts, int64_t part_stride, const float* G, float scale_b,
                           int use_log_flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* adam_scalars, float beta1, float beta2, float eps,
                           jd_stream_t stream);
int jd_grad_reduce_local(const float* parts, int n_parts, int64_t part_stride, const float* G, float scale_b, int fH,
                         int fW, const int32_t* shift_yx, int stride, int row_begin, int row_end, float* out,
                         jd_stream_t stream);

/* Fused variants used by the graph-captured step (one launch each instead of two):
 * jd_step_begin_flux = jd_step_begin + jd_flux_forward;
 * jd_adam_fold_step_dev = jd_patch_fold (gather col2im of G, scaled by scale_b) + jd_adam_step_dev. */
int jd_step_begin_flux(int32_t* counters, const int32_t* shift_table, int n_shifts, int32_t* shift_out,
                       int advance_adam, float lr, float beta1, float beta2, float* adam_scalars,
                       double* zero_acc, int n_acc, const float* theta, const uint8_t* mask, float* flux,
                       int64_t n, int use_log_flux, jd_stream_t stream);
int jd_adam_fold_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                          const float* dflux_a, const float* G, float scale_b, int use_log_flux, int fH, int fW,
                          const int32_t* shift_yx, int stride, int row_begin, int row_end,
                          const float* adam_scalars, float beta1, float beta2, float eps, jd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* JOLIDECO_B200_H */
