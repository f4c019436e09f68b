// Shared-memory FFT convolution of the NPred forward model and its adjoint (large PSFs).
//
// Same result as utils/torch.py:347-370 (rfft2 * rfft2 -> irfft2 -> centred crop) and, for the adjoint, as its
// autograd mirror, with three differences in how the work is organised:
//   * the PSF spectrum is computed ONCE per dataset (jd_fftconv_prepare_psf); the reference recomputes it on
//     every call;
//   * any FFT size >= the linear-convolution support gives the same answer, so both axes are padded to the
//     next r0 * 2^L, r0 in {1, 3, 5} (545 -> 640 instead of 1024, 1224 -> 1280 instead of 2048; JD_FFT_MIXED=0: r0 = 1)
//     and a Stockham autosort FFT (radix-4 stages, one radix-2 stage for odd L, one final twiddle-free radix-r0
//     stage; jd_fft_stages.cuh, checked on the host by tests/test_fft_host.py) runs entirely in shared memory;
//   * three kernels instead of rfft2 / multiply / irfft2 / slice:
//       rows   : [input transform] 2 real rows per complex FFT (even/odd split), half spectrum written
//                transposed  specT[kx][row]
//       cols   : per kx: column FFT . PSF^ (or conj PSF^ for the adjoint) . inverse column FFT fused, only the
//                cropped rows are written back (in place)
//       rows^-1: Hermitian rebuild, inverse FFT of 2 rows at a time, column crop and the output transform
//                (x exposure, accumulate) fused.
//   Crop / adjoint geometry: output index i reads circular index (i + off) mod S with off = +(k-1)/2 for the
//   forward model and off = -(k-1)/2 with conj(PSF^) for the adjoint (correlation), which reproduces the
//   asymmetric crop of even-sized PSFs exactly.
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_fft_stages.cuh"
#include "jd_likelihood.cuh"  // lik::poisson_px: the fused forward epilogue evaluates the same statistic

namespace jd {
namespace fft {

// tw[j] = exp(-2 pi i j / N), j < N/2
__device__ __forceinline__ void make_twiddles(float2* tw, int N) {
  for (int j = threadIdx.x; j < N / 2; j += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)j / (float)N, &s, &c);
    tw[j] = make_float2(c, -s);
  }
}

// Stockham autosort FFT of length N = r0 * 2^L in shared memory (ping-pong x <-> y).  INV conjugates the twiddles (no
// scaling).  Returns the buffer holding the result (natural order).  All threads of the block must call it; x must be
// fully written and synchronised.
template <bool INV>
__device__ __forceinline__ float2* fft_smem(float2* x, float2* y, const float2* tw, int N, int L, int r0) {
  return fft_mixed<INV>(x, y, tw, N, L, r0, [](int items, auto f) {
    for (int t = threadIdx.x; t < items; t += blockDim.x) f(t);
    __syncthreads();
  });
}

struct Plan {
  int fH, fW, Sy, Sx, logSy, logSx, ry, rx, ld;  // S = r * 2^log; ld = row stride (in complex) of specT: fH rounded up to 8
};

__device__ __forceinline__ int wrap_index(int i, int S) { return i >= S ? i - S : i; }  // i in [0, 2S)

enum { IN_FLUX = 0, IN_DPOOL = 1, IN_PSF = 2 };

// ---- pass 1: row FFTs.  CTA handles row pair (2 rows) per iteration, `pairs` pairs per CTA.
// Batched launches (jd_likelihood_*_fft): `table` != NULL, blockIdx.y = dataset, the per-dataset pointers come from its
// record.
template <int MODE>
__global__ void rows_fwd_kernel(const float* __restrict__ in, const float* __restrict__ scale, Plan pl, int nrows,
                                int ncols, int f, int H, int W, float2* __restrict__ specT, int pairs,
                                const jd_fftlik_dataset* __restrict__ table) {
  if (table) {
    const jd_fftlik_dataset& ds = table[blockIdx.y];
    in = MODE == IN_FLUX ? ds.lik.flux : ds.lik.dpool;
    scale = MODE == IN_FLUX ? ds.lik.exposure : nullptr;
    specT = reinterpret_cast<float2*>(ds.workspace);
  }
  extern __shared__ __align__(16) float2 sm[];
  float2* bx = sm;
  float2* by = sm + pl.Sx;
  float2* tw = sm + 2 * pl.Sx;
  make_twiddles(tw, pl.Sx);
  for (int it = 0; it < pairs; ++it) {
    const int ra = (blockIdx.x * pairs + it) * 2, rb = ra + 1;
    if (ra >= nrows) break;
    __syncthreads();
    for (int j = threadIdx.x; j < pl.Sx; j += blockDim.x) {
      float va = 0.f, vb = 0.f;
      if (j < ncols) {
        if (MODE == IN_FLUX) {
          va = in[(int64_t)ra * ncols + j] * (scale ? scale[(int64_t)ra * ncols + j] : 1.f);
          if (rb < nrows) vb = in[(int64_t)rb * ncols + j] * (scale ? scale[(int64_t)rb * ncols + j] : 1.f);
        } else if (MODE == IN_DPOOL) {
          int px = j / f;
          if (px < W) {
            if (ra / f < H) va = in[(int64_t)(ra / f) * W + px];
            if (rb < nrows && rb / f < H) vb = in[(int64_t)(rb / f) * W + px];
          }
        } else {
          va = in[(int64_t)ra * ncols + j];
          if (rb < nrows) vb = in[(int64_t)rb * ncols + j];
        }
      }
      bx[j] = make_float2(va, vb);
    }
    __syncthreads();
    float2* z = fft_smem<false>(bx, by, tw, pl.Sx, pl.logSx, pl.rx);
    // split: A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2i)
    for (int k = threadIdx.x; k <= pl.Sx / 2; k += blockDim.x) {
      const float2 zk = z[k], zn = z[wrap_index(pl.Sx - k, pl.Sx)];
      const float2 A = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      const float2 B = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
      float2* dst = specT + (int64_t)k * pl.ld + ra;
      dst[0] = A;
      if (rb < nrows) dst[1] = B;
    }
  }
}

// ---- pass 2: per kx column FFT, spectrum product, inverse column FFT, cropped write-back (in place).
// PSF_MODE: 0 = multiply by psf_hat, 1 = multiply by conj(psf_hat), 2 = no product (PSF spectrum setup: forward only)
template <int PSF_MODE>
__global__ void cols_kernel(float2* __restrict__ specT, const float2* __restrict__ psf_hat, Plan pl, int nrows_in,
                            int nrows_out, int off, float norm, float2* __restrict__ psf_out,
                            const jd_fftlik_dataset* __restrict__ table) {
  if (table) {
    const jd_fftlik_dataset& ds = table[blockIdx.y];
    specT = reinterpret_cast<float2*>(ds.workspace);
    psf_hat = reinterpret_cast<const float2*>(ds.psf_hat);
  }
  extern __shared__ __align__(16) float2 sm[];
  float2* bx = sm;
  float2* by = sm + pl.Sy;
  float2* tw = sm + 2 * pl.Sy;
  make_twiddles(tw, pl.Sy);
  const int k = blockIdx.x;
  float2* col = specT + (int64_t)k * pl.ld;
  for (int r = threadIdx.x; r < pl.Sy; r += blockDim.x) bx[r] = r < nrows_in ? col[r] : make_float2(0.f, 0.f);
  __syncthreads();
  float2* z = fft_smem<false>(bx, by, tw, pl.Sy, pl.logSy, pl.ry);
  if (PSF_MODE == 2) {
    float2* dst = psf_out + (int64_t)k * pl.Sy;
    for (int r = threadIdx.x; r < pl.Sy; r += blockDim.x) dst[r] = z[r];
    return;
  }
  const float2* ph = psf_hat + (int64_t)k * pl.Sy;
  for (int r = threadIdx.x; r < pl.Sy; r += blockDim.x) {
    float2 p = ph[r];
    if (PSF_MODE == 1) p.y = -p.y;
    const float2 v = cmul(z[r], p);
    z[r] = make_float2(v.x * norm, v.y * norm);
  }
  __syncthreads();
  float2* other = z == bx ? by : bx;
  float2* w = fft_smem<true>(z, other, tw, pl.Sy, pl.logSy, pl.ry);
  for (int i = threadIdx.x; i < nrows_out; i += blockDim.x) col[i] = w[wrap_index(i + off, pl.Sy)];
}

// ---- pass 3: inverse row FFTs of row pairs, column crop, output transform.
// OUT_MODE 0: out = value;  1: out (+)= value * scale;  2 (batched forward only): sum-pool f x f (f = 1, 2: the two
// rows of a pair are one pooling pair), Poisson statistic + gradient of the pooled pixel (lik::poisson_px, as the
// direct kernels' epilogue) - the convolution itself is never written
template <int OUT_MODE>
__global__ void rows_inv_kernel(const float2* __restrict__ specT, Plan pl, int nrows, int ncols, int off,
                                const float* __restrict__ scale, float* __restrict__ out, int accumulate, int pairs,
                                const jd_fftlik_dataset* __restrict__ table, int f, int H, int W, float eps,
                                float grad_scale) {
  const jd_fftlik_dataset* ds = table ? table + blockIdx.y : nullptr;
  if (ds) {
    specT = reinterpret_cast<const float2*>(ds->workspace);
    if (OUT_MODE == 1) {
      out = ds->lik.dflux;
      scale = ds->lik.exposure;
      accumulate = ds->lik.accumulate;
    }
  }
  float lacc = 0.f, bacc = 0.f;
  extern __shared__ __align__(16) float2 sm[];
  float2* bx = sm;
  float2* by = sm + pl.Sx;
  float2* tw = sm + 2 * pl.Sx;
  make_twiddles(tw, pl.Sx);
  for (int it = 0; it < pairs; ++it) {
    const int ra = (blockIdx.x * pairs + it) * 2, rb = ra + 1;
    if (ra >= nrows) break;
    __syncthreads();
    // Z[k] = A[k] + i B[k];  Z[N-k] = conj(A[k]) + i conj(B[k])
    for (int k = threadIdx.x; k <= pl.Sx / 2; k += blockDim.x) {
      const float2* src = specT + (int64_t)k * pl.ld + ra;
      const float2 A = src[0];
      const float2 B = rb < nrows ? src[1] : make_float2(0.f, 0.f);
      bx[k] = make_float2(A.x - B.y, A.y + B.x);
      if (k > 0 && k < pl.Sx / 2) bx[pl.Sx - k] = make_float2(A.x + B.y, B.x - A.y);
    }
    __syncthreads();
    float2* z = fft_smem<true>(bx, by, tw, pl.Sx, pl.logSx, pl.rx);
    for (int j = threadIdx.x; j < ncols; j += blockDim.x) {
      const float2 v = z[wrap_index(j + off, pl.Sx)];
      int64_t oa = (int64_t)ra * ncols + j, ob = (int64_t)rb * ncols + j;
      if (OUT_MODE == 2) {
        const float bnorm = ds->lik.bkg_log_norm ? expf(__ldg(ds->lik.bkg_log_norm)) : 1.f;
        if (f == 1) {
          const float da = lik::poisson_px(v.x, __ldg(ds->lik.background + oa) * bnorm, __ldg(ds->lik.counts + oa), eps,
                                           grad_scale, lacc, bacc);
          if (ds->lik.dpool) ds->lik.dpool[oa] = da;
          if (rb < nrows) {
            const float db = lik::poisson_px(v.y, __ldg(ds->lik.background + ob) * bnorm, __ldg(ds->lik.counts + ob), eps,
                                             grad_scale, lacc, bacc);
            if (ds->lik.dpool) ds->lik.dpool[ob] = db;
          }
        } else if ((j & 1) == 0 && (j >> 1) < W && (ra >> 1) < H) {  // f = 2: rows (ra, rb) x columns (j, j + 1)
          const float2 v1 = z[wrap_index(j + 1 + off, pl.Sx)];
          const float pool = ((v.x + v1.x) + v.y) + v1.y;
          const int64_t o = (int64_t)(ra >> 1) * W + (j >> 1);
          const float d = lik::poisson_px(pool, __ldg(ds->lik.background + o) * bnorm, __ldg(ds->lik.counts + o), eps,
                                          grad_scale, lacc, bacc);
          if (ds->lik.dpool) ds->lik.dpool[o] = d;
        }
      } else if (OUT_MODE == 0) {
        out[oa] = v.x;
        if (rb < nrows) out[ob] = v.y;
      } else {
        float xa = v.x * (scale ? scale[oa] : 1.f);
        out[oa] = accumulate ? out[oa] + xa : xa;
        if (rb < nrows) {
          float xb = v.y * (scale ? scale[ob] : 1.f);
          out[ob] = accumulate ? out[ob] + xb : xb;
        }
      }
    }
    // bx/by are rewritten at the top of the next iteration after a barrier
  }
  if (OUT_MODE == 2) {  // loss (+ d loss / d log background norm) of this CTA's rows
    __shared__ float s_red[2][32];
    lacc = warp_sum(lacc);
    bacc = warp_sum(bacc);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_red[0][w] = lacc, s_red[1][w] = bacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double l = 0.0, b = 0.0;
      for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) l += (double)s_red[0][i], b += (double)s_red[1][i];
      if (blockIdx.x == 0) l += ds->lik.loss_const;
      if (ds->lik.loss_sum) atomicAdd(ds->lik.loss_sum, l);
      if (ds->lik.dlogb) atomicAdd(ds->lik.dlogb, b);
    }
  }
}

// sizes 3 * 2^L and 5 * 2^L are allowed next to 2^L (2.5x less padded area for 545 or 1224 points); JD_FFT_MIXED=0
// restricts the plan to powers of two
static bool fft_mixed_sizes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JD_FFT_MIXED");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}

static int make_plan(const char* name, int fH, int fW, int kh, int kw, Plan* pl) {
  JD_CHECK_ARG(fH > 0 && fW > 0 && kh > 0 && kw > 0, "%s: bad shape", name);
  pl->fH = fH;
  pl->fW = fW;
  pl->Sy = fft_size(fH + kh - 1, fft_mixed_sizes(), &pl->logSy, &pl->ry);
  pl->Sx = fft_size(fW + kw - 1, fft_mixed_sizes(), &pl->logSx, &pl->rx);
  JD_CHECK_ARG(pl->Sy <= 8192 && pl->Sx <= 8192, "%s: padded FFT size %dx%d exceeds 8192", name, pl->Sy, pl->Sx);
  pl->ld = (fH + 7) & ~7;
  return JD_OK;
}

static size_t smem_for(int S) { return (size_t)(2 * S + S / 2) * sizeof(float2); }
static int threads_for(int S) {
  int t = S / 4;
  return t < 64 ? 64 : (t > 512 ? 512 : t);
}

template <typename Kern>
static void opt_in_smem(Kern kern, size_t bytes) {
  if (bytes > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

}  // namespace fft
}  // namespace jd

using namespace jd;
using namespace jd::fft;

extern "C" {

int jd_fftconv_sizes(int fH, int fW, int kh, int kw, int64_t* psf_hat_elems, int64_t* workspace_elems) {
  Plan pl;
  int rc = make_plan("jd_fftconv_sizes", fH, fW, kh, kw, &pl);
  if (rc) return rc;
  if (psf_hat_elems) *psf_hat_elems = (int64_t)(pl.Sx / 2 + 1) * pl.Sy * 2;      // floats
  if (workspace_elems) *workspace_elems = (int64_t)(pl.Sx / 2 + 1) * pl.ld * 2;  // floats
  return JD_OK;
}

int jd_fftconv_prepare_psf(const float* psf, int kh, int kw, int fH, int fW, float* psf_hat, float* workspace,
                           jd_stream_t stream) {
  JD_CHECK_ARG(psf && psf_hat && workspace, "jd_fftconv_prepare_psf: null pointer");
  Plan pl;
  int rc = make_plan("jd_fftconv_prepare_psf", fH, fW, kh, kw, &pl);
  if (rc) return rc;
  JD_CHECK_ARG(kh <= pl.ld, "jd_fftconv_prepare_psf: PSF taller than the image is not supported (kh=%d, fH=%d)", kh, fH);
  cudaStream_t st = to_stream(stream);
  const int pairs = 1;
  size_t smx = smem_for(pl.Sx), smy = smem_for(pl.Sy);
  opt_in_smem(rows_fwd_kernel<IN_PSF>, smx);
  opt_in_smem(cols_kernel<2>, smy);
  rows_fwd_kernel<IN_PSF><<<(kh + 2 * pairs - 1) / (2 * pairs), threads_for(pl.Sx), smx, st>>>(
      psf, nullptr, pl, kh, kw, 1, kh, kw, reinterpret_cast<float2*>(workspace), pairs, nullptr);
  JD_CHECK_LAUNCH("jd_fftconv_prepare_psf(rows)");
  cols_kernel<2><<<pl.Sx / 2 + 1, threads_for(pl.Sy), smy, st>>>(reinterpret_cast<float2*>(workspace), nullptr, pl, kh, 0,
                                                                 0, 1.f, reinterpret_cast<float2*>(psf_hat), nullptr);
  JD_CHECK_LAUNCH("jd_fftconv_prepare_psf(cols)");
  return JD_OK;
}

// mode 0: forward model, 1: adjoint.  table == NULL: one dataset through the pointer arguments; else `n` datasets per
// launch (grid.y), forward with the fused Poisson epilogue (lik_f / H / W / eps / grad_scale).
static int run_fftconv(const char* name, int mode, const float* in, const float* exposure, const float* psf_hat,
                       float* workspace, float* out, int accumulate, int fH, int fW, int kh, int kw, int f, int H, int W,
                       cudaStream_t st, const jd_fftlik_dataset* table = nullptr, int n = 1, float eps = 0.f,
                       float grad_scale = 0.f) {
  Plan pl;
  int rc = make_plan(name, fH, fW, kh, kw, &pl);
  if (rc) return rc;
  const int sy = (kh - 1) / 2, sx = (kw - 1) / 2;
  const float norm = 1.0f / ((float)pl.Sy * (float)pl.Sx);
  size_t smx = smem_for(pl.Sx), smy = smem_for(pl.Sy);
  JD_CHECK_ARG(smx <= 200 * 1024 && smy <= 200 * 1024, "%s: FFT size too large for shared memory", name);
  const int pairs = 1;  // one row pair per CTA: >= 2 CTAs per SM already at 512 rows, phases of different CTAs overlap
  const dim3 grid_rows((fH + 2 * pairs - 1) / (2 * pairs), n), grid_cols(pl.Sx / 2 + 1, n);
  float2* spec = reinterpret_cast<float2*>(workspace);
  const float2* ph = reinterpret_cast<const float2*>(psf_hat);
  if (mode == 0) {
    opt_in_smem(rows_fwd_kernel<IN_FLUX>, smx);
    opt_in_smem(cols_kernel<0>, smy);
    rows_fwd_kernel<IN_FLUX><<<grid_rows, threads_for(pl.Sx), smx, st>>>(in, exposure, pl, fH, fW, 1, fH, fW, spec, pairs,
                                                                          table);
    JD_CHECK_LAUNCH(name);
    cols_kernel<0><<<grid_cols, threads_for(pl.Sy), smy, st>>>(spec, ph, pl, fH, fH, sy, norm, nullptr, table);
    JD_CHECK_LAUNCH(name);
    if (table) {
      opt_in_smem(rows_inv_kernel<2>, smx);
      rows_inv_kernel<2><<<grid_rows, threads_for(pl.Sx), smx, st>>>(spec, pl, fH, fW, sx, nullptr, nullptr, 0, pairs,
                                                                      table, f, H, W, eps, grad_scale);
    } else {
      opt_in_smem(rows_inv_kernel<0>, smx);
      rows_inv_kernel<0><<<grid_rows, threads_for(pl.Sx), smx, st>>>(spec, pl, fH, fW, sx, nullptr, out, 0, pairs, nullptr,
                                                                      1, fH, fW, 0.f, 0.f);
    }
    JD_CHECK_LAUNCH(name);
  } else {
    opt_in_smem(rows_fwd_kernel<IN_DPOOL>, smx);
    opt_in_smem(cols_kernel<1>, smy);
    opt_in_smem(rows_inv_kernel<1>, smx);
    rows_fwd_kernel<IN_DPOOL><<<grid_rows, threads_for(pl.Sx), smx, st>>>(in, nullptr, pl, fH, fW, f, H, W, spec, pairs,
                                                                           table);
    JD_CHECK_LAUNCH(name);
    cols_kernel<1><<<grid_cols, threads_for(pl.Sy), smy, st>>>(spec, ph, pl, fH, fH, pl.Sy - sy, norm, nullptr, table);
    JD_CHECK_LAUNCH(name);
    rows_inv_kernel<1><<<grid_rows, threads_for(pl.Sx), smx, st>>>(spec, pl, fH, fW, pl.Sx - sx, exposure, out, accumulate,
                                                                    pairs, table, f, H, W, 0.f, 0.f);
    JD_CHECK_LAUNCH(name);
  }
  return JD_OK;
}

int jd_conv_forward_fft(const float* flux, const float* exposure, const float* psf_hat, float* workspace, float* conv,
                        int fH, int fW, int kh, int kw, jd_stream_t stream) {
  JD_CHECK_ARG(flux && psf_hat && workspace && conv, "jd_conv_forward_fft: null pointer");
  return run_fftconv("jd_conv_forward_fft", 0, flux, exposure, psf_hat, workspace, conv, 0, fH, fW, kh, kw, 1, fH, fW,
                     to_stream(stream));
}

int jd_conv_backward_fft(const float* dpool, const float* exposure, const float* psf_hat, float* workspace,
                         float* dflux, int accumulate, int fH, int fW, int kh, int kw, int f, int H, int W,
                         jd_stream_t stream) {
  JD_CHECK_ARG(dpool && psf_hat && workspace && dflux, "jd_conv_backward_fft: null pointer");
  JD_CHECK_ARG(f >= 1 && H * f <= fH && W * f <= fW, "jd_conv_backward_fft: bad shape");
  return run_fftconv("jd_conv_backward_fft", 1, dpool, exposure, psf_hat, workspace, dflux, accumulate, fH, fW, kh, kw, f,
                     H, W, to_stream(stream));
}

int jd_likelihood_forward_fft(const jd_fftlik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                              int H, int W, float eps, float grad_scale, jd_stream_t stream) {
  JD_CHECK_ARG(table_dev && n_datasets > 0 && n_datasets <= 65535, "jd_likelihood_forward_fft: bad dataset table");
  JD_CHECK_ARG((f == 1 || f == 2) && H * f == fH && W * f == fW,
               "jd_likelihood_forward_fft: upsampling factor %d (supported: 1, 2) or counts grid %dx%d != flux grid / f", f,
               H, W);
  return run_fftconv("jd_likelihood_forward_fft", 0, nullptr, nullptr, nullptr, nullptr, nullptr, 0, fH, fW, kh, kw, f, H,
                     W, to_stream(stream), table_dev, n_datasets, eps, grad_scale);
}

int jd_likelihood_backward_fft(const jd_fftlik_dataset* table_dev, int n_datasets, int fH, int fW, int kh, int kw, int f,
                               int H, int W, jd_stream_t stream) {
  JD_CHECK_ARG(table_dev && n_datasets > 0 && n_datasets <= 65535, "jd_likelihood_backward_fft: bad dataset table");
  JD_CHECK_ARG(f >= 1 && H * f <= fH && W * f <= fW, "jd_likelihood_backward_fft: bad shape");
  return run_fftconv("jd_likelihood_backward_fft", 1, nullptr, nullptr, nullptr, nullptr, nullptr, 0, fH, fW, kh, kw, f, H,
                     W, to_stream(stream), table_dev, n_datasets);
}

}  // extern "C"
