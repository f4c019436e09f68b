// HBM-bound pieces of the MAP step: flux parameterisation, pooled Poisson cash statistic with its
// gradient, patch-gradient fold (col2im, gather form) and the fused gradient-sum + Adam update.
#include <math_constants.h>
#include <stdarg.h>

#include "jd_common.cuh"
#include "jd_shift.cuh"

namespace jd {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cache[64] = {};  // per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int& n = cache[dev & 63];
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------
// a1: flux = exp(theta) * mask                                   (models/core.py:583-594)
// ------------------------------------------------------------------------------------------
__global__ void flux_kernel(const float* __restrict__ theta, const uint8_t* __restrict__ mask,
                            float* __restrict__ flux, int64_t n, int use_log) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float t = theta[i];
    float f = use_log ? expf(t) : t;
    if (mask) f *= (float)mask[i];
    flux[i] = f;
  }
}

// ------------------------------------------------------------------------------------------
// a3/a4/a6: sum-pool, clip, background, Poisson NLL (full Stirling term) and gradient.
// One thread per counts pixel; warp-shuffle + shared block reduction, one double atomic per block.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void poisson_pixel(float pool, float bkg, float c, float eps, float grad_scale,
                                              float& np_, float& dpool, float& acc, float& accb) {
  np_ = fmaxf(pool, 0.f) + bkg;
  const float ne = np_ + eps;
  float loss = np_ - c * logf(ne);
  if (c > 1.f) {  // Stirling term c log c - c + 1/2 log(2 pi c), with log(2 pi c) = log c + log 2 pi
    const float lc = logf(c);
    loss += fmaf(c, lc, -c) + 0.5f * (lc + 1.8378770664093453f);
  }
  acc += loss;
  const float d = (1.f - c / ne) * grad_scale;
  accb = fmaf(d, bkg, accb);
  dpool = pool >= 0.f ? d : 0.f;
}

// VEC: f == 1 and W % 4 == 0 -> four pixels per thread per pass with 128-bit loads/stores
template <bool VEC>
__global__ void __launch_bounds__(256)
poisson_kernel(const float* __restrict__ conv, const float* __restrict__ background,
               const float* __restrict__ bkg_log_norm, const float* __restrict__ counts,
               float* __restrict__ npred_out, float* __restrict__ dpool_out, double* __restrict__ loss_sum,
               double* __restrict__ dlogb, int H, int W, int f, int fW, float eps, float grad_scale) {
  __shared__ double red[32];
  const float bnorm = bkg_log_norm ? expf(bkg_log_norm[0]) : 1.0f;
  const int64_t n = (int64_t)H * W;
  // per-thread partial sums in FP32 (a thread sees at most a few dozen pixels), FP64 from the block reduction on
  float acc = 0.f, accb = 0.f;
  if (VEC) {
    // two 4-pixel chunks per thread and pass: six independent 128-bit loads in flight per thread (the kernel is
    // bound by memory latency x bytes in flight, not by arithmetic)
    const int64_t n4 = n >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += 2 * stride) {
      const int64_t j = i + stride;
      const bool two = j < n4;
      int64_t ci = i * 4;  // f == 1: conv row stride fW, equal to W unless the caller passes a padded buffer
      if (fW != W) {
        const int64_t row = ci / W;
        ci = row * fW + (ci - row * W);
      }
      const float4 pv = *reinterpret_cast<const float4*>(conv + ci);
      const float4 bv = *reinterpret_cast<const float4*>(background + i * 4);
      const float4 cv = *reinterpret_cast<const float4*>(counts + i * 4);
      float4 pv2 = pv, bv2 = bv, cv2 = cv;
      if (two) {
        int64_t cj = j * 4;
        if (fW != W) {
          const int64_t row2 = cj / W;
          cj = row2 * fW + (cj - row2 * W);
        }
        pv2 = *reinterpret_cast<const float4*>(conv + cj);
        bv2 = *reinterpret_cast<const float4*>(background + j * 4);
        cv2 = *reinterpret_cast<const float4*>(counts + j * 4);
      }
      float4 nv, dv;
      poisson_pixel(pv.x, bv.x * bnorm, cv.x, eps, grad_scale, nv.x, dv.x, acc, accb);
      poisson_pixel(pv.y, bv.y * bnorm, cv.y, eps, grad_scale, nv.y, dv.y, acc, accb);
      poisson_pixel(pv.z, bv.z * bnorm, cv.z, eps, grad_scale, nv.z, dv.z, acc, accb);
      poisson_pixel(pv.w, bv.w * bnorm, cv.w, eps, grad_scale, nv.w, dv.w, acc, accb);
      if (npred_out) *reinterpret_cast<float4*>(npred_out + i * 4) = nv;
      if (dpool_out) *reinterpret_cast<float4*>(dpool_out + i * 4) = dv;
      if (two) {
        poisson_pixel(pv2.x, bv2.x * bnorm, cv2.x, eps, grad_scale, nv.x, dv.x, acc, accb);
        poisson_pixel(pv2.y, bv2.y * bnorm, cv2.y, eps, grad_scale, nv.y, dv.y, acc, accb);
        poisson_pixel(pv2.z, bv2.z * bnorm, cv2.z, eps, grad_scale, nv.z, dv.z, acc, accb);
        poisson_pixel(pv2.w, bv2.w * bnorm, cv2.w, eps, grad_scale, nv.w, dv.w, acc, accb);
        if (npred_out) *reinterpret_cast<float4*>(npred_out + j * 4) = nv;
        if (dpool_out) *reinterpret_cast<float4*>(dpool_out + j * 4) = dv;
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
      float pool = 0.f;
      const float* src = conv + (int64_t)y * f * fW + (int64_t)x * f;
      for (int u = 0; u < f; ++u)
        for (int v = 0; v < f; ++v) pool += src[(int64_t)u * fW + v];
      float np_, d;
      poisson_pixel(pool, background[i] * bnorm, counts[i], eps, grad_scale, np_, d, acc, accb);
      if (npred_out) npred_out[i] = np_;
      if (dpool_out) dpool_out[i] = d;
    }
  }
  double s = block_sum((double)acc, red);
  if (threadIdx.x == 0 && loss_sum) atomicAdd(loss_sum, s);
  if (dlogb) {
    double sb = block_sum((double)accb, red);
    if (threadIdx.x == 0) atomicAdd(dlogb, sb);
  }
}

// ------------------------------------------------------------------------------------------
// col2im, gather form: every flux pixel sums the patch-gradient entries that cover it.
// Rolled coordinate ry = (y + sy) mod fH; covering patch rows iy with 0 <= ry - s*iy < 8.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fold_gather(const float* __restrict__ G, int y, int x, int fH, int fW, int sy,
                                             int sx, int stride, int ny, int nx, int row_begin, int row_end) {
  int ry = wrap(y + sy, fH), rx = wrap(x + sx, fW);
  int iy_hi = min(ry / stride, min(ny, row_end) - 1);
  int ix_hi = min(rx / stride, nx - 1);
  float acc = 0.f;
  for (int iy = iy_hi; iy >= row_begin && ry - iy * stride < PATCH; --iy) {
    int u = ry - iy * stride;
    for (int ix = ix_hi; ix >= 0 && rx - ix * stride < PATCH; --ix) {
      int v = rx - ix * stride;
      int64_t p = (int64_t)(iy - row_begin) * nx + ix;
      acc += G[p * PD + u * PATCH + v];
    }
  }
  return acc;
}

// The same sum for four consecutive pixels of one image row at stride 4 (8 x 8 patches every 4 pixels: a pixel lies in
// at most 2 x 2 patches).  Same summation order per pixel as fold_gather (patch rows high -> low, patch columns high ->
// low), so the results are bit-identical; the row part (rolled row, covering patch rows) is computed once and the
// divisions by the stride are shifts.
__device__ __forceinline__ void fold_gather4_s4(const float* __restrict__ G, int y, int x, int fH, int fW, int sy, int sx,
                                                int ny, int nx, int row_begin, int row_end, float (&out)[4]) {
  const int ry = wrap(y + sy, fH);
  const int iy_hi = min(ry >> 2, min(ny, row_end) - 1);
  // up to two covering patch rows: (iy_hi, u0) and (iy_hi - 1, u0 + 4)
  const int u0 = ry - 4 * iy_hi;
  const bool r0 = iy_hi >= row_begin && u0 < PATCH, r1 = r0 && iy_hi - 1 >= row_begin && u0 + 4 < PATCH;
  const float* g0 = G + ((int64_t)(iy_hi - row_begin) * nx) * PD + u0 * PATCH;  // row iy_hi, patch column 0
  const float* g1 = g0 - (int64_t)nx * PD + 4 * PATCH;                          // row iy_hi - 1
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int rx = wrap(x + c + sx, fW);
    const int ix_hi = min(rx >> 2, nx - 1);
    const int v0 = rx - 4 * ix_hi;
    const bool c0 = v0 < PATCH, c1 = c0 && ix_hi >= 1 && v0 + 4 < PATCH;
    const int o0 = ix_hi * PD + v0, o1 = o0 - PD + 4;
    float acc = 0.f;
    if (r0) {
      if (c0) acc += g0[o0];
      if (c1) acc += g0[o1];
    }
    if (r1) {
      if (c0) acc += g1[o0];
      if (c1) acc += g1[o1];
    }
    out[c] = acc;
  }
}

__global__ void fold_kernel(const float* __restrict__ G, int fH, int fW, const int32_t* __restrict__ shift_yx,
                            int stride, int row_begin, int row_end, float* __restrict__ dflux, int accumulate) {
  const int sy = shift_yx ? shift_yx[0] : 0, sx = shift_yx ? shift_yx[1] : 0;
  const int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  const int64_t n = (int64_t)fH * fW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(i / fW), x = (int)(i - (int64_t)y * fW);
    float g = fold_gather(G, y, x, fH, fW, sy, sx, stride, ny, nx, row_begin, row_end);
    dflux[i] = accumulate ? dflux[i] + g : g;
  }
}

// ------------------------------------------------------------------------------------------
// a12: gradient sum + chain rule + Adam, same operation order as torch.optim.Adam (single tensor):
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(v)/sqrt(bc2) + eps;
//   theta.addcdiv_(m, denom, -lr/bc1)
// ------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ flux, const uint8_t* __restrict__ mask,
                            const float* __restrict__ da, const float* __restrict__ db, float scale_b, int use_log,
                            int64_t n, float lr_over_bc1, float sqrt_bc2, float b1, float b2, float eps) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = da[i];
    if (db) g += scale_b * db[i];
    g *= use_log ? flux[i] : (mask ? (float)mask[i] : 1.f);
    float mi = m[i], vi = v[i];
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * g * g;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    theta[i] = theta[i] - lr_over_bc1 * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// sum-pool f x f (F.avg_pool2d(divisor_override=1), models/npred.py:181-184); pre-clip values
__global__ void pool_kernel(const float* __restrict__ conv, float* __restrict__ pool, int H, int W, int f, int fW) {
  const int64_t n = (int64_t)H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
    const float* src = conv + (int64_t)y * f * fW + (int64_t)x * f;
    float acc = 0.f;
    for (int u = 0; u < f; ++u)
      for (int v = 0; v < f; ++v) acc += src[(int64_t)u * fW + v];
    pool[i] = acc;
  }
}

// Device-side bookkeeping of one MAP step so the whole step replays from a CUDA graph:
// counters[0] = number of cycle-spin draws consumed, counters[1] = Adam step count t.
__global__ void step_begin_kernel(int32_t* __restrict__ counters, const int32_t* __restrict__ shift_table,
                                  int n_shifts, int32_t* __restrict__ shift_out, int advance_adam, float lr, float b1,
                                  float b2, float* __restrict__ adam_scalars, double* __restrict__ zero_acc,
                                  int n_acc) {
  if (threadIdx.x == 0) {
    if (shift_table && shift_out) {
      int d = counters[0];
      int idx = d < n_shifts ? d : n_shifts - 1;
      shift_out[0] = shift_table[2 * idx];
      shift_out[1] = shift_table[2 * idx + 1];
      counters[0] = d + 1;
    }
    if (advance_adam) {
      int t = counters[1] + 1;
      counters[1] = t;
      double bc1 = 1.0 - pow((double)b1, (double)t);
      double bc2 = 1.0 - pow((double)b2, (double)t);
      adam_scalars[0] = (float)((double)lr / bc1);
      adam_scalars[1] = (float)sqrt(bc2);
    }
  }
  for (int i = threadIdx.x; i < n_acc; i += blockDim.x) zero_acc[i] = 0.0;
}

__global__ void adam_dev_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                                const float* __restrict__ flux, const uint8_t* __restrict__ mask,
                                const float* __restrict__ da, const float* __restrict__ db, float scale_b, int use_log,
                                int64_t n, const float* __restrict__ scalars, float b1, float b2, float eps) {
  const float lr_over_bc1 = scalars[0], sqrt_bc2 = scalars[1];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = da[i];
    if (db) g += scale_b * db[i];
    g *= use_log ? flux[i] : (mask ? (float)mask[i] : 1.f);
    float mi = m[i], vi = v[i];
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * g * g;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    theta[i] = theta[i] - lr_over_bc1 * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// step_begin + flux in one launch: block 0 does the bookkeeping, every block its share of exp(theta)
__global__ void step_begin_flux_kernel(int32_t* __restrict__ counters, const int32_t* __restrict__ shift_table,
                                       int n_shifts, int32_t* __restrict__ shift_out, int advance_adam, float lr,
                                       float b1, float b2, float* __restrict__ adam_scalars,
                                       double* __restrict__ zero_acc, int n_acc, const float* __restrict__ theta,
                                       const uint8_t* __restrict__ mask, float* __restrict__ flux, int64_t n,
                                       int use_log) {
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      if (shift_table && shift_out) {
        int d = counters[0];
        int idx = d < n_shifts ? d : n_shifts - 1;
        shift_out[0] = shift_table[2 * idx];
        shift_out[1] = shift_table[2 * idx + 1];
        counters[0] = d + 1;
      }
      if (advance_adam) {
        int t = counters[1] + 1;
        counters[1] = t;
        double bc1 = 1.0 - pow((double)b1, (double)t);
        double bc2 = 1.0 - pow((double)b2, (double)t);
        adam_scalars[0] = (float)((double)lr / bc1);
        adam_scalars[1] = (float)sqrt(bc2);
      }
    }
    for (int i = threadIdx.x; i < n_acc; i += blockDim.x) zero_acc[i] = 0.0;
  }
  if (!flux) return;
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(flux)) & 15) == 0) {
    const int64_t n4 = n >> 2;  // four pixels per thread, 128-bit loads / stores
    for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x) {
      const float4 t = *reinterpret_cast<const float4*>(theta + 4 * i4);
      float4 f = use_log ? make_float4(expf(t.x), expf(t.y), expf(t.z), expf(t.w)) : t;
      if (mask) {
        const uchar4 mk = *reinterpret_cast<const uchar4*>(mask + 4 * i4);
        f.x *= (float)mk.x, f.y *= (float)mk.y, f.z *= (float)mk.z, f.w *= (float)mk.w;
      }
      *reinterpret_cast<float4*>(flux + 4 * i4) = f;
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t = theta[i];
    float f = use_log ? expf(t) : t;
    if (mask) f *= (float)mask[i];
    flux[i] = f;
  }
}

// fold (gather col2im of the patch gradients) + gradient sum + chain rule + Adam in one pass
__global__ void adam_fold_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                                 const float* __restrict__ flux, const uint8_t* __restrict__ mask,
                                 const float* __restrict__ da, const float* __restrict__ G, float scale_b, int use_log,
                                 int fH, int fW, const int32_t* __restrict__ shift_yx, int stride, int row_begin,
                                 int row_end, const float* __restrict__ scalars, float b1, float b2, float eps) {
  const float lr_over_bc1 = scalars[0], sqrt_bc2 = scalars[1];
  const int sy = shift_yx ? shift_yx[0] : 0, sx = shift_yx ? shift_yx[1] : 0;
  const int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  const int64_t n = (int64_t)fH * fW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int y = (int)(i / fW), x = (int)(i - (int64_t)y * fW);
    float g = da[i] + scale_b * fold_gather(G, y, x, fH, fW, sy, sx, stride, ny, nx, row_begin, row_end);
    g *= use_log ? flux[i] : (mask ? (float)mask[i] : 1.f);
    float mi = m[i], vi = v[i];
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * g * g;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    theta[i] = theta[i] - lr_over_bc1 * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// Joint step, one GPU: sum of the per-dataset likelihood gradients (`n_parts` images, `part_stride` floats apart,
// fixed order) + fold of the prior's patch gradients (G may be NULL) + chain rule + Adam, one pass.
// Four consecutive pixels per thread with 128-bit loads / stores (fW % 4 == 0, 16-byte aligned images); UPDATE = false
// only writes the gradient (this rank's partial gradient of a multi-GPU joint step).
template <bool UPDATE>
__global__ void __launch_bounds__(256)
joint_grad_kernel(float* __restrict__ theta, float* __restrict__ m, float* __restrict__ v,
                  const float* __restrict__ flux, const uint8_t* __restrict__ mask, const float* __restrict__ parts,
                  int n_parts, int64_t part_stride, const float* __restrict__ G, float scale_b, int use_log, int fH,
                  int fW, const int32_t* __restrict__ shift_yx, int stride, int row_begin, int row_end,
                  const float* __restrict__ scalars, float b1, float b2, float eps, float* __restrict__ out, int vec) {
  const float lr_over_bc1 = UPDATE ? scalars[0] : 0.f, sqrt_bc2 = UPDATE ? scalars[1] : 1.f;
  const int sy = shift_yx ? shift_yx[0] : 0, sx = shift_yx ? shift_yx[1] : 0;
  const int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  const int64_t n = (int64_t)fH * fW;
  const bool fold = G != nullptr && row_end > row_begin;
  __shared__ __align__(16) float s_fold[4 * 256];
  // one image row per CTA pass: fW = 4 x blockDim (every thread then runs the same number of passes)
  const bool row_fold = fold && vec && stride == 4 && fW == 4 * (int)blockDim.x && blockDim.x == 256;
  if (vec) {
    const int64_t n4 = n >> 2;
    for (int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * blockDim.x) {
      const int64_t i = i4 * 4;
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      // (groups of up to eight independent loads in flight, summed in the fixed order q = 0, 1, ...)
      for (int q0 = 0; q0 < n_parts; q0 += 8) {
        float4 p4[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          p4[u] = q0 + u < n_parts ? __ldcs(reinterpret_cast<const float4*>(parts + (q0 + u) * part_stride + i))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (q0 + u < n_parts) g[0] += p4[u].x, g[1] += p4[u].y, g[2] += p4[u].z, g[3] += p4[u].w;
      }
      if (fold) {
        const int y = (int)(i / fW), x = (int)(i - (int64_t)y * fW);
        if (row_fold) {
          // The CTA's 256 threads cover exactly one image row: fold the ROLLED row into shared memory with threads on
          // the patch grid (rolled columns 4t .. 4t+3 = one 16-byte piece of at most 2 x 2 patch rows: four 128-bit
          // loads instead of sixteen scattered 32-bit ones, same summation order per pixel), then read it back at the
          // pixel's rolled position.
          const int t = threadIdx.x;
          const int ry = wrap(y + sy, fH);
          const int iy_hi = min(ry >> 2, min(ny, row_end) - 1);
          const int u0 = ry - 4 * iy_hi;
          const bool r0 = iy_hi >= row_begin && u0 < PATCH, r1 = r0 && iy_hi - 1 >= row_begin && u0 + 4 < PATCH;
          const int ix_hi = min(t, nx - 1), v0 = 4 * (t - ix_hi);
          const bool c0 = v0 < PATCH, c1 = c0 && ix_hi >= 1 && v0 + 4 < PATCH;
          const float* g0 = G + ((int64_t)(iy_hi - row_begin) * nx + ix_hi) * PD + u0 * PATCH + v0;
          const float* g1 = g0 - (int64_t)nx * PD + 4 * PATCH;
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a00 = (r0 && c0) ? __ldg(reinterpret_cast<const float4*>(g0)) : z4;
          const float4 a01 = (r0 && c1) ? __ldg(reinterpret_cast<const float4*>(g0 - PD + 4)) : z4;
          const float4 a10 = (r1 && c0) ? __ldg(reinterpret_cast<const float4*>(g1)) : z4;
          const float4 a11 = (r1 && c1) ? __ldg(reinterpret_cast<const float4*>(g1 - PD + 4)) : z4;
          float4 acc = z4;  // skipped terms add +0.0f to a sum that starts at +0.0f: the same value as not adding them
          acc.x += a00.x, acc.y += a00.y, acc.z += a00.z, acc.w += a00.w;
          acc.x += a01.x, acc.y += a01.y, acc.z += a01.z, acc.w += a01.w;
          acc.x += a10.x, acc.y += a10.y, acc.z += a10.z, acc.w += a10.w;
          acc.x += a11.x, acc.y += a11.y, acc.z += a11.z, acc.w += a11.w;
          __syncthreads();  // the previous row has been read
          *reinterpret_cast<float4*>(s_fold + 4 * t) = acc;
          __syncthreads();
#pragma unroll
          for (int c = 0; c < 4; ++c) g[c] += scale_b * s_fold[wrap(x + c + sx, fW)];
        } else if (stride == 4) {
          float fg[4];
          fold_gather4_s4(G, y, x, fH, fW, sy, sx, ny, nx, row_begin, row_end, fg);
#pragma unroll
          for (int c = 0; c < 4; ++c) g[c] += scale_b * fg[c];
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            g[c] += scale_b * fold_gather(G, y, x + c, fH, fW, sy, sx, stride, ny, nx, row_begin, row_end);
        }
      }
      if (!UPDATE) {
        *reinterpret_cast<float4*>(out + i) = make_float4(g[0], g[1], g[2], g[3]);
        continue;
      }
      const float4 f4 = use_log ? *reinterpret_cast<const float4*>(flux + i) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float ff[4] = {f4.x, f4.y, f4.z, f4.w};
      const float4 t4 = *reinterpret_cast<const float4*>(theta + i);
      const float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
      float th[4] = {t4.x, t4.y, t4.z, t4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float gc = g[c] * (use_log ? ff[c] : (mask ? (float)mask[i + c] : 1.f));
        mm[c] = mm[c] + (gc - mm[c]) * (1.f - b1);
        vv[c] = vv[c] * b2 + (1.f - b2) * gc * gc;
        const float denom = sqrtf(vv[c]) / sqrt_bc2 + eps;
        th[c] = th[c] - lr_over_bc1 * (mm[c] / denom);
      }
      *reinterpret_cast<float4*>(theta + i) = make_float4(th[0], th[1], th[2], th[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = 0.f;
    for (int q = 0; q < n_parts; ++q) g += parts[q * part_stride + i];
    if (fold) {
      int y = (int)(i / fW), x = (int)(i - (int64_t)y * fW);
      g += scale_b * fold_gather(G, y, x, fH, fW, sy, sx, stride, ny, nx, row_begin, row_end);
    }
    if (!UPDATE) {
      out[i] = g;
      continue;
    }
    g *= use_log ? flux[i] : (mask ? (float)mask[i] : 1.f);
    float mi = m[i], vi = v[i];
    mi = mi + (g - mi) * (1.f - b1);
    vi = vi * b2 + (1.f - b2) * g * g;
    float denom = sqrtf(vi) / sqrt_bc2 + eps;
    theta[i] = theta[i] - lr_over_bc1 * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

// Adam on a handful of scalar calibration parameters (log background norm) whose gradients were accumulated in
// double by the Poisson kernel.  torch.optim.Adam keeps one step counter per parameter and only advances it when
// the parameter has a gradient (core.py:197-204, 229): `counter` is that per-parameter count.
__global__ void adam_scalar_kernel(float* __restrict__ param, float* __restrict__ m, float* __restrict__ v,
                                   const double* __restrict__ grad, int32_t* __restrict__ counter, int n, float lr,
                                   float b1, float b2, float eps) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int t = counter[0] + 1;
    counter[0] = t;
    const double bc1 = 1.0 - pow((double)b1, (double)t), bc2 = 1.0 - pow((double)b2, (double)t);
    const float step_size = (float)((double)lr / bc1), sqrt_bc2 = (float)sqrt(bc2);
    for (int i = 0; i < n; ++i) {
      const float g = (float)grad[i];
      float mi = m[i], vi = v[i];
      mi = mi + (g - mi) * (1.f - b1);
      vi = vi * b2 + (1.f - b2) * g * g;
      const float denom = sqrtf(vi) / sqrt_bc2 + eps;
      param[i] = param[i] - step_size * (mi / denom);
      m[i] = mi;
      v[i] = vi;
    }
  }
}

// ------------------------------------------------------------------------------------------
// NPredCalibration sub-pixel shift (utils/torch.py:196-223): 4-tap stencil with weights derived from the
// device-resident (shift_x, shift_y) pair; per-pixel arithmetic in jd_shift.cuh (host-checked against the oracle).
// ------------------------------------------------------------------------------------------
__global__ void shift_fwd_kernel(const float* __restrict__ flux, const float* __restrict__ shift_xy, int scale, int H,
                                 int W, float* __restrict__ out) {
  const ShiftTaps t = shift_taps(shift_xy[0], shift_xy[1], scale, H, W);
  const int64_t n = (int64_t)H * W;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / W), j = (int)(idx - (int64_t)i * W);
    out[idx] = shift_sample(flux, H, W, i, j, t);
  }
}

// dflux (+)= shift^T d;  dshift_xy += (sum d * d shifted / d shift_x, sum d * d shifted / d shift_y)
__global__ void __launch_bounds__(256)
shift_bwd_kernel(const float* __restrict__ d, const float* __restrict__ flux, const float* __restrict__ shift_xy,
                 int scale, int H, int W, float* __restrict__ dflux, int accumulate, double* __restrict__ dshift_xy) {
  __shared__ double red[32];
  const ShiftTaps t = shift_taps(shift_xy[0], shift_xy[1], scale, H, W);
  const int64_t n = (int64_t)H * W;
  float ax = 0.f, ay = 0.f;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / W), j = (int)(idx - (int64_t)i * W);
    const float g = shift_adjoint(d, H, W, i, j, t);
    dflux[idx] = accumulate ? dflux[idx] + g : g;
    if (dshift_xy) {
      float d_dy, d_dx;
      shift_dshift(flux, H, W, i, j, t, &d_dy, &d_dx);
      const float di = d[idx];
      ax = fmaf(di, d_dx, ax);
      ay = fmaf(di, d_dy, ay);
    }
  }
  if (dshift_xy) {
    const double sx = block_sum((double)ax, red);
    if (threadIdx.x == 0) atomicAdd(dshift_xy, sx);
    const double sy = block_sum((double)ay, red);
    if (threadIdx.x == 0) atomicAdd(dshift_xy + 1, sy);
  }
}

static inline int grid_for(int64_t n, int block) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)num_sms() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace jd

using namespace jd;

static int joint_vec_ok(int fW, int64_t part_stride, const void* a, const void* b, const void* c, const void* d,
                        const void* e, const void* f) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                         reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f);
  return (fW & 3) == 0 && (part_stride & 3) == 0 && (bits & 15) == 0;
}

extern "C" {

int jd_abi_version(void) { return JD_ABI_VERSION; }

const char* jd_last_error(void) { return jd::g_err; }

int jd_device_supported(int device) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) {
    set_error("jd_device_supported: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return JD_ERR_NO_DEVICE;
  }
  return major == 10 ? 1 : 0;
}

int jd_flux_forward(const float* theta, const uint8_t* mask, float* flux, int64_t n, int use_log_flux,
                    jd_stream_t stream) {
  JD_CHECK_ARG(theta && flux && n > 0, "jd_flux_forward: null pointer or n <= 0");
  flux_kernel<<<grid_for(n, 256), 256, 0, to_stream(stream)>>>(theta, mask, flux, n, use_log_flux);
  JD_CHECK_LAUNCH("jd_flux_forward");
  return JD_OK;
}

int jd_poisson_forward_backward(const float* conv, const float* background, const float* bkg_log_norm,
                                const float* counts, float* npred, float* dpool, double* loss_sum,
                                double* dlogb, int H, int W, int f, int fW, float eps, float grad_scale,
                                jd_stream_t stream) {
  JD_CHECK_ARG(conv && background && counts, "jd_poisson_forward_backward: null input");
  JD_CHECK_ARG(H > 0 && W > 0 && f >= 1 && fW >= W * f, "jd_poisson_forward_backward: bad shape H=%d W=%d f=%d fW=%d",
               H, W, f, fW);
  int64_t n = (int64_t)H * W;
  // few, fat blocks: one double atomic per block on the same accumulator would otherwise serialise
  const bool vec = f == 1 && (W & 3) == 0 && (fW & 3) == 0;
  int64_t want = ((vec ? n / 4 : n) + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > 4 * num_sms() ? 4 * num_sms() : want));
  if (vec)
    poisson_kernel<true><<<grid, 256, 0, to_stream(stream)>>>(conv, background, bkg_log_norm, counts, npred, dpool,
                                                              loss_sum, dlogb, H, W, f, fW, eps, grad_scale);
  else
    poisson_kernel<false><<<grid, 256, 0, to_stream(stream)>>>(conv, background, bkg_log_norm, counts, npred, dpool,
                                                               loss_sum, dlogb, H, W, f, fW, eps, grad_scale);
  JD_CHECK_LAUNCH("jd_poisson_forward_backward");
  return JD_OK;
}

int jd_pool_sum(const float* conv, float* pool, int H, int W, int f, int fW, jd_stream_t stream) {
  JD_CHECK_ARG(conv && pool && H > 0 && W > 0 && f >= 1 && fW >= W * f, "jd_pool_sum: bad arguments");
  pool_kernel<<<grid_for((int64_t)H * W, 256), 256, 0, to_stream(stream)>>>(conv, pool, H, W, f, fW);
  JD_CHECK_LAUNCH("jd_pool_sum");
  return JD_OK;
}

int jd_step_begin(int32_t* counters, const int32_t* shift_table, int n_shifts, int32_t* shift_out, int advance_adam,
                  float lr, float beta1, float beta2, float* adam_scalars, double* zero_acc, int n_acc,
                  jd_stream_t stream) {
  JD_CHECK_ARG(counters, "jd_step_begin: null counters");
  JD_CHECK_ARG(!advance_adam || adam_scalars, "jd_step_begin: adam_scalars required");
  JD_CHECK_ARG(!shift_table || n_shifts > 0, "jd_step_begin: empty shift table");
  JD_CHECK_ARG(n_acc == 0 || zero_acc, "jd_step_begin: null accumulator block");
  step_begin_kernel<<<1, 32, 0, to_stream(stream)>>>(counters, shift_table, n_shifts, shift_out, advance_adam, lr,
                                                     beta1, beta2, adam_scalars, zero_acc, n_acc);
  JD_CHECK_LAUNCH("jd_step_begin");
  return JD_OK;
}

int jd_step_begin_flux(int32_t* counters, const int32_t* shift_table, int n_shifts, int32_t* shift_out,
                       int advance_adam, float lr, float beta1, float beta2, float* adam_scalars, double* zero_acc,
                       int n_acc, const float* theta, const uint8_t* mask, float* flux, int64_t n, int use_log_flux,
                       jd_stream_t stream) {
  JD_CHECK_ARG(counters && theta && flux && n > 0, "jd_step_begin_flux: null pointer");
  JD_CHECK_ARG(!advance_adam || adam_scalars, "jd_step_begin_flux: adam_scalars required");
  JD_CHECK_ARG(!shift_table || n_shifts > 0, "jd_step_begin_flux: empty shift table");
  JD_CHECK_ARG(n_acc == 0 || zero_acc, "jd_step_begin_flux: null accumulator block");
  step_begin_flux_kernel<<<grid_for((n & 3) ? n : n >> 2, 256), 256, 0, to_stream(stream)>>>(counters, shift_table, n_shifts, shift_out,
                                                                           advance_adam, lr, beta1, beta2, adam_scalars,
                                                                           zero_acc, n_acc, theta, mask, flux, n,
                                                                           use_log_flux);
  JD_CHECK_LAUNCH("jd_step_begin_flux");
  return JD_OK;
}

int jd_adam_fold_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                          const float* dflux_a, const float* G, float scale_b, int use_log_flux, int fH, int fW,
                          const int32_t* shift_yx, int stride, int row_begin, int row_end, const float* adam_scalars,
                          float beta1, float beta2, float eps, jd_stream_t stream) {
  JD_CHECK_ARG(theta && m && v && dflux_a && G && adam_scalars, "jd_adam_fold_step_dev: null pointer");
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_fold_step_dev: flux required for the log parameterisation");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_adam_fold_step_dev: bad geometry");
  int ny = (fH - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin <= row_end, "jd_adam_fold_step_dev: bad row block");
  adam_fold_kernel<<<grid_for((int64_t)fH * fW, 256), 256, 0, to_stream(stream)>>>(
      theta, m, v, flux, mask, dflux_a, G, scale_b, use_log_flux, fH, fW, shift_yx, stride, row_begin, row_end,
      adam_scalars, beta1, beta2, eps);
  JD_CHECK_LAUNCH("jd_adam_fold_step_dev");
  return JD_OK;
}

int jd_adam_joint_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                           const float* parts, int n_parts, int64_t part_stride, const float* G, float scale_b,
                           int use_log_flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* adam_scalars, float beta1, float beta2, float eps,
                           jd_stream_t stream) {
  JD_CHECK_ARG(theta && m && v && adam_scalars && fH > 0 && fW > 0, "jd_adam_joint_step_dev: null pointer");
  JD_CHECK_ARG(n_parts >= 0 && (n_parts == 0 || parts), "jd_adam_joint_step_dev: bad gradient parts");
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_joint_step_dev: flux required for the log parameterisation");
  if (G) {
    JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_adam_joint_step_dev: bad geometry");
    int ny = (fH - PATCH) / stride + 1;
    JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin <= row_end, "jd_adam_joint_step_dev: bad row block");
  } else {
    stride = 1;
  }
  const int vec = joint_vec_ok(fW, part_stride, theta, m, v, flux, parts, nullptr);
  joint_grad_kernel<true><<<grid_for(((int64_t)fH * fW) >> (vec ? 2 : 0), 256), 256, 0, to_stream(stream)>>>(
      theta, m, v, flux, mask, parts, n_parts, part_stride, G, scale_b, use_log_flux, fH, fW, shift_yx, stride, row_begin,
      row_end, adam_scalars, beta1, beta2, eps, nullptr, vec);
  JD_CHECK_LAUNCH("jd_adam_joint_step_dev");
  return JD_OK;
}

int jd_grad_reduce_local(const float* parts, int n_parts, int64_t part_stride, const float* G, float scale_b, int fH,
                         int fW, const int32_t* shift_yx, int stride, int row_begin, int row_end, float* out,
                         jd_stream_t stream) {
  JD_CHECK_ARG(out && fH > 0 && fW > 0 && n_parts >= 0 && (n_parts == 0 || parts), "jd_grad_reduce_local: bad arguments");
  if (G) {
    JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_grad_reduce_local: bad geometry");
    int ny = (fH - PATCH) / stride + 1;
    JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin <= row_end, "jd_grad_reduce_local: bad row block");
  } else {
    stride = 1;
  }
  const int vec = joint_vec_ok(fW, part_stride, nullptr, nullptr, nullptr, nullptr, parts, out);
  joint_grad_kernel<false><<<grid_for(((int64_t)fH * fW) >> (vec ? 2 : 0), 256), 256, 0, to_stream(stream)>>>(
      nullptr, nullptr, nullptr, nullptr, nullptr, parts, n_parts, part_stride, G, scale_b, 0, fH, fW, shift_yx, stride,
      row_begin, row_end, nullptr, 0.f, 0.f, 0.f, out, vec);
  JD_CHECK_LAUNCH("jd_grad_reduce_local");
  return JD_OK;
}

int jd_shift_forward(const float* flux, const float* shift_xy, int scale, int fH, int fW, float* shifted,
                     jd_stream_t stream) {
  JD_CHECK_ARG(flux && shift_xy && shifted && flux != shifted, "jd_shift_forward: bad pointers");
  JD_CHECK_ARG(scale >= 1 && fH > 0 && fW > 0, "jd_shift_forward: bad shape");
  shift_fwd_kernel<<<grid_for((int64_t)fH * fW, 256), 256, 0, to_stream(stream)>>>(flux, shift_xy, scale, fH, fW, shifted);
  JD_CHECK_LAUNCH("jd_shift_forward");
  return JD_OK;
}

int jd_shift_backward(const float* dshifted, const float* flux, const float* shift_xy, int scale, int fH, int fW,
                      float* dflux, int accumulate, double* dshift_xy, jd_stream_t stream) {
  JD_CHECK_ARG(dshifted && flux && shift_xy && dflux && dflux != dshifted, "jd_shift_backward: bad pointers");
  JD_CHECK_ARG(scale >= 1 && fH > 0 && fW > 0, "jd_shift_backward: bad shape");
  shift_bwd_kernel<<<grid_for((int64_t)fH * fW, 256), 256, 0, to_stream(stream)>>>(dshifted, flux, shift_xy, scale, fH, fW,
                                                                                  dflux, accumulate, dshift_xy);
  JD_CHECK_LAUNCH("jd_shift_backward");
  return JD_OK;
}

int jd_adam_scalar_step_dev(float* param, float* m, float* v, const double* grad, int32_t* counter, int n, float lr,
                            float beta1, float beta2, float eps, jd_stream_t stream) {
  JD_CHECK_ARG(param && m && v && grad && counter && n > 0 && n <= 64, "jd_adam_scalar_step_dev: bad arguments");
  adam_scalar_kernel<<<1, 32, 0, to_stream(stream)>>>(param, m, v, grad, counter, n, lr, beta1, beta2, eps);
  JD_CHECK_LAUNCH("jd_adam_scalar_step_dev");
  return JD_OK;
}

int jd_adam_step_dev(float* theta, float* m, float* v, const float* flux, const uint8_t* mask, const float* dflux_a,
                     const float* dflux_b, float scale_b, int use_log_flux, int64_t n, const float* adam_scalars,
                     float beta1, float beta2, float eps, jd_stream_t stream) {
  JD_CHECK_ARG(theta && m && v && dflux_a && adam_scalars && n > 0, "jd_adam_step_dev: bad arguments");
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_step_dev: flux required for the log parameterisation");
  adam_dev_kernel<<<grid_for(n, 256), 256, 0, to_stream(stream)>>>(theta, m, v, flux, mask, dflux_a, dflux_b, scale_b,
                                                                    use_log_flux, n, adam_scalars, beta1, beta2, eps);
  JD_CHECK_LAUNCH("jd_adam_step_dev");
  return JD_OK;
}

int jd_patch_fold(const float* G, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                  int row_end, float* dflux, int accumulate, jd_stream_t stream) {
  JD_CHECK_ARG(G && dflux, "jd_patch_fold: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_patch_fold: bad geometry");
  int ny = (fH - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin <= row_end, "jd_patch_fold: bad row block [%d,%d) of %d",
               row_begin, row_end, ny);
  int64_t n = (int64_t)fH * fW;
  fold_kernel<<<grid_for(n, 256), 256, 0, to_stream(stream)>>>(G, fH, fW, shift_yx, stride, row_begin, row_end,
                                                                dflux, accumulate);
  JD_CHECK_LAUNCH("jd_patch_fold");
  return JD_OK;
}

int jd_adam_step(float* theta, float* m, float* v, const float* flux, const uint8_t* mask,
                 const float* dflux_a, const float* dflux_b, float scale_b, int use_log_flux, int64_t n,
                 int step, float lr, float beta1, float beta2, float eps, jd_stream_t stream) {
  JD_CHECK_ARG(theta && m && v && dflux_a && n > 0 && step >= 1, "jd_adam_step: bad arguments");
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_step: flux required for the log parameterisation");
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n, 256), 256, 0, to_stream(stream)>>>(theta, m, v, flux, mask, dflux_a, dflux_b, scale_b,
                                                                use_log_flux, n, (float)((double)lr / bc1),
                                                                (float)sqrt(bc2), beta1, beta2, eps);
  JD_CHECK_LAUNCH("jd_adam_step");
  return JD_OK;
}

}  // extern "C"
