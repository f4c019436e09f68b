// GMM patch prior on FP32 CUDA cores: the check path for the tcgen05 kernel (jd_gmm_tc.cu), the
// logsumexp backward, the max-mode gather-GEMV backward and the generic-D log-prob.
//
// Forward tile kernel: a CTA owns 128 patches.  The mean-subtracted patches are gathered straight
// from the flux image at rolled coordinates (no roll / unfold / reshape copies) into shared memory
// (feature-major), then for every mixture component the 64x64 matrix Lw_k is streamed through a
// double-buffered cp.async stage and a 128x64x64 register-tiled product (8x4 per thread) is formed.
// The epilogue never writes Y: it subtracts mw_k, squares, reduces over the 64 whitened features
// with half-warp shuffles and folds the component into a running max/argmax or online logsumexp.
#include <algorithm>
#include <cuda_pipeline.h>
#include <math_constants.h>

#include "jd_common.cuh"

namespace jd {

constexpr int TM = 128;        // patches per CTA
constexpr int NT = 256;        // threads per CTA
constexpr int AS = TM + 4;     // padded row length of the feature-major A tile

struct PatchGeom {
  int fH, fW, sy, sx, stride, nx, row_begin, P;  // P = local patch count
};

// Source offset of element (u,v) of patch (iy,ix): the rolled image r[y,x] = flux[(y-sy) mod fH, (x-sx) mod fW]
// (torch.roll, utils/torch.py:118-119) sampled at r[s*iy+u, s*ix+v] (unfold, utils/torch.py:226-275).
// Every kernel that touches patches goes through this one function.
__device__ __forceinline__ int patch_src_row(const PatchGeom& g, int iy, int u) {
  return wrap(iy * g.stride + u - g.sy, g.fH);
}
__device__ __forceinline__ int patch_src_col(const PatchGeom& g, int ix, int v) {
  return wrap(ix * g.stride + v - g.sx, g.fW);
}

// Gather + mean-subtract 128 patches into As[d][row]; returns validity per row in s_valid.
__device__ __forceinline__ void gather_patches(const float* __restrict__ flux, const PatchGeom& g, int64_t p0,
                                               float* __restrict__ As, int* __restrict__ s_valid) {
  // two threads per patch: each loads 4 patch rows (32 values)
  const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int64_t p = p0 + row;
  float vals[32];
  float sum = 0.f;
  bool ok = true;
  if (p < g.P) {
    int iy = (int)(p / g.nx) + g.row_begin, ix = (int)(p % g.nx);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float* src = flux + (int64_t)patch_src_row(g, iy, half * 4 + u) * g.fW;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        float val = __ldg(src + patch_src_col(g, ix, v));
        vals[u * 8 + v] = val;
        sum += val;
        ok = ok && (val > -1e5f);  // false for NaN too
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) vals[i] = 0.f;
    ok = false;
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const int ok_other = __shfl_xor_sync(0xffffffffu, (int)ok, 1);  // unconditional: no short-circuit around a shuffle
  ok = ok && ok_other;
  const float mean = sum * (1.f / 64.f);
#pragma unroll
  for (int i = 0; i < 32; ++i) As[(half * 32 + i) * AS + row] = ok ? vals[i] - mean : 0.f;
  if (half == 0) s_valid[row] = ok ? 1 : 0;
}

__device__ __forceinline__ void stage_B(const float* __restrict__ B, float* __restrict__ dst) {
  // 64x64 floats = 1024 float4, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int idx = threadIdx.x + i * NT;
    __pipeline_memcpy_async(dst + idx * 4, B + idx * 4, 16);
  }
}

// acc[i][j] = sum_d As[d][ty*8+i] * Bs[d][tx*4+j]
__device__ __forceinline__ void tile_mma(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx,
                                         float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int d = 0; d < PD; ++d) {
    float4 a0 = *reinterpret_cast<const float4*>(As + d * AS + ty * 8);
    float4 a1 = *reinterpret_cast<const float4*>(As + d * AS + ty * 8 + 4);
    float4 b = *reinterpret_cast<const float4*>(Bs + d * PD + tx * 4);
    float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
  }
}

// MODE 0: forward (max / logsumexp);  MODE 1: logsumexp backward (B = Lam, bias = bk)
template <int MODE>
__global__ void __launch_bounds__(NT)
gmm_tile_kernel(const float* __restrict__ flux, PatchGeom g, const int32_t* __restrict__ shift_yx,
                const float* __restrict__ Bmat, const float* __restrict__ bias, const float* __restrict__ ck, int K,
                int marginalize, float* __restrict__ value, int32_t* __restrict__ argmax, float* __restrict__ logp,
                double* __restrict__ sum, float scale, float* __restrict__ G) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                   // 64 x AS
  float* Bs = As + PD * AS;           // 2 x 64 x 64
  float* s_bias = Bs + 2 * PD * PD;   // 2 x 64
  int* s_valid = reinterpret_cast<int*>(s_bias + 2 * PD);  // 128
  __shared__ double red[32];

  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int64_t p0 = (int64_t)blockIdx.x * TM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  stage_B(Bmat, Bs);
  if (threadIdx.x < 16) __pipeline_memcpy_async(s_bias + threadIdx.x * 4, bias + threadIdx.x * 4, 16);
  __pipeline_commit();
  gather_patches(flux, g, p0, As, s_valid);

  // running state for the 8 rows of this thread (replicated over the 16 tx lanes)
  float run_m[8], run_s[8];
  int run_k[8];
  float gacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    run_m[i] = -CUDART_INF_F;
    run_s[i] = 0.f;
    run_k[i] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) gacc[i][j] = 0.f;
  }
  float lse[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t p = p0 + ty * 8 + i;
      lse[i] = p < g.P ? value[p] : 0.f;
    }
  }

  for (int k = 0; k < K; ++k) {
    const int buf = k & 1;
    if (k + 1 < K) {
      stage_B(Bmat + (int64_t)(k + 1) * PD * PD, Bs + (buf ^ 1) * PD * PD);
      if (threadIdx.x < 16)
        __pipeline_memcpy_async(s_bias + (buf ^ 1) * PD + threadIdx.x * 4, bias + (int64_t)(k + 1) * PD + threadIdx.x * 4,
                                16);
      __pipeline_commit();
      __pipeline_wait_prior(1);
    } else {
      __pipeline_wait_prior(0);
    }
    __syncthreads();

    float acc[8][4];
    tile_mma(As, Bs + buf * PD * PD, ty, tx, acc);
    const float4 bv = *reinterpret_cast<const float4*>(s_bias + buf * PD + tx * 4);
    const float bvv[4] = {bv.x, bv.y, bv.z, bv.w};

    if (MODE == 0) {
      const float c_k = ck[k];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float y = acc[i][j] - bvv[j];
          q = fmaf(y, y, q);
        }
        q += __shfl_xor_sync(0xffffffffu, q, 8);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        float lp = fmaf(-0.5f, q, c_k);
        if (logp && tx == 0) {
          int64_t p = p0 + ty * 8 + i;
          if (p < g.P) logp[p * K + k] = lp;
        }
        if (marginalize) {
          if (lp > run_m[i]) {
            run_s[i] = run_s[i] * expf(run_m[i] - lp) + 1.f;
            run_m[i] = lp;
            run_k[i] = k;
          } else {
            run_s[i] += expf(lp - run_m[i]);
          }
        } else if (lp > run_m[i]) {  // strict: first index wins ties, as torch.max
          run_m[i] = lp;
          run_k[i] = k;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int64_t p = p0 + ty * 8 + i;
        float r = p < g.P ? expf(logp[p * K + k] - lse[i]) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) gacc[i][j] = fmaf(r, acc[i][j] - bvv[j], gacc[i][j]);
      }
    }
    __syncthreads();
  }

  if (MODE == 0) {
    double part = 0.0;
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int row = ty * 8 + i;
        int64_t p = p0 + row;
        if (p < g.P) {
          bool ok = s_valid[row] != 0;
          float v = marginalize ? run_m[i] + logf(run_s[i]) : run_m[i];
          v = ok ? v : 0.f;
          if (value) value[p] = v;
          if (argmax) argmax[p] = ok ? run_k[i] : -1;
          part += (double)v;
        }
      }
    }
    double s = block_sum(part, red);
    if (threadIdx.x == 0 && sum) atomicAdd(sum, s);
  } else {
    // G = scale * gacc, minus the row mean; invalid patches get 0
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int row = ty * 8 + i;
      int64_t p = p0 + row;
      float rs = gacc[i][0] + gacc[i][1] + gacc[i][2] + gacc[i][3];
      rs += __shfl_xor_sync(0xffffffffu, rs, 8);
      rs += __shfl_xor_sync(0xffffffffu, rs, 4);
      rs += __shfl_xor_sync(0xffffffffu, rs, 2);
      rs += __shfl_xor_sync(0xffffffffu, rs, 1);
      float mean = rs * (1.f / 64.f);
      if (p < g.P) {
        bool ok = s_valid[row] != 0;
        float4 o;
        o.x = ok ? scale * (gacc[i][0] - mean) : 0.f;
        o.y = ok ? scale * (gacc[i][1] - mean) : 0.f;
        o.z = ok ? scale * (gacc[i][2] - mean) : 0.f;
        o.w = ok ? scale * (gacc[i][3] - mean) : 0.f;
        *reinterpret_cast<float4*>(G + p * PD + tx * 4) = o;
      }
    }
  }
}

// Max-mode backward: one warp per patch, G_p = scale * (xc_p Lam_k* - bk_k*) minus row mean.
__global__ void __launch_bounds__(256)
gmm_bwd_max_kernel(const float* __restrict__ flux, PatchGeom g, const int32_t* __restrict__ shift_yx,
                   const float* __restrict__ Lam, const float* __restrict__ bk, const int32_t* __restrict__ argmax,
                   float scale, float* __restrict__ G) {
  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < g.P; p += nwarps) {
    const int k = argmax[p];
    float* out = G + p * PD;
    if (k < 0) {
      out[lane] = 0.f;
      out[lane + 32] = 0.f;
      continue;
    }
    int iy = (int)(p / g.nx) + g.row_begin, ix = (int)(p % g.nx);
    // lane holds elements d = lane (u = lane/8, v = lane%8) and d = lane + 32 (u + 4)
    int u = lane >> 3, v = lane & 7;
    int x = patch_src_col(g, ix, v);
    float x0 = __ldg(flux + (int64_t)patch_src_row(g, iy, u) * g.fW + x);
    float x1 = __ldg(flux + (int64_t)patch_src_row(g, iy, u + 4) * g.fW + x);
    float mean = warp_sum(x0 + x1) * (1.f / 64.f);
    x0 -= mean;
    x1 -= mean;
    const float* L = Lam + (int64_t)k * PD * PD;
    float g0 = -bk[(int64_t)k * PD + lane], g1 = -bk[(int64_t)k * PD + lane + 32];
#pragma unroll 8
    for (int d = 0; d < 32; ++d) {
      float xa = __shfl_sync(0xffffffffu, x0, d);
      float xb = __shfl_sync(0xffffffffu, x1, d);
      g0 = fmaf(xa, __ldg(L + d * PD + lane), g0);
      g1 = fmaf(xa, __ldg(L + d * PD + lane + 32), g1);
      g0 = fmaf(xb, __ldg(L + (d + 32) * PD + lane), g0);
      g1 = fmaf(xb, __ldg(L + (d + 32) * PD + lane + 32), g1);
    }
    float gm = warp_sum(g0 + g1) * (1.f / 64.f);
    out[lane] = scale * (g0 - gm);
    out[lane + 32] = scale * (g1 - gm);
  }
}

// Max-mode backward for upper-triangular factors, from Lw_k* itself instead of Lam_k* = Lw Lw^T:
//     y = xc Lw - mw   (y_j = sum_{i<=j}),   G = y Lw^T   (G_i = sum_{j>=i} y_j Lw[i][j]),   G -= mean(G).
// The kernel above is bound by the L2 -> SM traffic of one 16 KB Lam matrix per patch; the triangular factor
// is read once (rows i < 32: both 128-byte halves, rows i >= 32: the upper half only = 12 KB, its zero half-rows
// skipped) and used for both products from registers.  Lane l owns columns l and l + 32; the 64 row sums of the
// second product are reduced across the warp with a 62-shuffle reduce-scatter (lane l ends with rows 2l, 2l + 1).
__global__ void __launch_bounds__(256)
gmm_bwd_max_tri_kernel(const float* __restrict__ flux, PatchGeom g, const int32_t* __restrict__ shift_yx,
                       const float* __restrict__ Lw, const float* __restrict__ mw, const int32_t* __restrict__ argmax,
                       float scale, float* __restrict__ G) {
  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < g.P; p += nwarps) {
    const int k = argmax[p];
    float2* out = reinterpret_cast<float2*>(G + p * PD) + lane;
    if (k < 0) {
      *out = make_float2(0.f, 0.f);
      continue;
    }
    const int iy = (int)(p / g.nx) + g.row_begin, ix = (int)(p % g.nx);
    // lane holds patch elements d = lane (u = lane/8, v = lane%8) and d = lane + 32 (u + 4)
    const int u = lane >> 3, v = lane & 7;
    const int x = patch_src_col(g, ix, v);
    float x0 = __ldg(flux + (int64_t)patch_src_row(g, iy, u) * g.fW + x);
    float x1 = __ldg(flux + (int64_t)patch_src_row(g, iy, u + 4) * g.fW + x);
    const float mean = warp_sum(x0 + x1) * (1.f / 64.f);
    x0 -= mean;
    x1 -= mean;
    const float* L = Lw + (int64_t)k * PD * PD;
    float a[32], b[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      a[i] = __ldg(L + i * PD + lane);
      b[i] = __ldg(L + i * PD + lane + 32);
    }
#pragma unroll
    for (int i = 32; i < 64; ++i) b[i] = __ldg(L + i * PD + lane + 32);
    float y0 = -__ldg(mw + (int64_t)k * PD + lane), y1 = -__ldg(mw + (int64_t)k * PD + lane + 32);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float xa = __shfl_sync(0xffffffffu, x0, i), xb = __shfl_sync(0xffffffffu, x1, i);
      y0 = fmaf(xa, a[i], y0);
      y1 = fmaf(xa, b[i], y1);
      y1 = fmaf(xb, b[i + 32], y1);
    }
    // row partials r_i = Lw[i][l] y_l + Lw[i][l+32] y_{l+32}, in place
#pragma unroll
    for (int i = 0; i < 32; ++i) b[i] = fmaf(a[i], y0, b[i] * y1);
#pragma unroll
    for (int i = 32; i < 64; ++i) b[i] *= y1;
    // reduce-scatter over the warp: after the step with mask m a lane keeps the half of its rows selected by (lane & m)
#pragma unroll
    for (int m = 16, c = 32; m >= 1; m >>= 1, c >>= 1) {
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int t = 0; t < c; ++t) {
        const float send = up ? b[t] : b[t + c];
        const float keep = up ? b[t + c] : b[t];
        b[t] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
    const float gm = warp_sum(b[0] + b[1]) * (1.f / 64.f);
    *out = make_float2(scale * (b[0] - gm), scale * (b[1] - gm));  // rows 2 lane, 2 lane + 1
  }
}

// ---- max-mode backward, bucketed by winning component ------------------------------------------------
// The warp-per-patch kernels read one 12-16 KB matrix per patch from L2 / L1: 1 GB per launch at 65 025 patches, which is
// what they cost.  Here the patches are grouped by argmax (shared-memory histogram -> one-warp scan -> scatter), then one
// CTA per (component, chunk of <= BCH patches) stages Lam_k once and runs a 64 x 64 x 64 register-tiled product:
//     G[p, :] = scale * ((x_p - mean) Lam_k - bk_k), minus its row mean.
// Work items are not materialised: item i belongs to the component k with item_base[k] <= i < item_base[k + 1]
// (binary search in shared memory), its patches are perm[cursor0[k] + BCH (i - item_base[k]) ...].
constexpr int BCH = 64;   // patches per work item
constexpr int BXS = 68;   // padded row length of the staged patches

__global__ void __launch_bounds__(256)
bwd_hist_kernel(const int32_t* __restrict__ argmax, int P, int K, int32_t* __restrict__ counts, float* __restrict__ G) {
  extern __shared__ int32_t s_cnt[];
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
    int k = argmax[p];
    if (k >= 0 && k < K) {
      atomicAdd(s_cnt + k, 1);
    } else {  // filtered patch: zero gradient row
      float4* row = reinterpret_cast<float4*>(G + (int64_t)p * PD);
#pragma unroll
      for (int i = 0; i < 16; ++i) row[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x)
    if (s_cnt[k]) atomicAdd(counts + k, s_cnt[k]);
}

// one warp: exclusive scans of the K counts (patch offsets) and of the per-component item counts.  The counts of eight
// 32-component chunks are loaded together (one L2 round trip per 256 components instead of one per 32).
__global__ void bwd_scan_kernel(int K, int32_t* __restrict__ counts, int32_t* __restrict__ cursor0,
                                int32_t* __restrict__ cursor, int32_t* __restrict__ item_base,
                                int32_t* __restrict__ next_item) {
  const int lane = threadIdx.x;
  int off = 0, ioff = 0;
  for (int kb = 0; kb < K; kb += 256) {
    int cnt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb + 32 * j + lane;
      cnt[j] = k < K ? counts[k] : 0;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kb + 32 * j + lane;
      const int c = cnt[j];
      const int ni = (c + BCH - 1) / BCH;
      int sc = c, si = ni;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, sc, o), b2 = __shfl_up_sync(0xffffffffu, si, o);
        if (lane >= o) sc += a, si += b2;
      }
      if (k < K) {
        cursor0[k] = off + sc - c;
        cursor[k] = off + sc - c;
        item_base[k] = ioff + si - ni;
        counts[k] = 0;  // ready for the next launch
      }
      off += __shfl_sync(0xffffffffu, sc, 31);
      ioff += __shfl_sync(0xffffffffu, si, 31);
    }
  }
  if (lane == 0) {
    item_base[K] = ioff;
    *next_item = 0;  // the bucket kernel's work counter
  }
}

__global__ void __launch_bounds__(256)
bwd_scatter_kernel(const int32_t* __restrict__ argmax, int P, int K, int32_t* __restrict__ cursor,
                   int32_t* __restrict__ perm) {
  // block-local ranks first (shared-memory atomics), one global reservation per (block, component)
  extern __shared__ int32_t s_buf[];
  int32_t* s_cnt = s_buf;
  int32_t* s_base = s_buf + K;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_cnt[k] = 0;
  __syncthreads();
  constexpr int PER = 4;
  const int p0 = blockIdx.x * blockDim.x * PER;
  int kk[PER], rk[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int p = p0 + u * blockDim.x + threadIdx.x;
    kk[u] = p < P ? argmax[p] : -1;
    rk[u] = (kk[u] >= 0 && kk[u] < K) ? atomicAdd(s_cnt + kk[u], 1) : 0;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_base[k] = s_cnt[k] ? atomicAdd(cursor + k, s_cnt[k]) : 0;
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int p = p0 + u * blockDim.x + threadIdx.x;
    if (kk[u] >= 0 && kk[u] < K) perm[s_base[kk[u]] + rk[u]] = p;
  }
}

__global__ void __launch_bounds__(256)
gmm_bwd_bucket_kernel(const float* __restrict__ flux, PatchGeom g, const int32_t* __restrict__ shift_yx,
                      const float* __restrict__ Lam, const float* __restrict__ bk, int K,
                      const int32_t* __restrict__ perm, const int32_t* __restrict__ cursor0,
                      const int32_t* __restrict__ cursor_end, const int32_t* __restrict__ item_base, float scale,
                      float* __restrict__ G) {
  __shared__ __align__(16) float Ls[PD * PD];   // Lam_k, row i = input feature
  __shared__ __align__(16) float Xs[BCH * BXS];  // centred patches of the item, row = patch
  __shared__ float bs[PD];
  __shared__ int s_p[BCH];
  __shared__ int s_k, s_start, s_cnt;
  if ((int)blockIdx.x >= item_base[K]) return;
  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  if (threadIdx.x == 0) {  // component of this item: last k with item_base[k] <= item
    int lo = 0, hi = K;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (item_base[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    const int first = cursor0[lo] + BCH * ((int)blockIdx.x - item_base[lo]);
    s_k = lo;
    s_start = first;
    s_cnt = min(BCH, cursor_end[lo] - first);  // after the scatter the running cursor of k is the end of its bucket
  }
  __syncthreads();
  const int k = s_k, start = s_start, cnt = s_cnt;
  const float4* Lsrc = reinterpret_cast<const float4*>(Lam + (int64_t)k * PD * PD);
#pragma unroll
  for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(Ls)[threadIdx.x + i * 256] = __ldg(Lsrc + threadIdx.x + i * 256);
  if (threadIdx.x < PD) bs[threadIdx.x] = bk[(int64_t)k * PD + threadIdx.x];
  if (threadIdx.x < BCH) s_p[threadIdx.x] = threadIdx.x < cnt ? perm[start + threadIdx.x] : -1;
  __syncthreads();
  {  // gather: 4 threads per patch, 2 patch rows (16 values) each; mean over the quad
    const int pi = threadIdx.x >> 2, qd = threadIdx.x & 3;
    const int p = s_p[pi];
    float v[16];
    float sm = 0.f;
    if (p >= 0) {
      const int iy = p / g.nx + g.row_begin, ix = p % g.nx;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float* src = flux + (int64_t)patch_src_row(g, iy, 2 * qd + u) * g.fW;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          v[8 * u + c] = __ldg(src + patch_src_col(g, ix, c));
          sm += v[8 * u + c];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = 0.f;
    }
    sm += __shfl_xor_sync(0xffffffffu, sm, 1);
    sm += __shfl_xor_sync(0xffffffffu, sm, 2);
    const float mean = sm * (1.f / 64.f);
#pragma unroll
    for (int i = 0; i < 16; i += 4)
      *reinterpret_cast<float4*>(Xs + pi * BXS + 16 * qd + i) =
          make_float4(v[i] - mean, v[i + 1] - mean, v[i + 2] - mean, v[i + 3] - mean);
  }
  __syncthreads();
  // product: thread (ty, tx) = 4 patches (ty + 16 r) x 4 outputs (4 tx + c)
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 4
  for (int i = 0; i < PD; i += 4) {
    float4 xr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) xr[r] = *reinterpret_cast<const float4*>(Xs + (ty + 16 * r) * BXS + i);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const float4 l = *reinterpret_cast<const float4*>(Ls + (i + ii) * PD + 4 * tx);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float x = ii == 0 ? xr[r].x : ii == 1 ? xr[r].y : ii == 2 ? xr[r].z : xr[r].w;
        acc[r][0] = fmaf(x, l.x, acc[r][0]);
        acc[r][1] = fmaf(x, l.y, acc[r][1]);
        acc[r][2] = fmaf(x, l.z, acc[r][2]);
        acc[r][3] = fmaf(x, l.w, acc[r][3]);
      }
    }
  }
  const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * tx);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    acc[r][0] -= b4.x, acc[r][1] -= b4.y, acc[r][2] -= b4.z, acc[r][3] -= b4.w;
    float sg = (acc[r][0] + acc[r][1]) + (acc[r][2] + acc[r][3]);  // row sum over the 16 threads tx of a half warp
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);
    const float gm = sg * (1.f / 64.f);
    const int p = s_p[ty + 16 * r];
    if (p >= 0)
      *reinterpret_cast<float4*>(G + (int64_t)p * PD + 4 * tx) =
          make_float4(scale * (acc[r][0] - gm), scale * (acc[r][1] - gm), scale * (acc[r][2] - gm), scale * (acc[r][3] - gm));
  }
}

// Second generation of the bucket kernel.  The first one (above) is bound by its load/store wavefronts (ncu: 68 % of
// the L1 data pipe, `short_scoreboard` + `barrier` stalls): a 4 x 4 register tile needs one LDS.128 per 8 FMAs and its
// gather issues 32 scattered 4-byte loads per instruction; 1.4 waves of CTAs leave half the machine idle in the second.
// Here 128 threads own 8 patches x 4 outputs each (one LDS.128 per 10.7 FMAs, the patch operand a one-wavefront
// broadcast), a gather instruction covers 4 rows x 8 columns of ONE patch (4-8 sectors instead of 32) as 4-byte
// cp.async straight into shared memory (ptxas serialises plain loads + shuffles patch by patch: one L2 round trip
// each), Lam_k arrives through cp.async while the gather runs, and a resident grid takes items from a device counter.
constexpr int BNT2 = 128;

__device__ __forceinline__ void bwd_cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void bwd_cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(BNT2, 5)
gmm_bwd_bucket8_kernel(const float* __restrict__ flux, PatchGeom g, const int32_t* __restrict__ shift_yx,
                       const float* __restrict__ Lam, const float* __restrict__ bk, int K,
                       const int32_t* __restrict__ perm, const int32_t* __restrict__ cursor0,
                       const int32_t* __restrict__ cursor_end, const int32_t* __restrict__ item_base,
                       int32_t* __restrict__ next_item, float scale, float* __restrict__ G) {
  __shared__ __align__(16) float Ls[PD * PD];    // Lam_k, row i = input feature
  __shared__ __align__(16) float Xs[BCH * BXS];  // centred patches of the item, row = patch
  __shared__ __align__(16) float bs[PD];
  __shared__ int s_p[BCH], s_y[BCH], s_x[BCH];
  __shared__ int s_item, s_k, s_start, s_cnt;
  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int n_items = item_base[K];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(next_item, 1);
    __syncthreads();  // also: every thread is done with the previous item's shared memory
    const int item = s_item;
    if (item >= n_items) break;
    // component of this item: the one k with item_base[k] <= item < item_base[k + 1] (independent loads, one round
    // trip instead of the dependent ones of a binary search)
    for (int kk = threadIdx.x; kk < K; kk += BNT2) {
      const int b0 = item_base[kk], b1 = item_base[kk + 1];
      if (b0 <= item && item < b1) {
        const int first = cursor0[kk] + BCH * (item - b0);
        s_k = kk;
        s_start = first;
        s_cnt = min(BCH, cursor_end[kk] - first);  // after the scatter the running cursor of k is the end of its bucket
      }
    }
    __syncthreads();
    const int k = s_k, start = s_start, cnt = s_cnt;
    {
      const float* Lsrc = Lam + (int64_t)k * PD * PD;
#pragma unroll
      for (int i = 0; i < PD * PD / (4 * BNT2); ++i)
        bwd_cp_async16(Ls + 4 * (threadIdx.x + i * BNT2), Lsrc + 4 * (threadIdx.x + i * BNT2));
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (threadIdx.x < BCH) {
      const int t = threadIdx.x;
      const int p = t < cnt ? perm[start + t] : -1;
      s_p[t] = p;
      const int iy = p >= 0 ? p / g.nx + g.row_begin : 0, ix = p >= 0 ? p % g.nx : 0;
      s_y[t] = iy * g.stride - g.sy;  // same arithmetic as patch_src_row / patch_src_col, the element offset added below
      s_x[t] = ix * g.stride - g.sx;
      bs[t] = bk[(int64_t)k * PD + t];
    }
    __syncthreads();
    {  // gather: a warp takes 16 patches; lane = (row r of a half patch, column c): two copies cover the 8 x 8 patch
      const int r = lane >> 3, c = lane & 7;
      constexpr int PW = BCH / (BNT2 / 32);
#pragma unroll 8
      for (int j = 0; j < PW; ++j) {
        const int pi = PW * w + j;
        float* dst = Xs + pi * BXS + 8 * r + c;
        if (s_p[pi] >= 0) {
          const int y0 = s_y[pi], col = wrap(s_x[pi] + c, g.fW);
          bwd_cp_async4(dst, flux + (int64_t)wrap(y0 + r, g.fH) * g.fW + col);
          bwd_cp_async4(dst + 32, flux + (int64_t)wrap(y0 + r + 4, g.fH) * g.fW + col);
        } else {
          dst[0] = 0.f;
          dst[32] = 0.f;
        }
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");  // also Lam_k
#pragma unroll 8
      for (int j = 0; j < PW; ++j) {  // every lane centres the two elements it copied itself
        float* dst = Xs + (PW * w + j) * BXS + 8 * r + c;
        const float a = dst[0], b = dst[32];
        const float mean = warp_sum(a + b) * (1.f / 64.f);
        dst[0] = a - mean;
        dst[32] = b - mean;
      }
    }
    __syncthreads();
    // product: thread (ty, tx) = 8 patches (ty + 8 r) x 4 outputs (4 tx + c)
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
#pragma unroll 2
    for (int i = 0; i < PD; i += 4) {
      float4 xr[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) xr[r] = *reinterpret_cast<const float4*>(Xs + (ty + 8 * r) * BXS + i);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const float4 l = *reinterpret_cast<const float4*>(Ls + (i + ii) * PD + 4 * tx);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float x = ii == 0 ? xr[r].x : ii == 1 ? xr[r].y : ii == 2 ? xr[r].z : xr[r].w;
          acc[r][0] = fmaf(x, l.x, acc[r][0]);
          acc[r][1] = fmaf(x, l.y, acc[r][1]);
          acc[r][2] = fmaf(x, l.z, acc[r][2]);
          acc[r][3] = fmaf(x, l.w, acc[r][3]);
        }
      }
    }
    const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * tx);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      acc[r][0] -= b4.x, acc[r][1] -= b4.y, acc[r][2] -= b4.z, acc[r][3] -= b4.w;
      float sg = (acc[r][0] + acc[r][1]) + (acc[r][2] + acc[r][3]);
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o);  // the 16 threads tx of a patch row
      const float gm = sg * (1.f / 64.f);
      const int p = s_p[ty + 8 * r];
      if (p >= 0)
        *reinterpret_cast<float4*>(G + (int64_t)p * PD + 4 * tx) =
            make_float4(scale * (acc[r][0] - gm), scale * (acc[r][1] - gm), scale * (acc[r][2] - gm), scale * (acc[r][3] - gm));
    }
  }
}

// Raw patch extraction (cycle_spin roll + view_as_overlapping_patches_torch): X[p', 8u+v].
__global__ void extract_patches_kernel(const float* __restrict__ flux, PatchGeom g,
                                       const int32_t* __restrict__ shift_yx, float* __restrict__ X) {
  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int64_t n = (int64_t)g.P * PD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t p = i >> 6;
    int d = (int)(i & 63);
    int iy = (int)(p / g.nx) + g.row_begin, ix = (int)(p % g.nx);
    X[i] = flux[(int64_t)patch_src_row(g, iy, d >> 3) * g.fW + patch_src_col(g, ix, d & 7)];
  }
}

// Generic-D log-prob: one warp per (sample, component).
__global__ void gmm_log_prob_kernel(const float* __restrict__ x, int64_t P, int D, int K, const float* __restrict__ Lw,
                                    const float* __restrict__ mw, const float* __restrict__ ck,
                                    float* __restrict__ logp) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < P * K; w += nwarps) {
    int64_t p = w / K;
    int k = (int)(w - p * K);
    const float* xp = x + p * D;
    const float* L = Lw + (int64_t)k * D * D;
    float q = 0.f;
    for (int j = lane; j < D; j += 32) {
      float y = -mw[(int64_t)k * D + j];
      for (int d = 0; d < D; ++d) y = fmaf(xp[d], L[(int64_t)d * D + j], y);
      q = fmaf(y, y, q);
    }
    q = warp_sum(q);
    if (lane == 0) logp[w] = fmaf(-0.5f, q, ck[k]);
  }
}

static int make_geom(const char* name, int fH, int fW, int stride, int row_begin, int row_end, PatchGeom* g) {
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH, "%s: image %dx%d smaller than a patch", name, fH, fW);
  JD_CHECK_ARG(stride >= 1 && stride <= PATCH, "%s: stride %d outside [1,8]", name, stride);
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end, "%s: bad patch-row block [%d,%d) of %d", name,
               row_begin, row_end, ny);
  g->fH = fH;
  g->fW = fW;
  g->sy = 0;
  g->sx = 0;
  g->stride = stride;
  g->nx = nx;
  g->row_begin = row_begin;
  g->P = (row_end - row_begin) * nx;
  return JD_OK;
}

constexpr size_t TILE_SMEM = (size_t)(PD * AS + 2 * PD * PD + 2 * PD) * sizeof(float) + TM * sizeof(int);

int gmm_prior_forward_simt(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* Lw, const float* mw, const float* ck, int K, int marginalize,
                           float* value, int32_t* argmax, float* logp, double* sum, cudaStream_t st) {
  PatchGeom g;
  int rc = make_geom("jd_gmm_prior_forward", fH, fW, stride, row_begin, row_end, &g);
  if (rc) return rc;
  auto kern = gmm_tile_kernel<0>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM);
  }
  int grid = (g.P + TM - 1) / TM;
  kern<<<grid, NT, TILE_SMEM, st>>>(flux, g, shift_yx, Lw, mw, ck, K, marginalize, value, argmax, logp, sum, 0.f,
                                    nullptr);
  JD_CHECK_LAUNCH("jd_gmm_prior_forward(simt)");
  return JD_OK;
}

}  // namespace jd

using namespace jd;

extern "C" {

int jd_gmm_log_prob(const float* x, int64_t P, int D, int K, const float* Lw, const float* mw, const float* ck,
                    float* logp, jd_stream_t stream) {
  JD_CHECK_ARG(x && Lw && mw && ck && logp, "jd_gmm_log_prob: null pointer");
  JD_CHECK_ARG(P > 0 && K > 0 && D > 0 && D <= 1024, "jd_gmm_log_prob: bad shape P=%lld D=%d K=%d", (long long)P, D, K);
  int64_t warps = P * K;
  int64_t blocks = (warps + 7) / 8;
  int64_t cap = (int64_t)num_sms() * 16;
  gmm_log_prob_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, to_stream(stream)>>>(x, P, D, K, Lw, mw, ck, logp);
  JD_CHECK_LAUNCH("jd_gmm_log_prob");
  return JD_OK;
}

int jd_extract_patches(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                       int row_end, float* X, jd_stream_t stream) {
  JD_CHECK_ARG(flux && X, "jd_extract_patches: null pointer");
  PatchGeom g;
  int rc = make_geom("jd_extract_patches", fH, fW, stride, row_begin, row_end, &g);
  if (rc) return rc;
  int64_t n = (int64_t)g.P * PD;
  int64_t blocks = (n + 255) / 256, cap = (int64_t)num_sms() * 16;
  extract_patches_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, to_stream(stream)>>>(flux, g, shift_yx, X);
  JD_CHECK_LAUNCH("jd_extract_patches");
  return JD_OK;
}

int64_t jd_gmm_backward_workspace_elems(int64_t P, int K) {
  // counts[K] (zero at entry, left at zero) cursor0[K] cursor[K] item_base[K + 1] next_item[1] perm[P]
  return 4 * (int64_t)K + 2 + P;
}

int jd_gmm_prior_backward(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                          int row_end, const float* Lam, const float* bk, int K, int marginalize,
                          const int32_t* argmax, const float* logp, const float* value, float scale, float* G,
                          int32_t* workspace, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Lam && bk && G && K > 0, "jd_gmm_prior_backward: null pointer");
  PatchGeom g;
  int rc = make_geom("jd_gmm_prior_backward", fH, fW, stride, row_begin, row_end, &g);
  if (rc) return rc;
  cudaStream_t st = to_stream(stream);
  if (marginalize) {
    JD_CHECK_ARG(logp && value, "jd_gmm_prior_backward: marginalize=1 needs logp and value from the forward");
    auto kern = gmm_tile_kernel<1>;
    static bool attr_set[64] = {};
    if (first_use_on_device(attr_set)) {
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM);
    }
    int grid = (g.P + TM - 1) / TM;
    kern<<<grid, NT, TILE_SMEM, st>>>(flux, g, shift_yx, Lam, bk, nullptr, K, 1, const_cast<float*>(value), nullptr,
                                      const_cast<float*>(logp), nullptr, scale, G);
  } else {
    JD_CHECK_ARG(argmax, "jd_gmm_prior_backward: marginalize=0 needs argmax from the forward");
    if (workspace) {
      JD_CHECK_ARG(K <= 4096, "jd_gmm_prior_backward: the bucketed backward supports up to 4096 components");
      JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Lam) & 15) == 0 && (reinterpret_cast<uintptr_t>(G) & 15) == 0,
                   "jd_gmm_prior_backward: Lam and G must be 16-byte aligned for the bucketed backward");
      int32_t* counts = workspace;
      int32_t* cursor0 = counts + K;
      int32_t* cursor = cursor0 + K;
      int32_t* item_base = cursor + K;
      int32_t* next_item = item_base + K + 1;
      int32_t* perm = next_item + 1;
      const int max_items = K + g.P / BCH + 1;
      const int hb = (int)std::min<int64_t>(((int64_t)g.P + 1023) / 1024, (int64_t)num_sms() * 4);
      bwd_hist_kernel<<<hb, 256, sizeof(int32_t) * K, st>>>(argmax, g.P, K, counts, G);
      bwd_scan_kernel<<<1, 32, 0, st>>>(K, counts, cursor0, cursor, item_base, next_item);
      bwd_scatter_kernel<<<(g.P + 1023) / 1024, 256, 2 * sizeof(int32_t) * K, st>>>(argmax, g.P, K, cursor, perm);
      static const bool v1 = getenv("JD_BWD_BUCKET_V1") && atoi(getenv("JD_BWD_BUCKET_V1")) != 0;  // A/B knob
      if (v1)
        gmm_bwd_bucket_kernel<<<max_items, 256, 0, st>>>(flux, g, shift_yx, Lam, bk, K, perm, cursor0, cursor, item_base,
                                                         scale, G);
      else
        gmm_bwd_bucket8_kernel<<<std::min(max_items, num_sms() * 5), BNT2, 0, st>>>(
            flux, g, shift_yx, Lam, bk, K, perm, cursor0, cursor, item_base, next_item, scale, G);
    } else {
      int64_t blocks = ((int64_t)g.P + 7) / 8;
      int64_t cap = (int64_t)num_sms() * 16;
      gmm_bwd_max_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, st>>>(flux, g, shift_yx, Lam, bk, argmax, scale,
                                                                             G);
    }
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_backward");
  return JD_OK;
}

int jd_gmm_prior_backward_max_tri(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                                  int row_end, const float* Lw, const float* mw, int K, const int32_t* argmax,
                                  float scale, float* G, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Lw && mw && argmax && G && K > 0, "jd_gmm_prior_backward_max_tri: null pointer");
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(G) & 7) == 0, "jd_gmm_prior_backward_max_tri: G must be 8-byte aligned");
  PatchGeom g;
  int rc = make_geom("jd_gmm_prior_backward_max_tri", fH, fW, stride, row_begin, row_end, &g);
  if (rc) return rc;
  int64_t blocks = ((int64_t)g.P + 7) / 8;
  int64_t cap = (int64_t)num_sms() * 8;
  gmm_bwd_max_tri_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, to_stream(stream)>>>(flux, g, shift_yx, Lw, mw,
                                                                                           argmax, scale, G);
  JD_CHECK_LAUNCH("jd_gmm_prior_backward_max_tri");
  return JD_OK;
}

}  // extern "C"
