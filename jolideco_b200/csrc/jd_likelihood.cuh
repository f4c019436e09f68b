// Batched likelihood kernels of the MAP step (sm_100a): the NPred forward model of every dataset of a joint
// iteration in ONE launch (grid.z = dataset) with the Poisson cash statistic and its gradient fused into the
// convolution epilogue, and the adjoint in a second launch.
//
//   forward  (models/npred.py:160-191, 234-261; loss.py:35-37):
//       conv = PSF (*) (flux . exposure)      direct form of rfft2*rfft2 -> irfft2 -> centred crop (utils/torch.py:337-370)
//       pool = sum over f x f blocks of conv  (registers: a thread owns whole blocks)
//       npred = max(pool, 0) + B exp(log b);  loss += npred - c log(npred + eps);  dpool = (1 - c/(npred+eps)) / (H W)
//     conv / npred never go to memory: 8 B/px read (c, B) + 4 B/px written (dpool) + the flux / exposure tiles.
//   backward:
//       dflux (+)= exposure . (PSF (*)^T up_f(dpool))
//
// Both directions are the offset correlation out[i,j] = sum_{a,b} Kc[a,b] in[i+oy+a, j+ox+b] of jd_conv.cu
// (forward: Kc = flip(psf), (oy,ox) = (s-(k-1)); adjoint: Kc = psf, (oy,ox) = -s; s = (k-1)/2), which keeps the
// asymmetric crop of even PSFs.
//
// Thread = RT x 8 outputs (RT = 4 for f = 1: 32 accumulators, 32 x 64 tiles; RT = 8 for f = 2 and rows of more than 32
// taps), CTA = 8 x 8 threads.  The input tile
// (64+kh-1) x (64+4 KG) and the taps live in shared memory; per staged input row t a thread loads its sliding
// window once (<= 10 LDS.128) and feeds the up to 8 output rows r with tap row a = t - r: 32 FFMA per broadcast
// LDS.128 of four taps -> the FMA pipe, not the LSU, is the limiter (the 4 x 4 tile of jd_conv.cu's conv3 kernel
// spends 5 LDS per 64 FFMA).  Tap rows hold `lead` (< 4) zero taps in front so that the tile origin is 16-byte
// aligned in global memory, KG groups of four taps, the last group with KT real taps (exact tap count at compile
// time: no multiply-by-zero work at the end of a row).  Shared-memory columns are skewed by 4 floats per 32
// (pcol) so that the 8 threads of a quarter warp, 8 columns apart, hit 8 distinct 16-byte bank groups.
#pragma once
#include "jd_common.cuh"

namespace jd {
namespace lik {

constexpr int CT = 8;                // output columns per thread; rows per thread: template parameter RT (8, or 4 for
                                     // launches that would leave the SMs with a couple of warps each)
constexpr int TYN = 8, TXN = 8;      // threads per CTA
constexpr int NTHR = TYN * TXN;      // 64
constexpr int TW = CT * TXN;         // 64 output columns per CTA, RT * TYN = 64 or 32 rows
enum { FWD = 0, BWD = 1 };

__host__ __device__ __forceinline__ int pcol(int c) { return c + ((c >> 5) << 2); }

template <int KG, int RT = 8>
struct Tile {
  static constexpr int TH = RT * TYN;
  static constexpr int KWP = 4 * KG;                                // tap row length in shared memory
  static constexpr int IW = TW + KWP;                               // staged input row, logical floats
  static constexpr int IWP = IW + (((IW - 1) >> 5) << 2);           // physical (skewed) row length, multiple of 4
  static constexpr int NVEC = IW / 4;
  static __host__ __device__ size_t smem_bytes(int kh) { return ((size_t)(TH + kh - 1) * IWP + (size_t)kh * KWP) * 4; }
};

// acc[r][c] += sum_{a,b} tap[a][b] * tile[ty*8 + r + a][tx*8 + c + b]
template <int KG, int KT, int RT>
__device__ __forceinline__ void conv_core(const float* __restrict__ s_in, const float* __restrict__ s_k, int kh,
                                          int ty, int tx, float (&acc)[RT][CT]) {
  using T = Tile<KG, RT>;
  constexpr int NWF = CT + 4 * (KG - 1) + KT - 1;  // window floats a thread needs per input row
  constexpr int NW = (NWF + 3) / 4;
  int off[NW];
#pragma unroll
  for (int j = 0; j < NW; ++j) off[j] = pcol(tx * CT + 4 * j);
  const float* rowp = s_in + (ty * RT) * T::IWP;
  const int nt = RT + kh - 1;
#pragma unroll 1
  for (int t = 0; t < nt; ++t, rowp += T::IWP) {
    float win[4 * NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(rowp + off[j]);
      win[4 * j] = v.x, win[4 * j + 1] = v.y, win[4 * j + 2] = v.z, win[4 * j + 3] = v.w;
    }
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int a = t - r;
      if ((unsigned)a < (unsigned)kh) {  // warp-uniform
        const float* kr = s_k + a * T::KWP;
#pragma unroll
        for (int g = 0; g < KG; ++g) {
          const float4 kv = *reinterpret_cast<const float4*>(kr + 4 * g);
          const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int bb = 0; bb < (g == KG - 1 ? KT : 4); ++bb)
#pragma unroll
            for (int c = 0; c < CT; ++c) acc[r][c] = fmaf(kk[bb], win[4 * g + bb + c], acc[r][c]);
        }
      }
    }
  }
}

// Poisson cash statistic of one pixel without the counts-only Stirling term (a constant of the dataset, added once
// from jd_lik_dataset.loss_const).  MUFU-based log / reciprocal: relative error ~3e-7 on c log(n+eps) and c/(n+eps).
__device__ __forceinline__ float poisson_px(float pool, float bkg, float c, float eps, float gs, float& lacc,
                                            float& bacc) {
  const float np_ = fmaxf(pool, 0.f) + bkg;
  const float ne = np_ + eps;
  lacc += np_ - c * __logf(ne);
  const float d = (1.f - __fdividef(c, ne)) * gs;
  bacc = fmaf(d, bkg, bacc);
  return pool >= 0.f ? d : 0.f;
}

template <int MODE, int F, int KG, int KT, int RT = 8>
__global__ void __launch_bounds__(NTHR)
lik_kernel(const jd_lik_dataset* __restrict__ table, int fH, int fW, int kh, int kw, int H, int W, float eps,
           float grad_scale) {
  using T = Tile<KG, RT>;
  constexpr int TH = T::TH;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_k = smem + (TH + kh - 1) * T::IWP;
  __shared__ float s_red[4];

  const jd_lik_dataset ds = table[blockIdx.z];
  const int tid = threadIdx.x, tx = tid & (TXN - 1), ty = tid / TXN;
  const int sy = (kh - 1) / 2, sx = (kw - 1) / 2;
  const int oy = MODE == FWD ? sy - (kh - 1) : -sy;
  const int ox = MODE == FWD ? sx - (kw - 1) : -sx;
  const int lead = ((ox % 4) + 4) % 4;
  const int tile_y = blockIdx.y * TH, tile_x = blockIdx.x * TW;
  const int Y0 = tile_y + oy, X0 = tile_x + ox - lead;  // X0 % 4 == 0

  // ---- taps: s_k[a][lead + b] = Kc[a][b]
  for (int i = tid; i < kh * T::KWP; i += NTHR) {
    const int a = i / T::KWP, b = i - a * T::KWP - lead;
    float v = 0.f;
    if (b >= 0 && b < kw) v = MODE == FWD ? __ldg(ds.psf + (kh - 1 - a) * kw + (kw - 1 - b)) : __ldg(ds.psf + a * kw + b);
    s_k[i] = v;
  }

  // ---- input tile: rows [Y0, Y0 + TH + kh - 1), cols [X0, X0 + IW), zero outside the image.  Register-staged in
  // batches so that every thread keeps 16 independent 128-bit loads in flight (the tile load is pure latency).
  const int ih = TH + kh - 1;
  constexpr int SB = 8;  // staging batch: SB (x2 in the forward) independent 128-bit loads in flight per thread
  const int total = ih * T::NVEC;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (MODE == FWD) {
    const float* __restrict__ in = ds.flux;
    const float* __restrict__ sc = ds.exposure;
    const bool vec = (fW & 3) == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(sc)) & 15) == 0;
    if (vec) {
      for (int base = tid; base < total; base += SB * NTHR) {
        float4 fv[SB], ev[SB];
#pragma unroll
        for (int u = 0; u < SB; ++u) {
          const int i = base + u * NTHR;
          const int ry = i / T::NVEC, v = i - ry * T::NVEC;
          const int y = Y0 + ry, x = X0 + 4 * v;
          const bool ok = i < total && y >= 0 && y < fH && x >= 0 && x < fW;
          const int64_t o = ok ? (int64_t)y * fW + x : 0;
          fv[u] = ok ? __ldg(reinterpret_cast<const float4*>(in + o)) : zero4;
          ev[u] = (ok && sc) ? __ldg(reinterpret_cast<const float4*>(sc + o)) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int u = 0; u < SB; ++u) {
          const int i = base + u * NTHR;
          if (i < total) {
            const int ry = i / T::NVEC, v = i - ry * T::NVEC;
            *reinterpret_cast<float4*>(s_in + ry * T::IWP + pcol(4 * v)) =
                make_float4(fv[u].x * ev[u].x, fv[u].y * ev[u].y, fv[u].z * ev[u].z, fv[u].w * ev[u].w);
          }
        }
      }
    } else {
      for (int i = tid; i < total; i += NTHR) {
        const int ry = i / T::NVEC, v = i - ry * T::NVEC;
        const int y = Y0 + ry, x = X0 + 4 * v;
        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
        if (y >= 0 && y < fH) {
          const int64_t o = (int64_t)y * fW + x;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int xx = x + e;
            if (xx >= 0 && xx < fW) tmp[e] = __ldg(in + o + e) * (sc ? __ldg(sc + o + e) : 1.f);
          }
        }
        *reinterpret_cast<float4*>(s_in + ry * T::IWP + pcol(4 * v)) = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
    }
  } else {
    const float* __restrict__ in = ds.dpool;  // H x W, replicated f x f (adjoint of the sum-pool)
    const bool vec = F == 1 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (vec) {
      for (int base = tid; base < total; base += 2 * SB * NTHR) {
        float4 dv[2 * SB];
#pragma unroll
        for (int u = 0; u < 2 * SB; ++u) {
          const int i = base + u * NTHR;
          const int ry = i / T::NVEC, v = i - ry * T::NVEC;
          const int y = Y0 + ry, x = X0 + 4 * v;
          const bool ok = i < total && y >= 0 && y < H && x >= 0 && x < W;
          dv[u] = ok ? __ldg(reinterpret_cast<const float4*>(in + (int64_t)y * W + x)) : zero4;
        }
#pragma unroll
        for (int u = 0; u < 2 * SB; ++u) {
          const int i = base + u * NTHR;
          if (i < total) {
            const int ry = i / T::NVEC, v = i - ry * T::NVEC;
            *reinterpret_cast<float4*>(s_in + ry * T::IWP + pcol(4 * v)) = dv[u];
          }
        }
      }
    } else {
      for (int i = tid; i < total; i += NTHR) {
        const int ry = i / T::NVEC, v = i - ry * T::NVEC;
        const int y = Y0 + ry, x = X0 + 4 * v;
        float tmp[4] = {0.f, 0.f, 0.f, 0.f};
        if (y >= 0 && y < fH) {
          const int py = y / F;
          if (py < H) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int xx = x + e;
              const int px = xx / F;
              if (xx >= 0 && xx < fW && px < W) tmp[e] = __ldg(in + (int64_t)py * W + px);
            }
          }
        }
        *reinterpret_cast<float4*>(s_in + ry * T::IWP + pcol(4 * v)) = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
    }
  }
  __syncthreads();

  float acc[RT][CT];
#pragma unroll
  for (int r = 0; r < RT; ++r)
#pragma unroll
    for (int c = 0; c < CT; ++c) acc[r][c] = 0.f;
  conv_core<KG, KT, RT>(s_in, s_k, kh, ty, tx, acc);

  const int y0 = tile_y + ty * RT, x0 = tile_x + tx * CT;
  if (MODE == BWD) {
    // ---- dflux (+)= exposure . acc
    float* __restrict__ out = ds.dflux;
    const float* __restrict__ sc = ds.exposure;
    const bool vec = (fW & 3) == 0 && x0 + CT <= fW &&
                     ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(sc)) & 15) == 0;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const int y = y0 + r;
      if (y >= fH) break;
      const int64_t o = (int64_t)y * fW + x0;
      if (vec) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 val = make_float4(acc[r][4 * h], acc[r][4 * h + 1], acc[r][4 * h + 2], acc[r][4 * h + 3]);
          if (sc) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(sc + o + 4 * h));
            val.x *= e.x, val.y *= e.y, val.z *= e.z, val.w *= e.w;
          }
          if (ds.accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(out + o + 4 * h);
            val.x += old.x, val.y += old.y, val.z += old.z, val.w += old.w;
          }
          *reinterpret_cast<float4*>(out + o + 4 * h) = val;
        }
      } else {
#pragma unroll
        for (int c = 0; c < CT; ++c) {
          if (x0 + c >= fW) continue;
          float val = acc[r][c];
          if (sc) val *= __ldg(sc + o + c);
          if (ds.accumulate) val += out[o + c];
          out[o + c] = val;
        }
      }
    }
    return;
  }

  // ---- forward epilogue: sum-pool in registers, Poisson loss + gradient
  constexpr int PR = RT / F, PC = CT / F;  // pooled pixels per thread
  const float bnorm = ds.bkg_log_norm ? expf(__ldg(ds.bkg_log_norm)) : 1.f;
  float lacc = 0.f, bacc = 0.f;
  const int py0 = y0 / F, px0 = x0 / F;
  const bool vec = (W & 3) == 0 && px0 + PC <= W && (PC & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(ds.counts) | reinterpret_cast<uintptr_t>(ds.background) |
                     reinterpret_cast<uintptr_t>(ds.dpool)) & 15) == 0;
#pragma unroll
  for (int i = 0; i < PR; ++i) {
    const int py = py0 + i;
    if (py >= H) break;
    float pool[PC];
#pragma unroll
    for (int j = 0; j < PC; ++j) {
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < F; ++u)
#pragma unroll
        for (int v = 0; v < F; ++v) s += acc[F * i + u][F * j + v];
      pool[j] = s;
    }
    const int64_t o = (int64_t)py * W + px0;
    if (vec) {
#pragma unroll
      for (int h = 0; h < PC / 4; ++h) {
        const float4 cv = __ldg(reinterpret_cast<const float4*>(ds.counts + o + 4 * h));
        const float4 bv = __ldg(reinterpret_cast<const float4*>(ds.background + o + 4 * h));
        float4 dv;
        dv.x = poisson_px(pool[4 * h], bv.x * bnorm, cv.x, eps, grad_scale, lacc, bacc);
        dv.y = poisson_px(pool[4 * h + 1], bv.y * bnorm, cv.y, eps, grad_scale, lacc, bacc);
        dv.z = poisson_px(pool[4 * h + 2], bv.z * bnorm, cv.z, eps, grad_scale, lacc, bacc);
        dv.w = poisson_px(pool[4 * h + 3], bv.w * bnorm, cv.w, eps, grad_scale, lacc, bacc);
        if (ds.dpool) *reinterpret_cast<float4*>(ds.dpool + o + 4 * h) = dv;
      }
    } else {
#pragma unroll
      for (int j = 0; j < PC; ++j) {
        if (px0 + j >= W) continue;
        const float d = poisson_px(pool[j], __ldg(ds.background + o + j) * bnorm, __ldg(ds.counts + o + j), eps,
                                   grad_scale, lacc, bacc);
        if (ds.dpool) ds.dpool[o + j] = d;
      }
    }
  }
  lacc = warp_sum(lacc);
  bacc = warp_sum(bacc);
  if ((tid & 31) == 0) s_red[(tid >> 5) * 2] = lacc, s_red[(tid >> 5) * 2 + 1] = bacc;
  __syncthreads();
  if (tid == 0) {
    double l = (double)s_red[0] + (double)s_red[2];
    if (blockIdx.x == 0 && blockIdx.y == 0) l += ds.loss_const;
    if (ds.loss_sum) atomicAdd(ds.loss_sum, l);
    if (ds.dlogb) atomicAdd(ds.dlogb, (double)s_red[1] + (double)s_red[3]);
  }
}

// host side: launch one instantiation
template <int MODE, int F, int KG, int KT, int RT = 8>
int launch(const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W, float eps,
           float grad_scale, cudaStream_t st) {
  using T = Tile<KG, RT>;
  constexpr int TH = T::TH;
  const size_t sm = T::smem_bytes(kh);
  JD_CHECK_ARG(sm <= 200 * 1024, "jd_likelihood: PSF too tall for the direct kernel (kh=%d)", kh);
  auto kern = lik_kernel<MODE, F, KG, KT, RT>;
  static bool attr_set[64] = {};  // per device: opt in to > 48 KB of dynamic shared memory
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set[dev & 63] = true;
  }
  dim3 grid((fW + TW - 1) / TW, (fH + TH - 1) / TH, n_datasets);
  kern<<<grid, NTHR, sm, st>>>(table, fH, fW, kh, kw, H, W, eps, grad_scale);
  JD_CHECK_LAUNCH("jd_likelihood");
  return JD_OK;
}

}  // namespace lik
}  // namespace jd
