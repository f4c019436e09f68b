// Direct (shared-memory tiled) PSF convolution of the NPred forward model and its adjoint.
//
// Both directions are one "offset correlation"
//     out[i,j] = sum_{a<kh, b<kw} Kc[a,b] * in[i+oy+a, j+ox+b]        (in = 0 outside the image)
//   forward : Kc = flip(psf), (oy,ox) = (sy-(kh-1), sx-(kw-1)), in = flux*exposure      s = (k-1)/2
//   adjoint : Kc = psf,       (oy,ox) = (-sy, -sx),             in[y,x] = dpool[y/f, x/f], out *= E
// which reproduces rfft2*rfft2 -> irfft2 -> centred crop of utils/torch.py:337-370 exactly, including
// the asymmetric crop of even-sized PSFs.
//
// Tiling: a CTA of TY x TX threads owns an (R TY) x (4 TX) output tile; each thread an R x 4 register
// block (R = 2 by default).  The PSF is processed in chunks of KC rows: per chunk the input tile
// (4 TY + KC - 1) x (4 TX + kw_pad - 1) and the chunk rows are staged in shared memory.  For every
// staged input row t the thread loads a sliding 4+4 window once and feeds the 4 output rows r with
// kernel row a = t - r (zero rows pad the chunk so no predicate is needed): 64 FMA per
// 1 window LDS.128 + 4 broadcast LDS.128.
#include <stdlib.h>

#include "jd_common.cuh"

namespace jd {

constexpr int C = 4;   // output cols per thread
constexpr int KC = 16; // PSF rows per chunk

enum { CONV_FWD = 0, CONV_BWD = 1 };

template <int MODE, int TY, int TX, int R>  // R = output rows per thread
__global__ void __launch_bounds__(TY * TX)
conv_kernel(const float* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ psf,
            float* __restrict__ out, int fH, int fW, int kh, int kw, int oy, int ox, int f, int H, int W,
            int accumulate, int kchunk) {
  constexpr int TH = R * TY, TW = C * TX;
  constexpr int NT = TY * TX;
  extern __shared__ __align__(16) float smem[];
  const int kwp = (kw + 3) & ~3;            // kernel row padded to a multiple of 4
  const int iw = TW + kwp;                  // staged input row length (multiple of 4)
  const int ih = TH + kchunk - 1;
  float* s_in = smem;                       // ih x iw
  float* s_k = smem + ih * iw;              // (kchunk + 2(R-1)) x kwp, rows [R-1, R-1+kchunk) hold the chunk

  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int tile_y = blockIdx.y * TH, tile_x = blockIdx.x * TW;
  // staging layout: 32 consecutive threads walk a row (coalesced), NT/32 rows at a time
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  constexpr int LROWS = NT / 32;

  float acc[R][C];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) acc[r][c] = 0.f;

  // zero the kernel staging buffer once: pad rows/cols stay zero for every chunk
  for (int i = threadIdx.x; i < (kchunk + 2 * (R - 1)) * kwp; i += NT) s_k[i] = 0.f;

  for (int a0 = 0; a0 < kh; a0 += kchunk) {
    const int kc = min(kchunk, kh - a0);
    __syncthreads();
    // stage kernel chunk (MODE fwd: flipped psf); rows >= kc of the chunk are zeroed
    for (int a = ly; a < kchunk; a += LROWS)
      for (int b = lx; b < kwp; b += 32) {
        float val = 0.f;
        if (a < kc && b < kw) {
          int aa = a0 + a;
          val = MODE == CONV_FWD ? __ldg(psf + (kh - 1 - aa) * kw + (kw - 1 - b)) : __ldg(psf + aa * kw + b);
        }
        s_k[(a + R - 1) * kwp + b] = val;
      }
    // stage input tile rows [tile_y + oy + a0, +TH+kc-1), cols [tile_x + ox, +iw): 4 rows per pass so that
    // 4 (8 with the exposure) independent global loads are in flight per thread
    const int rows = TH + kc - 1;
    const int y0 = tile_y + oy + a0, x0 = tile_x + ox;
    for (int rb = ly; rb < rows; rb += 4 * LROWS) {
      for (int rx = lx; rx < iw; rx += 32) {
        const int x = x0 + rx;
        const bool xin = x >= 0 && x < fW;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ry = rb + j * LROWS, y = y0 + ry;
          float val = 0.f;
          if (ry < rows && xin && y >= 0 && y < fH) {
            if (MODE == CONV_FWD) {
              val = __ldg(in + (int64_t)y * fW + x);
              if (scale) val *= __ldg(scale + (int64_t)y * fW + x);
            } else {
              int py = y / f, px = x / f;
              if (py < H && px < W) val = __ldg(in + (int64_t)py * W + px);
            }
          }
          v[j] = val;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int ry = rb + j * LROWS;
          if (ry < rows) s_in[ry * iw + rx] = v[j];
        }
      }
    }
    __syncthreads();

    // t = staged input row relative to the thread's first output row
    for (int t = 0; t < R - 1 + kc; ++t) {
      const float* in_row = s_in + (ty * R + t) * iw + tx * C;
      const float* k_rows = s_k + (t + R - 1) * kwp;   // row for r = 0; row for r is k_rows - r*kwp
      float4 lo = *reinterpret_cast<const float4*>(in_row);
#pragma unroll 2
      for (int b0 = 0; b0 < kwp; b0 += 4) {
        float4 hi = *reinterpret_cast<const float4*>(in_row + b0 + 4);
        float win[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float4 kv = *reinterpret_cast<const float4*>(k_rows - r * kwp + b0);
          float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
          for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[r][c] = fmaf(kk[bb], win[bb + c], acc[r][c]);
        }
        lo = hi;
      }
    }
  }

#pragma unroll
  for (int r = 0; r < R; ++r) {
    int y = tile_y + ty * R + r;
    if (y >= fH) continue;
    int x = tile_x + tx * C;
    int64_t o = (int64_t)y * fW + x;
    if (x + C <= fW && (fW & 3) == 0) {
      float4 val = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      if (MODE == CONV_BWD) {
        if (scale) {
          float4 sc = *reinterpret_cast<const float4*>(scale + o);
          val.x *= sc.x, val.y *= sc.y, val.z *= sc.z, val.w *= sc.w;
        }
        if (accumulate) {
          float4 old = *reinterpret_cast<const float4*>(out + o);
          val.x += old.x, val.y += old.y, val.z += old.z, val.w += old.w;
        }
      }
      *reinterpret_cast<float4*>(out + o) = val;
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (x + c >= fW) continue;
        float val = acc[r][c];
        if (MODE == CONV_BWD) {
          if (scale) val *= scale[o + c];
          if (accumulate) val += out[o + c];
        }
        out[o + c] = val;
      }
    }
  }
}

template <int MODE, int TY, int TX, int R>
static int launch_conv_t(const float* in, const float* scale, const float* psf, float* out, int fH, int fW, int kh,
                         int kw, int oy, int ox, int f, int H, int W, int accumulate, cudaStream_t st,
                         const char* name) {
  const int kwp = (kw + 3) & ~3;
  // whole PSF in one chunk when it fits ~44 KB of shared memory (several CTAs stay resident per SM);
  // otherwise chunks of KC rows
  auto smem_for = [&](int kc) {
    return ((size_t)(R * TY + kc - 1) * (C * TX + kwp) + (size_t)(kc + 2 * (R - 1)) * kwp) * sizeof(float);
  };
  int kchunk = kh;
  if (smem_for(kchunk) > 44 * 1024) kchunk = KC;
  size_t sm = smem_for(kchunk);
  JD_CHECK_ARG(sm <= 200 * 1024, "%s: PSF too wide for the direct kernel (kw=%d)", name, kw);
  auto kern = conv_kernel<MODE, TY, TX, R>;
  static bool attr_set = false;  // once per process and instantiation: opt in to the 200 KB limit checked above
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  dim3 grid((fW + C * TX - 1) / (C * TX), (fH + R * TY - 1) / (R * TY));
  kern<<<grid, TY * TX, sm, st>>>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, kchunk);
  JD_CHECK_LAUNCH(name);
  return JD_OK;
}

template <int MODE>
static int launch_conv(const float* in, const float* scale, const float* psf, float* out, int fH, int fW, int kh,
                       int kw, int oy, int ox, int f, int H, int W, int accumulate, cudaStream_t st,
                       const char* name) {
  static int tile = -1;
  if (tile < 0) {
    const char* e = getenv("JD_CONV_TILE");  // tuning knob: rows per thread / CTA shape, see below
    tile = e ? atoi(e) : 1;
  }
  if (tile == 1)  // default: 2 rows per thread, 16x16 threads -> 32x64 tiles (best of the sweep, tools/conv_exp.py)
    return launch_conv_t<MODE, 16, 16, 2>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
  if (tile == 2)  // 2 rows per thread, 8x16 threads -> 16x64 tiles
    return launch_conv_t<MODE, 8, 16, 2>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
  if (tile == 3)  // 1 row per thread, 16x16 threads -> 16x64 tiles
    return launch_conv_t<MODE, 16, 16, 1>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
  return launch_conv_t<MODE, 8, 16, 4>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
}

}  // namespace jd

using namespace jd;

extern "C" {

int jd_conv_forward_direct(const float* flux, const float* exposure, const float* psf, float* conv, int fH,
                           int fW, int kh, int kw, jd_stream_t stream) {
  JD_CHECK_ARG(flux && psf && conv, "jd_conv_forward_direct: null pointer");
  JD_CHECK_ARG(fH > 0 && fW > 0 && kh > 0 && kw > 0, "jd_conv_forward_direct: bad shape");
  int sy = (kh - 1) / 2, sx = (kw - 1) / 2;
  return launch_conv<CONV_FWD>(flux, exposure, psf, conv, fH, fW, kh, kw, sy - (kh - 1), sx - (kw - 1), 1, fH, fW,
                               0, to_stream(stream), "jd_conv_forward_direct");
}

int jd_conv_backward_direct(const float* dpool, const float* exposure, const float* psf, float* dflux,
                            int accumulate, int fH, int fW, int kh, int kw, int f, int H, int W,
                            jd_stream_t stream) {
  JD_CHECK_ARG(dpool && psf && dflux, "jd_conv_backward_direct: null pointer");
  JD_CHECK_ARG(fH > 0 && fW > 0 && kh > 0 && kw > 0 && f >= 1 && H * f <= fH && W * f <= fW,
               "jd_conv_backward_direct: bad shape");
  int sy = (kh - 1) / 2, sx = (kw - 1) / 2;
  return launch_conv<CONV_BWD>(dpool, exposure, psf, dflux, fH, fW, kh, kw, -sy, -sx, f, H, W, accumulate,
                               to_stream(stream), "jd_conv_backward_direct");
}

}  // extern "C"
