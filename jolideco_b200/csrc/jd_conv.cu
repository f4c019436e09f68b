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
  static bool attr_set[64] = {};  // once per process and instantiation: opt in to the 200 KB limit checked above
  if (first_use_on_device(attr_set)) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 grid((fW + C * TX - 1) / (C * TX), (fH + R * TY - 1) / (R * TY));
  kern<<<grid, TY * TX, sm, st>>>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, kchunk);
  JD_CHECK_LAUNCH(name);
  return JD_OK;
}


// ------------------------------------------------------------------------------------------------
// v3: whole PSF in shared memory, asynchronous (cp.async, zero-filling) tile staging, 4x4 register
// blocks, exact tap count (no zero-padded kernel rows / columns in the FMA loop) and an optional
// split of the kernel rows over S thread groups of one CTA (small images: more warps per SM).
//
//   block = S groups x (TY x TX) threads; every group owns the whole (4 TY) x (4 TX) output tile for the
//   kernel rows [q kh/S, (q+1) kh/S); the partial tiles are summed through shared memory.
//   Staging: 16-byte cp.async when the tile origin is 16-byte aligned in global memory (the kernel is
//   shifted right by dx = (ox mod 4) zero taps to get there), else 4-byte cp.async; out-of-image
//   chunks are zero-filled by the copy itself (src-size 0), so the loop carries no predicates.
//   Forward: flux and exposure tiles are both copied and multiplied in place by the copying thread.
//   FMA loop: per staged input row t one sliding 4+4 window (LDS.128) feeds the output rows r with
//   kernel row a = t - r (4 broadcast LDS.128): 64 FFMA per 5 LDS.  The kernel rows sit at a
//   compile-time stride (KS) so all four are immediate offsets of one pointer; the first / last three
//   input rows of a group (fewer than 4 output rows take part) are separate instantiations with a
//   compile-time row range instead of predicates.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

constexpr int R3 = 4;   // output rows per thread (v3)
constexpr int TY3 = 8;  // thread rows per group (v3): 32-row tiles
constexpr int KS = 68;  // shared-memory stride of a kernel row (floats): kw + dx <= 68

// one staged input row against the kernel rows a = t - r of the output rows r in [RLO, RHI]
// (krow0 = kernel row t; row t - r sits r * KS floats below).  CG column groups of 4 outputs, GS floats apart:
// every kernel vector feeds CG x 16 FFMA.
template <int KT, int RLO, int RHI, int CG, int GS>
__device__ __forceinline__ void conv3_row(float (&acc)[R3][CG][C], const float* __restrict__ ip,
                                          const float* __restrict__ kp, int ng_full) {
  float4 lo[CG];
#pragma unroll
  for (int cg = 0; cg < CG; ++cg) lo[cg] = *reinterpret_cast<const float4*>(ip + cg * GS);
#pragma unroll 2
  for (int g = 0; g < ng_full; ++g) {
    ip += 4;
    float win[CG][8];
#pragma unroll
    for (int cg = 0; cg < CG; ++cg) {
      const float4 hi = *reinterpret_cast<const float4*>(ip + cg * GS);
      win[cg][0] = lo[cg].x, win[cg][1] = lo[cg].y, win[cg][2] = lo[cg].z, win[cg][3] = lo[cg].w;
      win[cg][4] = hi.x, win[cg][5] = hi.y, win[cg][6] = hi.z, win[cg][7] = hi.w;
      lo[cg] = hi;
    }
#pragma unroll
    for (int r = RLO; r <= RHI; ++r) {
      const float4 kv = *reinterpret_cast<const float4*>(kp - r * KS);
      const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int cg = 0; cg < CG; ++cg)
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
#pragma unroll
          for (int c = 0; c < C; ++c) acc[r][cg][c] = fmaf(kk[bb], win[cg][bb + c], acc[r][cg][c]);
    }
    kp += 4;
  }
  if (KT > 0) {  // last tap group holds KT (< 4) taps
    float win[CG][8];
#pragma unroll
    for (int cg = 0; cg < CG; ++cg) {
      float4 hi = make_float4(0.f, 0.f, 0.f, 0.f);
      if (KT > 1) hi = *reinterpret_cast<const float4*>(ip + 4 + cg * GS);
      win[cg][0] = lo[cg].x, win[cg][1] = lo[cg].y, win[cg][2] = lo[cg].z, win[cg][3] = lo[cg].w;
      win[cg][4] = hi.x, win[cg][5] = hi.y, win[cg][6] = hi.z, win[cg][7] = hi.w;
    }
#pragma unroll
    for (int r = RLO; r <= RHI; ++r) {
      const float4 kv = *reinterpret_cast<const float4*>(kp - r * KS);
      const float kk[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int cg = 0; cg < CG; ++cg)
#pragma unroll
        for (int bb = 0; bb < KT; ++bb)
#pragma unroll
          for (int c = 0; c < C; ++c) acc[r][cg][c] = fmaf(kk[bb], win[cg][bb + c], acc[r][cg][c]);
    }
  }
}

template <int MODE, int TY, int TX, int CG, int KT>
__global__ void __launch_bounds__(TY * TX * 4)
conv3_kernel(const float* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ psf,
             float* __restrict__ out, int fH, int fW, int kh, int kw, int oy, int ox, int f, int H, int W,
             int accumulate, int dx, int vec, int S) {
  constexpr int GS = C * TX;                                   // column-group stride (floats)
  constexpr int TH = R3 * TY, TW = GS * CG, NG = TY * TX;      // NG threads per group
  extern __shared__ __align__(16) float smem[];
  const int kwp = (kw + dx + 3) & ~3;  // kernel row: dx leading zero taps, padded to a multiple of 4
  const int ng = kwp >> 2, ng_full = KT ? ng - 1 : ng;
  const int iw = TW + kwp, ih = TH + kh - 1;
  const bool has_e = MODE == CONV_FWD && scale != nullptr;
  float* s_k = smem;                 // kh x KS
  float* s_in = smem + kh * KS;      // ih x iw
  float* s_e = s_in + ih * iw;       // ih x iw (forward with exposure), also the reduction buffer

  const int tid = threadIdx.x, nt = NG * S;
  const int q = tid / NG, tg = tid - q * NG;
  const int tx = tg % TX, ty = tg / TX;
  const int tile_y = blockIdx.y * TH, tile_x = blockIdx.x * TW;

  // kernel (forward: flipped PSF), shifted right by dx, zero padded up to kwp
  for (int i = tid; i < kh * KS; i += nt) {
    const int a = i / KS, c = i - a * KS, b = c - dx;
    if (c < kwp) {
      float val = 0.f;
      if (b >= 0 && b < kw) val = MODE == CONV_FWD ? __ldg(psf + (kh - 1 - a) * kw + (kw - 1 - b)) : __ldg(psf + a * kw + b);
      s_k[i] = val;
    }
  }

  // input tile rows [y0, y0 + ih), cols [x0, x0 + iw): flattened chunk index -> (row, chunk) by multiply-high
  const int y0 = tile_y + oy, x0 = tile_x + ox - dx;
  if (vec) {
    const int ncx = iw >> 2, total = ih * ncx;
    const uint32_t inv = 0xFFFFFFFFu / (uint32_t)ncx + 1u;  // exact floor(c / ncx) for c * ncx < 2^32
    for (int c = tid; c < total; c += nt) {
      const int ry = (int)__umulhi((uint32_t)c, inv), cx = c - ry * ncx;
      const int y = y0 + ry, x = x0 + 4 * cx;
      const bool ok = y >= 0 && y < fH && x >= 0 && x + 3 < fW;
      const int64_t off = ok ? (int64_t)y * fW + x : 0;
      cp_async16(smem_addr(s_in + ry * iw + 4 * cx), in + off, ok ? 16 : 0);
      if (has_e) cp_async16(smem_addr(s_e + ry * iw + 4 * cx), scale + off, ok ? 16 : 0);
    }
    cp_async_wait_all();
    if (has_e) {  // g = flux * exposure, in place, by the thread that copied the chunk
      for (int c = tid; c < total; c += nt) {
        float4* pa = reinterpret_cast<float4*>(s_in) + c;
        const float4 a = *pa, e = reinterpret_cast<const float4*>(s_e)[c];
        *pa = make_float4(a.x * e.x, a.y * e.y, a.z * e.z, a.w * e.w);
      }
    }
  } else {
    const int total = ih * iw;
    const uint32_t inv = 0xFFFFFFFFu / (uint32_t)iw + 1u;
    for (int c = tid; c < total; c += nt) {
      const int ry = (int)__umulhi((uint32_t)c, inv), rx = c - ry * iw;
      const int y = y0 + ry, x = x0 + rx;
      bool ok = y >= 0 && y < fH && x >= 0 && x < fW;
      int64_t off = 0;
      if (MODE == CONV_FWD) {
        off = ok ? (int64_t)y * fW + x : 0;
      } else if (ok) {
        const int py = y / f, px = x / f;
        ok = py < H && px < W;
        off = ok ? (int64_t)py * W + px : 0;
      }
      cp_async4(smem_addr(s_in + c), in + off, ok ? 4 : 0);
      if (has_e) cp_async4(smem_addr(s_e + c), scale + off, ok ? 4 : 0);
    }
    cp_async_wait_all();
    if (has_e) {
      for (int c = tid; c < total; c += nt) s_in[c] *= s_e[c];
    }
  }
  __syncthreads();

  float acc[R3][CG][C];
#pragma unroll
  for (int r = 0; r < R3; ++r)
#pragma unroll
    for (int cg = 0; cg < CG; ++cg)
#pragma unroll
      for (int c = 0; c < C; ++c) acc[r][cg][c] = 0.f;

  // this group's kernel rows [a_lo, a_hi): input rows t = a + r, t in [a_lo, a_hi + R3 - 1)
  const int a_lo = (q * kh) / S, a_hi = ((q + 1) * kh) / S;
  const float* ip = s_in + (ty * R3 + a_lo) * iw + tx * C;  // staged row t, this thread's first column
  const float* kp = s_k + a_lo * KS;                        // kernel row t
  if (a_hi - a_lo >= R3 - 1) {
    conv3_row<KT, 0, 0, CG, GS>(acc, ip, kp, ng_full);
    conv3_row<KT, 0, 1, CG, GS>(acc, ip + iw, kp + KS, ng_full);
    conv3_row<KT, 0, 2, CG, GS>(acc, ip + 2 * iw, kp + 2 * KS, ng_full);
    ip += 3 * iw, kp += 3 * KS;
    for (int t = a_lo + 3; t < a_hi; ++t, ip += iw, kp += KS) conv3_row<KT, 0, 3, CG, GS>(acc, ip, kp, ng_full);
    conv3_row<KT, 1, 3, CG, GS>(acc, ip, kp, ng_full);
    conv3_row<KT, 2, 3, CG, GS>(acc, ip + iw, kp + KS, ng_full);
    conv3_row<KT, 3, 3, CG, GS>(acc, ip + 2 * iw, kp + 2 * KS, ng_full);
  } else {
    // fewer than 3 kernel rows in this group: one output row at a time, row by row
    for (int a = a_lo; a < a_hi; ++a) {
      conv3_row<KT, 0, 0, CG, GS>(acc, s_in + (ty * R3 + a) * iw + tx * C, s_k + a * KS, ng_full);
      conv3_row<KT, 1, 1, CG, GS>(acc, s_in + (ty * R3 + a + 1) * iw + tx * C, s_k + (a + 1) * KS, ng_full);
      conv3_row<KT, 2, 2, CG, GS>(acc, s_in + (ty * R3 + a + 2) * iw + tx * C, s_k + (a + 2) * KS, ng_full);
      conv3_row<KT, 3, 3, CG, GS>(acc, s_in + (ty * R3 + a + 3) * iw + tx * C, s_k + (a + 3) * KS, ng_full);
    }
  }

  if (S > 1) {  // sum the groups' partial tiles (accumulator-major: conflict-free)
    constexpr int NA = R3 * CG * C;
    __syncthreads();
    float* s_red = s_in;  // (S - 1) x NA x NG floats; the host sizes s_in + s_e for it
    if (q > 0) {
#pragma unroll
      for (int r = 0; r < R3; ++r)
#pragma unroll
        for (int cg = 0; cg < CG; ++cg)
#pragma unroll
          for (int c = 0; c < C; ++c) s_red[((q - 1) * NA + (r * CG + cg) * C + c) * NG + tg] = acc[r][cg][c];
    }
    __syncthreads();
    if (q > 0) return;
    for (int qq = 0; qq < S - 1; ++qq)
#pragma unroll
      for (int r = 0; r < R3; ++r)
#pragma unroll
        for (int cg = 0; cg < CG; ++cg)
#pragma unroll
          for (int c = 0; c < C; ++c) acc[r][cg][c] += s_red[(qq * NA + (r * CG + cg) * C + c) * NG + tg];
  }

  const bool vec_out = (fW & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                       (scale == nullptr || (reinterpret_cast<uintptr_t>(scale) & 15) == 0);
#pragma unroll
  for (int r = 0; r < R3; ++r) {
    const int y = tile_y + ty * R3 + r;
    if (y >= fH) continue;
#pragma unroll
    for (int cg = 0; cg < CG; ++cg) {
      const int x = tile_x + cg * GS + tx * C;
      const int64_t o = (int64_t)y * fW + x;
      if (x + C <= fW && vec_out) {
        float4 val = make_float4(acc[r][cg][0], acc[r][cg][1], acc[r][cg][2], acc[r][cg][3]);
        if (MODE == CONV_BWD) {
          if (scale) {
            const float4 sc = *reinterpret_cast<const float4*>(scale + o);
            val.x *= sc.x, val.y *= sc.y, val.z *= sc.z, val.w *= sc.w;
          }
          if (accumulate) {
            const float4 old = *reinterpret_cast<const float4*>(out + o);
            val.x += old.x, val.y += old.y, val.z += old.z, val.w += old.w;
          }
        }
        *reinterpret_cast<float4*>(out + o) = val;
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (x + c >= fW) continue;
          float val = acc[r][cg][c];
          if (MODE == CONV_BWD) {
            if (scale) val *= scale[o + c];
            if (accumulate) val += out[o + c];
          }
          out[o + c] = val;
        }
      }
    }
  }
}

constexpr size_t CONV3_SMEM_MAX = 200 * 1024;

struct Conv3Plan {
  int tx, cg, S, dx, vec, kt;
  size_t smem;
};

static size_t conv3_smem(int mode, bool has_scale, int kh, int kw, int dx, int tx, int cg, int S) {
  const int kwp = (kw + dx + 3) & ~3, NG = TY3 * tx;
  size_t buf = (size_t)(R3 * TY3 + kh - 1) * (C * tx * cg + kwp) * ((mode == CONV_FWD && has_scale) ? 2 : 1);
  const size_t red = (size_t)(S - 1) * R3 * C * cg * NG;
  if (buf < red) buf = red;
  return ((size_t)kh * KS + buf) * sizeof(float);
}

// tile shape / kernel-row split: enough warps per SM for the FMA pipe, and a last wave that is not mostly idle
static void conv3_auto(int mode, bool has_scale, int fH, int fW, int kh, int kw, Conv3Plan* p) {
  const int nsm = num_sms();
  constexpr int TH = R3 * TY3;
  const long rows = (fH + TH - 1) / TH;
  const long tiles64 = (long)((fW + 63) / 64) * rows, tiles32 = (long)((fW + 31) / 32) * rows;
  // many waves: 64-column tiles with two column groups per thread (32 outputs per thread, least staging and
  // fewest shared-memory reads per FMA); otherwise 32-column tiles for the balance of the last wave
  p->tx = 8;
  p->cg = tiles64 >= 6L * nsm ? 2 : 1;
  const long tiles = p->cg == 2 ? tiles64 : tiles32;
  const long per_sm = (tiles + nsm - 1) / nsm;
  const int warps_per_group = TY3 * p->tx / 32;
  // PSF-row split S: about 24 resident warps per SM keep the FMA pipe busy (tools/conv_exp.py sweep)
  int S = 1;
  for (;;) {
    const size_t smem = conv3_smem(mode, has_scale, kh, kw, p->dx, p->tx, p->cg, S) + 1024;
    long resident = (long)(220 * 1024 / smem);
    if (resident > per_sm) resident = per_sm;
    if (resident < 1) resident = 1;
    if (S >= 4 || resident * warps_per_group * S >= 24 || kh < 8 * S) break;
    S *= 2;
  }
  p->S = S;
}

static bool conv3_plan(int mode, const float* in, const float* scale, const float* out, int fH, int fW, int kh, int kw,
                       int ox, int f, int H, int W, Conv3Plan* p) {
  const bool aligned = (fW & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                       (scale == nullptr || (reinterpret_cast<uintptr_t>(scale) & 15) == 0) &&
                       (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  p->vec = aligned && (mode == CONV_FWD || (f == 1 && H == fH && W == fW));
  p->dx = p->vec ? ((ox % 4) + 4) % 4 : 0;
  if (kw + p->dx > KS) return false;
  p->kt = (kw + p->dx) & 3;
  p->tx = 8;
  p->cg = 1;
  p->S = 1;
  conv3_auto(mode, scale != nullptr, fH, fW, kh, kw, p);
  p->smem = conv3_smem(mode, scale != nullptr, kh, kw, p->dx, p->tx, p->cg, p->S);
  return p->smem <= CONV3_SMEM_MAX;
}

template <int MODE, int TX, int CG, int KT>
static int launch_conv3_t(const Conv3Plan& p, const float* in, const float* scale, const float* psf, float* out, int fH,
                          int fW, int kh, int kw, int oy, int ox, int f, int H, int W, int accumulate, cudaStream_t st,
                          const char* name) {
  constexpr int TY = TY3;
  auto kern = conv3_kernel<MODE, TY, TX, CG, KT>;
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CONV3_SMEM_MAX);
    if (e != cudaSuccess) {
      set_error("%s: cannot reserve shared memory: %s", name, cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
  }
  dim3 grid((fW + C * TX * CG - 1) / (C * TX * CG), (fH + R3 * TY - 1) / (R3 * TY));
  kern<<<grid, TY * TX * p.S, p.smem, st>>>(in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, p.dx, p.vec,
                                            p.S);
  JD_CHECK_LAUNCH(name);
  return JD_OK;
}

template <int MODE, int TX, int CG>
static int launch_conv3_kt(const Conv3Plan& p, const float* in, const float* scale, const float* psf, float* out, int fH,
                           int fW, int kh, int kw, int oy, int ox, int f, int H, int W, int accumulate, cudaStream_t st,
                           const char* name) {
  switch (p.kt) {
    case 0: return launch_conv3_t<MODE, TX, CG, 0>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
    case 1: return launch_conv3_t<MODE, TX, CG, 1>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
    case 2: return launch_conv3_t<MODE, TX, CG, 2>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
    default: return launch_conv3_t<MODE, TX, CG, 3>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
  }
}

// tuning knobs (environment at first use, or jd_conv_tuning): JD_CONV_TILE = shape of the previous kernel,
// JD_CONV_V3 = 0 selects the previous (synchronously staged, chunked) kernel, JD_CONV_TX = 8 | 16 | 28 forces the v3
// tile (32 | 64 columns with one column group per thread | 64 columns with two), JD_CONV_S = 1 | 2 | 4 forces the v3 kernel-row split
static int g_tile = -1, g_v3 = 1, g_tx = 0, g_s = 0;
static void conv_tuning_init() {
  if (g_tile >= 0) return;
  const char* e = getenv("JD_CONV_TILE");
  g_tile = e ? atoi(e) : 1;
  if ((e = getenv("JD_CONV_V3"))) g_v3 = atoi(e);
  if ((e = getenv("JD_CONV_TX"))) g_tx = atoi(e);
  if ((e = getenv("JD_CONV_S"))) g_s = atoi(e);
}

template <int MODE>
static int launch_conv(const float* in, const float* scale, const float* psf, float* out, int fH, int fW, int kh,
                       int kw, int oy, int ox, int f, int H, int W, int accumulate, cudaStream_t st,
                       const char* name) {
  conv_tuning_init();
  const int tile = g_tile, v3 = g_v3, force_tx = g_tx, force_s = g_s;
  Conv3Plan p;
  if (v3 && conv3_plan(MODE, in, scale, out, fH, fW, kh, kw, ox, f, H, W, &p)) {
    if (force_tx == 8 || force_tx == 16 || force_tx == 28 || force_s == 1 || force_s == 2 || force_s == 4) {
      if (force_tx == 8 || force_tx == 16) p.tx = force_tx, p.cg = 1;
      if (force_tx == 28) p.tx = 8, p.cg = 2;
      if (force_s) p.S = force_s;
      p.smem = conv3_smem(MODE, scale != nullptr, kh, kw, p.dx, p.tx, p.cg, p.S);
    }
    if (p.smem <= CONV3_SMEM_MAX) {
      if (p.cg == 2)
        return launch_conv3_kt<MODE, 8, 2>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
      if (p.tx == 16)
        return launch_conv3_kt<MODE, 16, 1>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
      return launch_conv3_kt<MODE, 8, 1>(p, in, scale, psf, out, fH, fW, kh, kw, oy, ox, f, H, W, accumulate, st, name);
    }
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

int jd_conv_tuning(int v3, int tx, int split) {
  JD_CHECK_ARG((tx == 0 || tx == 8 || tx == 16 || tx == 28) && (split == 0 || split == 1 || split == 2 || split == 4),
               "jd_conv_tuning: tx must be 0, 8, 16 or 28 and split 0, 1, 2 or 4");
  conv_tuning_init();
  g_v3 = v3 ? 1 : 0;
  g_tx = tx;
  g_s = split;
  return JD_OK;
}

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
