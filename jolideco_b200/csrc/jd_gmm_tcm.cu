// GMM patch prior forward on tcgen05, third generation: split-TF32 with the two correction products on the FP16
// pipe, persistent stream-K work decomposition, and the patch gather of the NEXT tile overlapped with the tensor
// work of the current one (double-buffered A operand in TMEM, dedicated gather warps).
//
// Precision recipe.  With power-of-two scales (per patch row for x, per component for L; max |.| in [2^13, 2^14))
//     x' = xt + xr,  xt = tf32(x'),  |xr| <= 2^-11 |x'|          L' = Lt + Lr likewise
//     x' L' = xt Lt            kind::tf32, 8 MMAs (K = 8)      exact products, FP32 accumulation in TMEM
//           + xr L'            kind::f16,  4 MMAs (K = 16)     operands half(xr), half(L'):  error 2^-11 * 2^-11
//           + x' Lr            kind::f16,  4 MMAs (K = 16)     operands half(x'), half(Lr):  error 2^-11 * 2^-11
//           + xr Lr            dropped: 2^-22
// i.e. the same 2^-21..2^-22 relative accuracy as the 3 x TF32 split of jd_gmm_tc.cu (the three products there are
// lo.hi + hi.lo + hi.hi in TF32), with the two small products at twice the tensor rate and half the operand bytes:
// 1 + 1/2 + 1/2 = 2 TF32-product equivalents instead of 3 (x 0.625 with the triangular trim).  All three accumulate
// into the same FP32 TMEM columns.  The epilogue undoes the scales with one multiply.
//
// Why it matters: with 3 x TF32 the kernel is tensor-pipe bound (88 % busy, ~730 clk per component and CTA); the
// next floor is the epilogue's TMEM read, 128 x 64 FP32 = 32 KB per component at 64 B/clk = 512 clk.  This kernel
// issues ~460 clk of tensor work per component and therefore sits on the TMEM-read floor.
//
// Work decomposition: the (tile pair, component) space is linearised and cut into equal chunks, one per CTA pair
// (2-CTA cluster, B multicast), as in jd_gmm_prior_forward_tc_sk; a chunk is a sequence of segments (component
// range of one tile pair).  Unlike that kernel a segment boundary does not drain the pipeline: warps 4-7 gather the
// patches of segment s+1 into the other A buffer while the MMAs of segment s run; the MMA warps move on as soon as
// `afull` of the next buffer has fired.  Accumulator slots, the B ring and every mbarrier phase run on the CTA's
// global position counter.
//
// Warps (512 threads, one CTA per SM): 0-1 bulk-TMA producers | 2-3 MMA issuers | 4-7 gather | 8-11 epilogue
// group A (even positions) | 12-15 epilogue group B (odd positions).
// TMEM (512 columns): [0,128) A buffer 0 | [128,256) A buffer 1 | [256,512) four 64-column accumulator slots.
// A buffer: [0,64) tf32(x') | [64,96) half2(xr) | [96,128) half2(x').
// B image per component (32 KB, 128B-swizzled K-major): [0,16K) tf32(L') | [16K,24K) half(L') | [24K,32K) half(Lr).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_tc_ptx.cuh"

namespace jd {
namespace tcm {

using namespace tcx;

constexpr int TM = 128;
constexpr int NSTAGE = 6;
constexpr int NSLOT = 4;
constexpr int SLOT_COLS = 64;
constexpr int A_COLS = 128;
constexpr int ACC0 = 2 * A_COLS;
constexpr int TMEM_COLS = 512;
constexpr int KBLOCK_BYTES = 64 * 128;           // 8 KB: 64 rows x 128 B
constexpr int B_TF32_BYTES = 2 * KBLOCK_BYTES;   // 64 x 64 tf32
constexpr int B_F16_BYTES = KBLOCK_BYTES;        // 64 x 64 half
constexpr int B_BYTES = B_TF32_BYTES + 2 * B_F16_BYTES;
constexpr int CLUSTER = 2;
constexpr int NPROD = 2;
constexpr int NMMA = 2;
static_assert(NSTAGE % NPROD == 0 && NSLOT % NMMA == 0 && NSLOT % 2 == 0, "fixed barrier ownership (jd_gmm_tc.cu)");
constexpr int M0 = NPROD;      // first MMA warp (owns the TMEM allocation)
constexpr int G0 = 4;          // first gather warp
constexpr int E0 = 8;          // first epilogue warp
constexpr int NTHREADS = 512;
constexpr int MW_BYTES = 64 * 4;
constexpr int NBAR = 2 * NSTAGE + 3 * NSLOT + 4;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + NSTAGE * B_BYTES + NSLOT * MW_BYTES + 6144 /*barriers, flags*/;

__device__ __host__ constexpr uint32_t idesc_tf32(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __host__ constexpr uint32_t idesc_f16(uint32_t n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------- setup: pack Lw_k^T, scaled, as tf32 + two halves
// One CTA per component.  Row n of every image = whitened feature j, K index = input feature i.
__global__ void pack_bm_kernel(const float* __restrict__ Lw, int K, uint8_t* __restrict__ out,
                               float* __restrict__ binv) {
  __shared__ float s_max[32];
  const int k = blockIdx.x;
  const float* L = Lw + (size_t)k * 4096;
  float m = 0.f;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) m = fmaxf(m, fabsf(L[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_max[w]);
  int e = m > 0.f ? ilogbf(m) : 13;
  e = max(-100, min(100, e));
  const float sB = ldexpf(1.f, 13 - e);
  if (threadIdx.x == 0) binv[k] = ldexpf(1.f, e - 13);
  uint8_t* base = out + (size_t)k * B_BYTES;
  for (int idx = threadIdx.x; idx < 4096; idx += blockDim.x) {
    const int i = idx >> 6, j = idx & 63;  // Lw[k][i][j]
    const float v = L[idx] * sB;
    const float t = tf32_rna(v);
    const float r = v - t;
    {  // tf32 image: two 128B-swizzled k-blocks of 32 input features
      const int kb = i >> 5, c = (i & 31) >> 2, el = i & 3;
      const uint32_t off = kb * KBLOCK_BYTES + (j >> 3) * 1024 + (j & 7) * 128 + ((c ^ (j & 7)) << 4) + el * 4;
      *reinterpret_cast<float*>(base + off) = t;
    }
    const uint32_t off16 = (j >> 3) * 1024 + (j & 7) * 128 + (((i >> 3) ^ (j & 7)) << 4) + (i & 7) * 2;
    *reinterpret_cast<__half*>(base + B_TF32_BYTES + off16) = __float2half_rn(v);
    *reinterpret_cast<__half*>(base + B_TF32_BYTES + B_F16_BYTES + off16) = __float2half_rn(r);
  }
}

__device__ __forceinline__ int seg_rotation(int cl, int len, unsigned mul) {
  return (int)(((unsigned)cl * mul) % (unsigned)len);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  return (uint32_t)__half_as_ushort(__float2half_rn(a)) | ((uint32_t)__half_as_ushort(__float2half_rn(b)) << 16);
}

// ---------------------------------------------------------------- the forward kernel
template <bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tcm_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                   const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck,
                   const float* __restrict__ binv, int K, int marginalize, int chunk, int smax, unsigned rot_mul,
                   unsigned* __restrict__ counters, float* __restrict__ ws_m, float* __restrict__ ws_s,
                   int* __restrict__ ws_k, float* __restrict__ value, int32_t* __restrict__ argmax,
                   float* __restrict__ logp, double* __restrict__ sum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                            // NSTAGE x 32 KB
  float* sMW = reinterpret_cast<float*>(sB + NSTAGE * B_BYTES);  // NSLOT x 64 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES + NSLOT * MW_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + NBAR);
  int* s_flag = reinterpret_cast<int*>(s_tmem + 2);
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);         // 2 x 128 ints
  float* s_rinv = reinterpret_cast<float*>(s_valid + 2 * TM);  // 2 x 128: 1 / (row scale)
  float* s_mm = s_rinv + 2 * TM;                              // merge buffers of epilogue group B: max,
  float* s_ms = s_mm + TM;                                    //   sum-exp,
  int* s_mk = reinterpret_cast<int*>(s_ms + TM);              //   argmax

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
  // profiling knobs (JD_TC_DEBUG, results are wrong with any of them): 1 = no epilogue TMEM loads, 2 = no FP16
  // correction products, 4 = no TF32 product, 8 = one MMA per component
  const int dbg = marginalize >> 8;
  marginalize &= 1;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NSLOT + s); };
  auto mwfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + s); };
  auto afull_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 3 * NSLOT + b); };
  auto aempty_bar = [&](int b) { return bar0 + 8u * (2 * NSTAGE + 3 * NSLOT + 2 + b); };
  const uint32_t crank = cluster_ctarank();

  // this CTA pair's chunk of the linearised (tile pair, component) space
  const int cl = blockIdx.x / CLUSTER;
  const int n_tiles = (g.P + TM - 1) / TM, n_pairs = (n_tiles + CLUSTER - 1) / CLUSTER;
  const long long w_tot = (long long)n_pairs * K;
  const long long lin_begin = (long long)cl * chunk;
  const long long lin_end = lin_begin + chunk < w_tot ? lin_begin + chunk : w_tot;
  const int tp_first = (int)(lin_begin / K), tp_last = (int)((lin_end - 1) / K);

  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CLUSTER);  // released by the MMA commits of both CTAs of the pair
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
      mbar_init(mwfull_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(afull_bar(b), 4);             // one arrive per gather warp
      mbar_init(aempty_bar(b), NMMA + 1);     // last MMAs of the segment (both issuers) + the epilogue has read the flags
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M0) tmem_alloc(smem_u32(s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int q = warp & 3;
  const int row = q * 32 + lane;

  // segment geometry of tile pair tp: positions [pos_lo, pos_hi) of this CTA's chunk, first component ka
#define JD_SEGMENT(tp)                                                                              \
  const long long comp0 = (long long)(tp)*K;                                                         \
  const int pos_lo = (int)((lin_begin > comp0 ? lin_begin : comp0) - lin_begin);                     \
  const int pos_hi = (int)((lin_end < comp0 + K ? lin_end : comp0 + K) - lin_begin);                 \
  const int ka = (int)(lin_begin + pos_lo - comp0), len = pos_hi - pos_lo;                           \
  const int rot = seg_rotation(cl, len, rot_mul);                                                    \
  const int sidx = (tp)-tp_first, buf = sidx & 1, use = sidx >> 1;                                   \
  (void)ka; (void)rot; (void)buf; (void)use;

  if (warp < NPROD) {
    // ===================== bulk-TMA producers ======================================================
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      int pos = pos_lo + ((pos_lo % NPROD) == warp ? 0 : (warp - (pos_lo % NPROD) + NPROD) % NPROD);
      for (; pos < pos_hi; pos += NPROD) {
        int idx = pos - pos_lo + rot;
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int s = pos % NSTAGE, t = pos % NSLOT;
        mbar_wait(empty_bar(s), ((pos / NSTAGE) & 1) ^ 1);
        if (!ZERO_MEAN) mbar_wait(tempty_bar(t), ((pos / NSLOT) & 1) ^ 1);
        if (elect_one()) {
          // this CTA fetches half `crank` of the image (tf32 part | the two half parts) for both CTAs of the pair.
          // Upper-triangular factors: whitened features 0..31 do not depend on input features 32..63, i.e. rows 0..31
          // of the second tf32 k-block (4 KB) are zeros no MMA reads (k-steps 4..7 start at row 32): not copied.
          mbar_arrive_expect_tx(full_bar(s), (TRI && !(dbg & 32)) ? B_BYTES - KBLOCK_BYTES / 2 : B_BYTES);
          const uint32_t dst = smem_u32(sB + s * B_BYTES) + crank * (B_BYTES / CLUSTER);
          const uint8_t* src = Bt + (size_t)kc * B_BYTES + crank * (B_BYTES / CLUSTER);
          const uint16_t both = (uint16_t)((1u << CLUSTER) - 1);
          if (TRI && crank == 0 && !(dbg & 32)) {
            bulk_g2s_mc(dst, src, KBLOCK_BYTES, full_bar(s), both);
            bulk_g2s_mc(dst + KBLOCK_BYTES + KBLOCK_BYTES / 2, src + KBLOCK_BYTES + KBLOCK_BYTES / 2, KBLOCK_BYTES / 2,
                        full_bar(s), both);
          } else {
            bulk_g2s_mc(dst, src, B_BYTES / CLUSTER, full_bar(s), both);
          }
          if (!ZERO_MEAN) {
            mbar_arrive_expect_tx(mwfull_bar(t), MW_BYTES);
            bulk_g2s(smem_u32(sMW + t * 64), mw + (size_t)kc * 64, MW_BYTES, mwfull_bar(t));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < M0 + NMMA) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =========
    const int w = warp - M0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      mbar_wait(afull_bar(buf), use & 1);  // the gather warps have stored this segment's A operand
      tc_fence_after();
      const uint32_t a_base = tmem_u + buf * A_COLS;
      int pos = pos_lo + ((pos_lo % NMMA) == w ? 0 : (w - (pos_lo % NMMA) + NMMA) % NMMA);
      bool released = false;
      for (; pos < pos_hi; pos += NMMA) {
        const int s = pos % NSTAGE, t = pos % NSLOT;
        const bool last = pos + NMMA >= pos_hi;
        mbar_wait(tempty_bar(t), ((pos / NSLOT) & 1) ^ 1);
        mbar_wait(full_bar(s), (pos / NSTAGE) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_t = sB_lo0 + s * (B_BYTES >> 4);
          const uint32_t b_h = b_t + (B_TF32_BYTES >> 4), b_r = b_h + (B_F16_BYTES >> 4);
          const uint32_t d = tmem_u + ACC0 + t * SLOT_COLS;
          uint32_t acc = 0;
          // small terms first: xr . half(L'), half(x') . half(Lr), then tf32(x') . tf32(L')
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const uint32_t a_col = pass == 0 ? 64u : 96u;
            const uint32_t b_base = pass == 0 ? b_h : b_r;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if ((dbg & 2) && (kk > 0 || pass > 0 || !(dbg & 4))) continue;
              if ((dbg & 8) && (kk > 0 || pass > 0)) continue;
              // upper-triangular Lw: input features [16kk, 16kk+16) only reach whitened features >= 16kk
              const uint32_t n0 = TRI ? 16u * kk : 0u;
              const uint32_t off16 = (kk * 32 + n0 * 128) >> 4;
              umma_f16_ts(d + n0, a_base + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_f16(64 - n0), acc);
              acc = 1;
            }
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if ((dbg & 12) != 0) continue;
            const uint32_t n0 = TRI ? 16u * (kk >> 1) : 0u;
            const uint32_t off16 = ((kk >> 2) * KBLOCK_BYTES + (kk & 3) * 32 + n0 * 128) >> 4;
            umma_tf32_ts(d + n0, a_base + kk * 8, desc_from_lo(b_t + off16), idesc_tf32(64 - n0), acc);
            acc = 1;
          }
          umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));  // stage free in both CTAs of the pair
          umma_commit(tfull_bar(t));                                       // accumulator slot complete
          if (last) umma_commit(aempty_bar(buf));  // this warp's last read of the A buffer (same thread as its MMAs)
        }
        released = released || last;
        __syncwarp();
      }
      if (!released) {  // no position of this segment fell to this warp
        if (elect_one()) mbar_arrive(aempty_bar(buf));
        __syncwarp();
      }
    }
  } else if (warp < E0) {
    // ===================== gather: thread = patch row; 64 loads, mean, scale, split, tcgen05.st ==============
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      const int tile = tp * CLUSTER + (int)crank;
      const int64_t p = (int64_t)tile * TM + row;
      float vals[64];
      float sm = 0.f;
      bool ok = p < g.P;
      if (ok) {
        int iy = (int)(p / g.nx) + g.row_begin, ix = (int)(p % g.nx);
        int cols[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) cols[v] = src_col(g, ix, v);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float* src = flux + (int64_t)src_row(g, iy, u) * g.fW;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float x = __ldg(src + cols[v]);
            vals[u * 8 + v] = x;
            sm += x;
            ok = ok && (x > -1e5f);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) vals[i] = 0.f;
      }
      const float mean = sm * (1.f / 64.f);
      float amax = 0.f;
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        vals[i] = ok ? vals[i] - mean : 0.f;
        amax = fmaxf(amax, fabsf(vals[i]));
      }
      int e = amax > 0.f ? ilogbf(amax) : 13;
      e = max(-100, min(100, e));
      const float sA = ldexpf(1.f, 13 - e);
      // the A buffer (and its flags) are free once the MMAs of segment sidx - 2 have completed and its epilogue is done
      mbar_wait(aempty_bar(buf), (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + buf * A_COLS;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float xt[32], xr[16], xh[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a0 = vals[h * 32 + 2 * i] * sA, a1 = vals[h * 32 + 2 * i + 1] * sA;
          const float t0 = tf32_rna(a0), t1 = tf32_rna(a1);
          xt[2 * i] = t0, xt[2 * i + 1] = t1;
          xr[i] = __uint_as_float(pack_half2(a0 - t0, a1 - t1));
          xh[i] = __uint_as_float(pack_half2(a0, a1));
        }
        tmem_st32(a_lane + h * 32, xt);
        // 16 packed words each: features [32h, 32h+32) land in columns [64 + 16h, +16) and [96 + 16h, +16)
        float pk[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = xr[i], pk[16 + i] = xh[i];
        // two 16-column stores through one 32-wide helper would overlap: store xr and xh halves separately below
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a_lane + 64 + h * 16),
            "f"(pk[0]), "f"(pk[1]), "f"(pk[2]), "f"(pk[3]), "f"(pk[4]), "f"(pk[5]), "f"(pk[6]), "f"(pk[7]), "f"(pk[8]),
            "f"(pk[9]), "f"(pk[10]), "f"(pk[11]), "f"(pk[12]), "f"(pk[13]), "f"(pk[14]), "f"(pk[15])
            : "memory");
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
            "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a_lane + 96 + h * 16),
            "f"(pk[16]), "f"(pk[17]), "f"(pk[18]), "f"(pk[19]), "f"(pk[20]), "f"(pk[21]), "f"(pk[22]), "f"(pk[23]),
            "f"(pk[24]), "f"(pk[25]), "f"(pk[26]), "f"(pk[27]), "f"(pk[28]), "f"(pk[29]), "f"(pk[30]), "f"(pk[31])
            : "memory");
      }
      tmem_st_wait();
      s_valid[buf * TM + row] = ok ? 1 : 0;
      s_rinv[buf * TM + row] = ldexpf(1.f, e - 13);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull_bar(buf));
    }
  } else {
    // ===================== epilogue: group A (warps 8-11) even positions, group B (12-15) odd ========
    const int grp = warp >= E0 + 4 ? 1 : 0;
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      const int tile = tp * CLUSTER + (int)crank;
      const int64_t p = (int64_t)tile * TM + row;
      float run_m = -CUDART_INF_F, run_s = 0.f;
      int run_k = 0x7fffffff;
      const int pos_first = pos_lo + ((pos_lo ^ grp) & 1);
      float row_inv = 0.f;
      for (int pos = pos_first; pos < pos_hi; pos += 2) {
        int idx = pos - pos_lo + rot;  // same rotated component order as the producers
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int t = pos % NSLOT;
        const float c_k = __ldg(ck + kc);
        const float b_inv = __ldg(binv + kc);
        if (!ZERO_MEAN) mbar_wait(mwfull_bar(t), (pos / NSLOT) & 1);
        mbar_wait(tfull_bar(t), (pos / NSLOT) & 1);
        tc_fence_after();
        if (pos == pos_first) row_inv = s_rinv[buf * TM + row];  // written by the gather warps before the first MMA
        const float inv = row_inv * b_inv;  // undo the row and component scales
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ACC0 + t * SLOT_COLS;
        float y0[32], y1[32];
        if (dbg & 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) y0[i] = y1[i] = (float)lane;
        } else {
          tmem_ld32(taddr, y0);
          tmem_ld32(taddr + 32, y1);
          tmem_ld_wait();
        }
        float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
        if (ZERO_MEAN) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            qa = fmaf(y0[i], y0[i], qa);
            qb = fmaf(y1[i], y1[i], qb);
            qc = fmaf(y0[i + 1], y0[i + 1], qc);
            qd = fmaf(y1[i + 1], y1[i + 1], qd);
          }
          const float i2 = inv * inv;
          qa *= i2, qb *= i2, qc *= i2, qd *= i2;
        } else {
          const float4* mwk = reinterpret_cast<const float4*>(sMW + t * 64);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 b0 = mwk[c4], b1 = mwk[8 + c4];
            float d0 = fmaf(y0[4 * c4], inv, -b0.x), d1 = fmaf(y0[4 * c4 + 1], inv, -b0.y);
            float d2 = fmaf(y0[4 * c4 + 2], inv, -b0.z), d3 = fmaf(y0[4 * c4 + 3], inv, -b0.w);
            float e0 = fmaf(y1[4 * c4], inv, -b1.x), e1 = fmaf(y1[4 * c4 + 1], inv, -b1.y);
            float e2 = fmaf(y1[4 * c4 + 2], inv, -b1.z), e3 = fmaf(y1[4 * c4 + 3], inv, -b1.w);
            qa = fmaf(d0, d0, qa);
            qb = fmaf(e0, e0, qb);
            qc = fmaf(d1, d1, qc);
            qd = fmaf(e1, e1, qd);
            qa = fmaf(d2, d2, qa);
            qb = fmaf(e2, e2, qb);
            qc = fmaf(d3, d3, qc);
            qd = fmaf(e3, e3, qd);
          }
        }
        const float lp = fmaf(-0.5f, (qa + qb) + (qc + qd), c_k);
        tc_fence_before();  // slot and mw row are free once consumed (lp depends on every load, see mbar_arrive_after)
        __syncwarp();
        if (lane == 0) mbar_arrive_after(tempty_bar(t), lp, rt_zero);
        if (logp && p < g.P) logp[(size_t)kc * g.P + p] = lp;
        if (marginalize) {
          if (lp > run_m) {
            run_s = run_s * expf(run_m - lp) + 1.f;
            run_m = lp;
            run_k = kc;
          } else {
            run_s += expf(lp - run_m);
          }
        } else if (lp > run_m || (lp == run_m && kc < run_k)) {
          run_m = lp;
          run_k = kc;
        }
      }

      // ---- segment end: group B -> group A through shared memory (both groups stay within one segment of each other)
      if (grp == 1) {
        s_mm[row] = run_m;
        s_ms[row] = run_s;
        s_mk[row] = run_k;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        asm volatile("bar.sync 3, 256;" ::: "memory");  // group A has read the buffers
        continue;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const bool ok = s_valid[buf * TM + row] != 0;
      {
        const float om = s_mm[row], os = s_ms[row];
        const int ok_ = s_mk[row];
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (marginalize) {
          const float m = fmaxf(run_m, om);
          run_s = (run_m == -CUDART_INF_F ? 0.f : run_s * expf(run_m - m)) + (om == -CUDART_INF_F ? 0.f : os * expf(om - m));
          run_k = om > run_m ? ok_ : run_k;
          run_m = m;
        } else if (om > run_m || (om == run_m && ok_ < run_k)) {
          run_m = om;
          run_k = ok_;
        }
      }
      // both groups are past their last accumulator of the segment and the flags are in registers: together with the
      // MMA warps' commits this frees the A buffer for the gather of segment sidx + 2
      if (threadIdx.x == E0 * 32) mbar_arrive(aempty_bar(buf));
      bool final_here = len == K;
      if (!final_here) {
        const int c_first = (int)(comp0 / chunk), c_last = (int)((comp0 + K - 1) / chunk);
        const int nseg = c_last - c_first + 1;
        const size_t base = ((size_t)tile * smax) * TM + row;
        const size_t widx = base + (size_t)(cl - c_first) * TM;
        ws_m[widx] = run_m;
        ws_s[widx] = run_s;
        ws_k[widx] = run_k;
        __threadfence();
        asm volatile("bar.sync 4, 128;" ::: "memory");
        if (row == 0) *s_flag = (atomicAdd(&counters[tile], 1u) + 1u == (unsigned)nseg) ? 1 : 0;
        asm volatile("bar.sync 4, 128;" ::: "memory");
        final_here = *s_flag != 0;
        if (final_here) {  // last segment of this tile to arrive: merge the slots in component order
          __threadfence();
          run_m = -CUDART_INF_F, run_s = 0.f, run_k = 0x7fffffff;
          for (int j = 0; j < nseg; ++j) {
            const float om = __ldcg(ws_m + base + (size_t)j * TM), os = __ldcg(ws_s + base + (size_t)j * TM);
            const int ok_ = __ldcg(ws_k + base + (size_t)j * TM);
            if (marginalize) {
              const float m = fmaxf(run_m, om);
              run_s = (run_m == -CUDART_INF_F ? 0.f : run_s * expf(run_m - m)) +
                      (om == -CUDART_INF_F ? 0.f : os * expf(om - m));
              run_k = om > run_m ? ok_ : run_k;
              run_m = m;
            } else if (om > run_m || (om == run_m && ok_ < run_k)) {
              run_m = om;
              run_k = ok_;
            }
          }
          if (row == 0) counters[tile] = 0;  // every segment has arrived: ready for the next launch
        }
        asm volatile("bar.sync 4, 128;" ::: "memory");  // s_flag is read before the next segment rewrites it
      }
      if (final_here) {
        double part = 0.0;
        if (p < g.P) {
          float v = marginalize ? run_m + logf(run_s) : run_m;
          v = ok ? v : 0.f;
          if (value) value[p] = v;
          if (argmax) argmax[p] = ok ? run_k : -1;
          part = (double)v;
        }
        part = warp_sum(part);
        if (lane == 0 && sum) atomicAdd(sum, part);
      }
    }
  }
#undef JD_SEGMENT

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until here
  if (warp == M0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- plan: CTA pairs, chunk, workspace
struct Plan {
  int n_clusters, chunk, smax, n_tiles2;
  size_t off_m, off_s, off_k, bytes;
};
static Plan plan(int64_t P, int K) {
  Plan p;
  const int64_t n_tiles = (P + TM - 1) / TM, n_pairs = (n_tiles + CLUSTER - 1) / CLUSTER;
  const int64_t w_tot = n_pairs * K;
  int C = num_sms() / CLUSTER;  // one CTA per SM, two SMs per pair
  if (const char* e = getenv("JD_TCM_CLUSTERS")) C = atoi(e) > 0 ? atoi(e) : C;
  int64_t chunk = (w_tot + C - 1) / C;
  if (K % 8 == 0) chunk = (chunk + 7) / 8 * 8;  // no segment shorter than 8 components
  if (chunk < 1) chunk = 1;
  p.chunk = (int)chunk;
  p.n_clusters = (int)((w_tot + chunk - 1) / chunk);
  p.smax = (int)((K + chunk - 1) / chunk) + 1;
  p.n_tiles2 = (int)(n_pairs * CLUSTER);
  const size_t cnt = ((size_t)p.n_tiles2 * sizeof(unsigned) + 255) / 256 * 256;
  const size_t part = (size_t)p.n_tiles2 * p.smax * TM * sizeof(float);
  p.off_m = cnt;
  p.off_s = cnt + part;
  p.off_k = cnt + 2 * part;
  p.bytes = cnt + 3 * part;
  return p;
}

}  // namespace tcm
}  // namespace jd

using namespace jd;

extern "C" {

size_t jd_gmm_tcm_packed_bytes(int K) { return (size_t)K * tcm::B_BYTES; }

int jd_gmm_tcm_pack(const float* Lw, int K, void* Bt, float* binv, jd_stream_t stream) {
  JD_CHECK_ARG(Lw && Bt && binv && K > 0, "jd_gmm_tcm_pack: bad arguments");
  tcm::pack_bm_kernel<<<K, 256, 0, to_stream(stream)>>>(Lw, K, reinterpret_cast<uint8_t*>(Bt), binv);
  JD_CHECK_LAUNCH("jd_gmm_tcm_pack");
  return JD_OK;
}

int64_t jd_gmm_tcm_workspace_bytes(int64_t P, int K) {
  if (P <= 0 || K <= 0) return 0;
  return (int64_t)tcm::plan(P, K).bytes;
}

int jd_gmm_prior_forward_tcm(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                             int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                             int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                             int32_t* argmax, float* logp, double* sum, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && binv && mw && ck && workspace && K > 0, "jd_gmm_prior_forward_tcm: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tcm: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tcm: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "jd_gmm_prior_forward_tcm: Bt and mw must be 16-byte aligned, the workspace 256-byte aligned");
  tcx::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaSuccess;
    const void* kerns[4] = {(const void*)tcm::gmm_fwd_tcm_kernel<false, false>, (const void*)tcm::gmm_fwd_tcm_kernel<false, true>,
                            (const void*)tcm::gmm_fwd_tcm_kernel<true, false>, (const void*)tcm::gmm_fwd_tcm_kernel<true, true>};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcm::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tcm: cannot reserve %zu B of shared memory: %s", tcm::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
    attr_set[dev & 63] = true;
  }
  const tcm::Plan p = tcm::plan(g.P, K);
  static int rot_env = -1;
  if (rot_env < 0) {
    const char* e = getenv("JD_TC_SK_ROT");  // 0: visit the components of a segment in ascending order
    rot_env = e ? atoi(e) : 40503;
  }
  auto kern = upper_tri ? (zero_mean ? tcm::gmm_fwd_tcm_kernel<true, true> : tcm::gmm_fwd_tcm_kernel<true, false>)
                        : (zero_mean ? tcm::gmm_fwd_tcm_kernel<false, true> : tcm::gmm_fwd_tcm_kernel<false, false>);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_clusters * tcm::CLUSTER);
  cfg.blockDim = dim3(tcm::NTHREADS);
  cfg.dynamicSmemBytes = tcm::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tcm::CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t* bt8 = reinterpret_cast<const uint8_t*>(Bt);
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  unsigned* counters = reinterpret_cast<unsigned*>(ws);
  float* ws_m = reinterpret_cast<float*>(ws + p.off_m);
  float* ws_s = reinterpret_cast<float*>(ws + p.off_s);
  int* ws_k = reinterpret_cast<int*>(ws + p.off_k);
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("JD_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  if (dbg & 16) kern = zero_mean ? tcm::gmm_fwd_tcm_kernel<false, true> : tcm::gmm_fwd_tcm_kernel<false, false>;  // dense
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, flux, g, shift_yx, bt8, mw, ck, binv, K, (marginalize ? 1 : 0) | (dbg << 8), p.chunk,
                                      p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_forward_tcm: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tcm");
  return JD_OK;
}

}  // extern "C"
