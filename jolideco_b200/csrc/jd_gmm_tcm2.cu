// GMM patch prior forward on tcgen05, fourth generation: TWO patch tiles per CTA against every staged operand image,
// round-robin issuer warps with a stripped issue loop, one epilogue group per tile, in two precision recipes:
//   R = 0  mixed TF32 / FP16 split of jd_gmm_tcm.cu (operand image of jd_gmm_tcm_pack, 32 KB per component)
//   R = 1  split FP16 of jd_gmm_tc16.cu          (operand image of jd_gmm_tc16_pack, 16 KB per component)
// Both carry 22 significand bits per operand (2^-21..2^-22 relative accuracy of every log-probability, FP32
// accumulation in TMEM); R = 1 issues 12 instead of 16 MMAs per tile and component, needs half the TMEM columns for the A
// operand (64 instead of 128 per tile) - which buys a third accumulator slot - and half the operand bytes.
//
// Why (measurements on B200, 1024^2 image, K = 256; profiles/r02_summary.md):
//   * ncu on the one-tile kernel jd_gmm_tcm.cu: tensor pipe 65 % busy, 3.75 GB of operand images cross into the SMs
//     per launch (41 B/clk/SM); switching its epilogue TMEM loads off, or all MMAs but one, barely changes the time;
//   * tools/ubench.cu: TMEM reads sustain > 400 B/clk/SM with 16 warps, an mbarrier hand-over costs ~140 clk, bulk-TMA
//     delivers > 90 B/clk/SM - none of these is the limiter;
//   * clock64 stamps of every hand-over (JD_TCM_TRACE build, tools/tcm_trace.py) on the first two-tile version: ONE
//     elected lane needs ~36 clk per tcgen05.mma (R2UR + descriptor arithmetic + the debug-knob branches around each
//     instruction): 1141 clk to issue the 32 MMAs of a position whose tensor work is 640 clk; and an accumulator slot
//     turns around in issue + ~150 (commit -> epilogue) + ~700 (TMEM loads 460, squares 230) + ~300 (release -> issuer)
//     clk, so with two slots the issuer and the epilogue of a slot simply alternate.
//   * second version (one issuer warp per tile, stripped issue loop: UTCHMMAs back to back): issuing 12 MMAs takes 340
//     clk, but the issuer's loop still needs ~1100 clk per position - every mbarrier wait costs ~100 clk even when the
//     phase is already complete, and the wait that follows the commits stalls another ~300 clk: ~450 clk of fixed
//     overhead per issuer iteration against 240 clk of tensor work.
// Hence: the issue loop is stripped to descriptor adds + UTCHMMA (no run-time knobs, warp index made uniform with a
// shuffle so that operands live in uniform registers); an issuer iteration covers BOTH tiles of a position (24 / 32 MMAs
// per pair of waits and commits), NMMA = NSLOT issuer warps take the positions round-robin so that their fixed
// overheads overlap, and recipe 1 runs three accumulator slots.
//
// Work decomposition (as jd_gmm_tcm.cu): the (tile group, component) space is linearised and cut into equal chunks, one
// per CTA pair (2-CTA cluster, each CTA fetches half of every operand image and multicasts it to both); tile group =
// CLUSTER x TPC = 4 tiles; a chunk is a sequence of segments (component range of one tile group); partial (max, sum-exp,
// argmax) rows of split tiles go through a workspace, the last segment of a tile to arrive merges them in component order.
// The A operand is single-buffered: at a segment boundary the gather warps (tile 0's patch already in registers) store
// once the issuers' last commits of the previous segment have fired.
//
// Warps (512 threads, one CTA per SM): 0 bulk-TMA producer | 1..NMMA MMA issuers | 4-7 gather | 8-11 epilogue of tile
// 0 | 12-15 epilogue of tile 1.
// TMEM (512 columns): [0, TPC A_COLS) A operands | then NSLOT slots of TPC x 64 accumulator columns.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_tc_ptx.cuh"

namespace jd {
namespace tcx2 {

using namespace tcx;

constexpr int TM = 128;
constexpr int TPC = 2;  // patch tiles per CTA and staged operand image
constexpr int ACC_COLS = 64;
constexpr int TMEM_COLS = 512;
constexpr int KBLOCK_BYTES = 64 * 128;  // 8 KB: 64 rows x 128 B (one swizzle atom wide)
constexpr int CLUSTER = 2;
constexpr int M0 = 1;  // first MMA-issuer warp (owns the TMEM allocation); position pos is issued by warp M0 + pos % NMMA
constexpr int G0 = 4;  // first gather warp
constexpr int E0 = 8;  // first epilogue warp
constexpr int NTHREADS = 512;

template <int R>
struct Recipe;
template <>
struct Recipe<0> {  // tf32(x') tf32(L') + half(xr) half(L') + half(x') half(Lr)
  static constexpr int A_COLS = 128;  // [0,64) tf32(x') | [64,96) half2(xr) | [96,128) half2(x')
  static constexpr int NSLOT = 2;   // accumulator slots (both tiles of a position) = issuer warps
  static constexpr int NSTAGE = 6;
  static constexpr int B_BYTES = 4 * KBLOCK_BYTES;  // [0,16K) tf32(L') | [16K,24K) half(L') | [24K,32K) half(Lr)
};
template <>
struct Recipe<1> {  // lo.hi + hi.lo + hi.hi in FP16
  static constexpr int A_COLS = 64;  // [0,32) half2(hi) | [32,64) half2(lo)
  static constexpr int NSLOT = 3;
  static constexpr int NSTAGE = 10;
  static constexpr int B_BYTES = 2 * KBLOCK_BYTES;  // hi | lo
};

template <int R>
struct Layout {
  using Rc = Recipe<R>;
  static constexpr int NMMA = Rc::NSLOT;  // a slot always belongs to the same issuer warp (fixed barrier ownership)
  static constexpr int ACC0 = TPC * Rc::A_COLS;
  static constexpr int SLOT_COLS = TPC * ACC_COLS;
  static constexpr int NBAR = 2 * Rc::NSTAGE + 2 * Rc::NSLOT + TPC + 1;
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)Rc::NSTAGE * Rc::B_BYTES + 6144 /*barriers, flags*/;
  static_assert(ACC0 + Rc::NSLOT * SLOT_COLS <= TMEM_COLS, "TMEM budget");
  static_assert(M0 + NMMA <= G0, "warp roles");
};

__device__ __host__ constexpr uint32_t idesc_tf32(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __host__ constexpr uint32_t idesc_f16(uint32_t n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem]; ACC is a compile-time constant (no predicate arithmetic in the issue loop)
template <bool F16, bool ACC>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  if (F16) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0)
        : "memory");
  }
}

// Experiment build (JD_NVCC_EXTRA=-DJD_TCM_TRACE, tools/tcm_trace.py): clock64 stamps of the hand-overs of CTA 0, per
// position: 0 producer past `empty` | 1 issuer past `tempty` | 2 issuer past `full` | 3 issuer done | 4 epilogue 0 past
// `tfull` | 5 epilogue 0 loads landed | 6 epilogue 0 released the slot | 7 epilogue 1 past `tfull` | 8 MMAs issued |
// 9 stage commit issued | 10 issuer warp reconverged | 11 epilogue 1 released the slot
#if defined(JD_TCM_TRACE)
constexpr int TRACE_POS = 2048;
__device__ long long g_trace[TRACE_POS * 16];
#define JD_TRACE(ev, pos)                                                                   \
  do {                                                                                      \
    if (blockIdx.x == 0 && (pos) < TRACE_POS) g_trace[(pos)*16 + (ev)] = clock64();         \
  } while (0)
#else
#define JD_TRACE(ev, pos) \
  do {                    \
  } while (0)
#endif

// Packed FP32 pairs (fma.rn.f32x2, sm_100): the epilogue is bound by the instructions its warps can issue, the squares
// of an accumulator row take 32 instead of 64 FMAs.
__device__ __forceinline__ void fma2_sq(float2& acc, float a0, float a1) {  // acc += (a0^2, a1^2)
  asm("{\n\t.reg .b64 va, vc;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vc, {%0, %1};\n\t"
      "fma.rn.f32x2 vc, va, va, vc;\n\t"
      "mov.b64 {%0, %1}, vc;\n\t}"
      : "+f"(acc.x), "+f"(acc.y)
      : "f"(a0), "f"(a1));
}
// (d0, d1) = (a0, a1) * s - (b0, b1)
__device__ __forceinline__ void fma2_sub(float& d0, float& d1, float a0, float a1, float s, float b0, float b1) {
  asm("{\n\t.reg .b64 va, vs, vb, vd;\n\t"
      "mov.b64 va, {%2, %3};\n\t"
      "mov.b64 vs, {%4, %4};\n\t"
      "mov.b64 vb, {%5, %6};\n\t"
      "fma.rn.f32x2 vd, va, vs, vb;\n\t"
      "mov.b64 {%0, %1}, vd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(s), "f"(-b0), "f"(-b1));
}

__device__ __forceinline__ int seg_rotation(int cl, int len, unsigned mul) {
  return (int)(((unsigned)cl * mul) % (unsigned)len);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  return (uint32_t)__half_as_ushort(__float2half_rn(a)) | ((uint32_t)__half_as_ushort(__float2half_rn(b)) << 16);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
      "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}

// ---------------------------------------------------------------- the MMAs of one tile and component
// a: TMEM address of the tile's A operand, d: its accumulator, b: low descriptor word of the staged operand image.
template <int R, bool TRI>
__device__ __forceinline__ void issue_tile(uint32_t d, uint32_t a, uint32_t b) {
  if (R == 0) {
    const uint32_t b_h = b + (2 * KBLOCK_BYTES >> 4), b_r = b_h + (KBLOCK_BYTES >> 4);
    // small terms first: xr . half(L'), half(x') . half(Lr), then tf32(x') . tf32(L').  Upper-triangular Lw: input
    // features [16kk, 16kk+16) only reach whitened features >= 16kk, the MMA's N extent shrinks accordingly
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t n0 = TRI ? 16u * kk : 0u;
      const uint32_t off16 = (kk * 32 + n0 * 128) >> 4;
      if (kk == 0)
        umma_ts<true, false>(d + n0, a + 64 + kk * 8, desc_from_lo(b_h + off16), idesc_f16(64 - n0));
      else
        umma_ts<true, true>(d + n0, a + 64 + kk * 8, desc_from_lo(b_h + off16), idesc_f16(64 - n0));
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t n0 = TRI ? 16u * kk : 0u;
      const uint32_t off16 = (kk * 32 + n0 * 128) >> 4;
      umma_ts<true, true>(d + n0, a + 96 + kk * 8, desc_from_lo(b_r + off16), idesc_f16(64 - n0));
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t n0 = TRI ? 16u * (kk >> 1) : 0u;
      const uint32_t off16 = ((kk >> 2) * KBLOCK_BYTES + (kk & 3) * 32 + n0 * 128) >> 4;
      umma_ts<false, true>(d + n0, a + kk * 8, desc_from_lo(b + off16), idesc_tf32(64 - n0));
    }
  } else {
    const uint32_t b_hi = b, b_lo = b + (KBLOCK_BYTES >> 4);
    // small terms first: lo . hi, hi . lo, then hi . hi
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t a_col = pass == 0 ? 32u : 0u;
      const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t n0 = TRI ? 16u * kk : 0u;
        const uint32_t off16 = (kk * 32 + n0 * 128) >> 4;
        if (pass == 0 && kk == 0)
          umma_ts<true, false>(d + n0, a + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_f16(64 - n0));
        else
          umma_ts<true, true>(d + n0, a + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_f16(64 - n0));
      }
    }
  }
}

// ---------------------------------------------------------------- the forward kernel
template <int R, bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tcx2_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                    const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck,
                    const float* __restrict__ binv, int K, int marginalize, int chunk, int smax, unsigned rot_mul,
                    unsigned* __restrict__ counters, float* __restrict__ ws_m, float* __restrict__ ws_s,
                    int* __restrict__ ws_k, float* __restrict__ value, int32_t* __restrict__ argmax,
                    float* __restrict__ logp, double* __restrict__ sum) {
  using Rc = Recipe<R>;
  using L = Layout<R>;
  constexpr int NSTAGE = Rc::NSTAGE, NSLOT = Rc::NSLOT, A_COLS = Rc::A_COLS, B_BYTES = Rc::B_BYTES, ACC0 = L::ACC0;
  constexpr int NMMA = L::NMMA, SLOT_COLS = L::SLOT_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;  // NSTAGE operand images
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + L::NBAR);
  int* s_flag = reinterpret_cast<int*>(s_tmem + 2);  // one per epilogue group
  // per-row flags of the gather warps, double-buffered on the segment parity: [parity][tile][row]
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);
  float* s_rinv = reinterpret_cast<float*>(s_valid + 2 * TPC * TM);  // 1 / (row scale)

  // warp index through a shuffle: the compiler then knows it is warp-uniform (role dispatch, positions, stage / slot
  // indices and with them every MMA operand stay in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
#if defined(JD_TCM_TRACE)
  const int dbg = marginalize >> 8;  // trace build, JD_TC_DEBUG & 1 (timing only, wrong results): no epilogue TMEM loads
#endif
  marginalize &= 1;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int t) { return bar0 + 8u * (2 * NSTAGE + t); };
  auto tempty_bar = [&](int t) { return bar0 + 8u * (2 * NSTAGE + NSLOT + t); };
  auto afull_bar = [&](int h) { return bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + h); };  // A operand of tile h stored
  const uint32_t aempty_bar = bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + TPC);             // last MMAs of the segment done
  const uint32_t crank = cluster_ctarank();

  // this CTA pair's chunk of the linearised (tile group, component) space; tile group tp = tiles [4 tp, 4 tp + 4):
  // CTA `crank` of the pair takes tiles 4 tp + 2 crank + {0, 1}
  const int cl = blockIdx.x / CLUSTER;
  const int n_tiles = (g.P + TM - 1) / TM, n_groups = (n_tiles + CLUSTER * TPC - 1) / (CLUSTER * TPC);
  const long long w_tot = (long long)n_groups * K;
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
    for (int t = 0; t < NSLOT; ++t) {
      mbar_init(tfull_bar(t), 1);
      mbar_init(tempty_bar(t), 4 * TPC);  // one arrive per epilogue warp (both tiles' accumulators share the slot)
    }
    for (int h = 0; h < TPC; ++h) mbar_init(afull_bar(h), 4);  // one arrive per gather warp
    mbar_init(aempty_bar, NMMA);                                // last MMAs of the segment, every issuer
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

  // segment geometry of tile group tp: positions [pos_lo, pos_hi) of this CTA's chunk, first component ka
#define JD_SEGMENT(tp)                                                                              \
  const long long comp0 = (long long)(tp)*K;                                                         \
  const int pos_lo = (int)((lin_begin > comp0 ? lin_begin : comp0) - lin_begin);                     \
  const int pos_hi = (int)((lin_end < comp0 + K ? lin_end : comp0 + K) - lin_begin);                 \
  const int ka = (int)(lin_begin + pos_lo - comp0), len = pos_hi - pos_lo;                           \
  const int rot = seg_rotation(cl, len, rot_mul);                                                    \
  const int sidx = (tp)-tp_first, par = sidx & 1;                                                    \
  (void)ka; (void)rot; (void)par;

  if (warp == 0) {
    // ===================== bulk-TMA producer (one warp, every position) ============================
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      for (int pos = pos_lo; pos < pos_hi; ++pos) {
        int idx = pos - pos_lo + rot;
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int s = pos % NSTAGE;
        mbar_wait(empty_bar(s), ((pos / NSTAGE) & 1) ^ 1);
        if (elect_one()) {
          JD_TRACE(0, pos);
          // this CTA fetches half `crank` of the image for both CTAs of the pair.  Recipe 0, upper-triangular factors:
          // whitened features 0..31 do not depend on input features 32..63, i.e. rows 0..31 of the second tf32 k-block
          // (4 KB) are zeros no MMA reads (k-steps 4..7 start at row 32): not copied.
          const bool trim = R == 0 && TRI;
          mbar_arrive_expect_tx(full_bar(s), trim ? B_BYTES - KBLOCK_BYTES / 2 : B_BYTES);
          const uint32_t dst = smem_u32(sB + s * B_BYTES) + crank * (B_BYTES / CLUSTER);
          const uint8_t* src = Bt + (size_t)kc * B_BYTES + crank * (B_BYTES / CLUSTER);
          const uint16_t both = (uint16_t)((1u << CLUSTER) - 1);
          if (trim && crank == 0) {
            bulk_g2s_mc(dst, src, KBLOCK_BYTES, full_bar(s), both);
            bulk_g2s_mc(dst + KBLOCK_BYTES + KBLOCK_BYTES / 2, src + KBLOCK_BYTES + KBLOCK_BYTES / 2, KBLOCK_BYTES / 2,
                        full_bar(s), both);
          } else {
            bulk_g2s_mc(dst, src, B_BYTES / CLUSTER, full_bar(s), both);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp >= M0 && warp < M0 + NMMA) {
    // ===================== MMA issuers: position pos (both tiles) is issued by warp M0 + pos % NMMA into slot
    // pos % NSLOT (warp-uniform control flow, one elected lane issues) =================================
    const int w = warp - M0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
#pragma unroll
      for (int h = 0; h < TPC; ++h) mbar_wait(afull_bar(h), par);  // the gather warps have stored this segment's A operands
      tc_fence_after();
      int pos = pos_lo + ((pos_lo % NMMA) == w ? 0 : (w - (pos_lo % NMMA) + NMMA) % NMMA);
      bool released = false;
      for (; pos < pos_hi; pos += NMMA) {
        const int s = pos % NSTAGE, t = pos % NSLOT;
        const bool last = pos + NMMA >= pos_hi;
        mbar_wait(tempty_bar(t), ((pos / NSLOT) & 1) ^ 1);
        if (lane == 0) JD_TRACE(1, pos);
        mbar_wait(full_bar(s), (pos / NSTAGE) & 1);
        if (lane == 0) JD_TRACE(2, pos);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b = sB_lo0 + s * (B_BYTES >> 4), d = tmem_u + ACC0 + t * SLOT_COLS;
#pragma unroll
          for (int h = 0; h < TPC; ++h)  // both tiles against the same staged image
            issue_tile<R, TRI>(d + h * ACC_COLS, tmem_u + h * A_COLS, b);
          JD_TRACE(8, pos);
          umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));  // stage free in both CTAs of the pair
          JD_TRACE(9, pos);
          umma_commit(tfull_bar(t));                                       // both accumulators of the slot complete
          if (last) umma_commit(aempty_bar);  // this warp's last read of the A operands (same thread as its MMAs)
          JD_TRACE(3, pos);
        }
        released = released || last;
        __syncwarp();
        if (lane == 0) JD_TRACE(10, pos);
      }
      if (!released) {  // no position of this segment fell to this warp
        if (elect_one()) mbar_arrive(aempty_bar);
        __syncwarp();
      }
    }
  } else if (warp >= G0 && warp < E0) {
    // ===================== gather: thread = patch row; 64 loads, mean, scale, split, tcgen05.st ==============
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
#pragma unroll 1
      for (int h = 0; h < TPC; ++h) {
        const int tile = (tp * CLUSTER + (int)crank) * TPC + h;
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
        // the A operands are free once the last MMAs of the previous segment have completed (tile 0's patch is already in
        // registers by then)
        if (h == 0) {
          mbar_wait(aempty_bar, par ^ 1);
          tc_fence_after();
        }
        const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16) + h * A_COLS;
        if (R == 0) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float xt[32], pk[32];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a0 = vals[hh * 32 + 2 * i] * sA, a1 = vals[hh * 32 + 2 * i + 1] * sA;
              const float t0 = tf32_rna(a0), t1 = tf32_rna(a1);
              xt[2 * i] = t0, xt[2 * i + 1] = t1;
              pk[i] = __uint_as_float(pack_half2(a0 - t0, a1 - t1));
              pk[16 + i] = __uint_as_float(pack_half2(a0, a1));
            }
            tmem_st32(a_lane + hh * 32, xt);
            // 16 packed words each: features [32hh, 32hh+32) land in columns [64 + 16hh, +16) and [96 + 16hh, +16)
            tmem_st16(a_lane + 64 + hh * 16, pk);
            tmem_st16(a_lane + 96 + hh * 16, pk + 16);
          }
        } else {
          float hi[32], lo[32];  // 32 packed half2 words each (64 features)
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float x0 = vals[2 * c] * sA, x1 = vals[2 * c + 1] * sA;
            const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
            hi[c] = __uint_as_float((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16));
            lo[c] = __uint_as_float(pack_half2(x0 - __half2float(h0), x1 - __half2float(h1)));
          }
          tmem_st32(a_lane, hi);
          tmem_st32(a_lane + 32, lo);
        }
        tmem_st_wait();
        s_valid[(par * TPC + h) * TM + row] = ok ? 1 : 0;
        s_rinv[(par * TPC + h) * TM + row] = ldexpf(1.f, e - 13);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(afull_bar(h));
      }
    }
  } else if (warp >= E0) {
    // ===================== epilogue: group 0 (warps 8-11) tile 0, group 1 (12-15) tile 1, every position ========
    const int grp = warp >= E0 + 4 ? 1 : 0;
    const int bar_id = 2 + grp;  // named barrier of this group's 128 threads
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      JD_SEGMENT(tp)
      const int tile = (tp * CLUSTER + (int)crank) * TPC + grp;
      const int64_t p = (int64_t)tile * TM + row;
      float run_m = -CUDART_INF_F, run_s = 0.f;
      int run_k = 0x7fffffff;
      float row_inv = 0.f;
      bool ok = false;
      for (int pos = pos_lo; pos < pos_hi; ++pos) {
        int idx = pos - pos_lo + rot;  // same rotated component order as the producers
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int t = pos % NSLOT;
        const float c_k = __ldg(ck + kc);
        const float b_inv = __ldg(binv + kc);
        mbar_wait(tfull_bar(t), (pos / NSLOT) & 1);
        if (lane == 0 && (warp == E0 || warp == E0 + 4)) JD_TRACE(warp == E0 ? 4 : 7, pos);
        tc_fence_after();
        if (pos == pos_lo) {  // written by the gather warps before the first MMA of the segment
          row_inv = s_rinv[(par * TPC + grp) * TM + row];
          ok = s_valid[(par * TPC + grp) * TM + row] != 0;
        }
        const float inv = row_inv * b_inv;  // undo the row and component scales
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ACC0 + t * SLOT_COLS + grp * ACC_COLS;
        float y0[32], y1[32];
#if defined(JD_TCM_TRACE)
        if (dbg & 1) {
#pragma unroll
          for (int i = 0; i < 32; ++i) y0[i] = y1[i] = (float)lane;
        } else
#endif
        {
          tmem_ld32(taddr, y0);
          tmem_ld32(taddr + 32, y1);
          tmem_ld_wait();
        }
        if (lane == 0 && warp == E0) JD_TRACE(5, pos);
        float2 qa = make_float2(0.f, 0.f), qb = qa, qc = qa, qd = qa;  // eight independent chains of packed squares
        float lp;
        if (ZERO_MEAN) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            fma2_sq(qa, y0[i], y0[i + 1]);
            fma2_sq(qb, y1[i], y1[i + 1]);
            fma2_sq(qc, y0[i + 2], y0[i + 3]);
            fma2_sq(qd, y1[i + 2], y1[i + 3]);
          }
          // the scales are powers of two: applying them to the sum is exact
          lp = fmaf(-0.5f * (inv * inv), ((qa.x + qa.y) + (qb.x + qb.y)) + ((qc.x + qc.y) + (qd.x + qd.y)), c_k);
        } else {
          // (mw_k through the read-only path: every lane reads the same 16 x 16 B, an L1 broadcast)
          const float4* mwk = reinterpret_cast<const float4*>(mw + (size_t)kc * 64);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 b0 = __ldg(mwk + c4), b1 = __ldg(mwk + 8 + c4);
            float d0, d1, d2, d3, e0, e1, e2, e3;
            fma2_sub(d0, d1, y0[4 * c4], y0[4 * c4 + 1], inv, b0.x, b0.y);
            fma2_sub(d2, d3, y0[4 * c4 + 2], y0[4 * c4 + 3], inv, b0.z, b0.w);
            fma2_sub(e0, e1, y1[4 * c4], y1[4 * c4 + 1], inv, b1.x, b1.y);
            fma2_sub(e2, e3, y1[4 * c4 + 2], y1[4 * c4 + 3], inv, b1.z, b1.w);
            fma2_sq(qa, d0, d1);
            fma2_sq(qb, e0, e1);
            fma2_sq(qc, d2, d3);
            fma2_sq(qd, e2, e3);
          }
          lp = fmaf(-0.5f, ((qa.x + qa.y) + (qb.x + qb.y)) + ((qc.x + qc.y) + (qd.x + qd.y)), c_k);
        }
        tc_fence_before();  // the slot is free once consumed (lp depends on every load, see mbar_arrive_after)
        __syncwarp();
        if (lane == 0) mbar_arrive_after(tempty_bar(t), lp, rt_zero);
        if (lane == 0 && (warp == E0 || warp == E0 + 4)) JD_TRACE(warp == E0 ? 6 : 11, pos);
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

      // ---- segment end: this group's tile is complete over [ka, ka + len)
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
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (row == 0) s_flag[grp] = (atomicAdd(&counters[tile], 1u) + 1u == (unsigned)nseg) ? 1 : 0;
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        final_here = s_flag[grp] != 0;
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
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");  // s_flag is read before the next segment rewrites it
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

// ---------------------------------------------------------------- plan: CTA pairs, chunk of (tile group, component), workspace
struct Plan {
  int n_clusters, chunk, smax, n_tiles2;
  size_t off_m, off_s, off_k, bytes;
};
static Plan plan(int64_t P, int K, int clusters = 0) {
  Plan p;
  const int64_t n_tiles = (P + TM - 1) / TM, n_groups = (n_tiles + CLUSTER * TPC - 1) / (CLUSTER * TPC);
  const int64_t w_tot = n_groups * K;
  int C = num_sms() / CLUSTER;  // one CTA per SM, two SMs per pair
  if (const char* e = getenv("JD_TCM_CLUSTERS")) C = atoi(e) > 0 ? atoi(e) : C;
  if (clusters > 0 && clusters < C) C = clusters;  // fewer pairs: longer chunks, never a larger workspace
  int64_t chunk = (w_tot + C - 1) / C;
  if (K % 8 == 0) chunk = (chunk + 7) / 8 * 8;  // no segment shorter than 8 components
  if (chunk < 1) chunk = 1;
  p.chunk = (int)chunk;
  p.n_clusters = (int)((w_tot + chunk - 1) / chunk);
  p.smax = (int)((K + chunk - 1) / chunk) + 1;
  p.n_tiles2 = (int)(n_groups * CLUSTER * TPC);
  const size_t cnt = ((size_t)p.n_tiles2 * sizeof(unsigned) + 255) / 256 * 256;
  const size_t part = (size_t)p.n_tiles2 * p.smax * TM * sizeof(float);
  p.off_m = cnt;
  p.off_s = cnt + part;
  p.off_k = cnt + 2 * part;
  p.bytes = cnt + 3 * part;
  return p;
}

template <int R>
static int launch(const char* who, const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                  int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K, int upper_tri,
                  int zero_mean, int marginalize, void* workspace, float* value, int32_t* argmax, float* logp,
                  double* sum, jd_stream_t stream, int clusters = 0) {
  JD_CHECK_ARG(flux && Bt && binv && mw && ck && workspace && K > 0, "%s: null pointer", who);
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "%s: bad geometry", who);
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end, "%s: bad patch-row block [%d,%d) of %d", who,
               row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "%s: Bt and mw must be 16-byte aligned, the workspace 256-byte aligned", who);
  tcx::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  const void* kerns[4] = {(const void*)gmm_fwd_tcx2_kernel<R, false, false>, (const void*)gmm_fwd_tcx2_kernel<R, false, true>,
                          (const void*)gmm_fwd_tcx2_kernel<R, true, false>, (const void*)gmm_fwd_tcx2_kernel<R, true, true>};
  static bool attr_set[64] = {};  // per device (and per recipe: the static lives in the template instance)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Layout<R>::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("%s: cannot reserve %zu B of shared memory: %s", who, Layout<R>::SMEM_BYTES, cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
    attr_set[dev & 63] = true;
  }
  const Plan p = plan(g.P, K, clusters);
  static int rot_env = -1, dbg = -1;
  if (rot_env < 0) {
    const char* e = getenv("JD_TC_SK_ROT");  // 0: visit the components of a segment in ascending order
    rot_env = e ? atoi(e) : 40503;
  }
  if (dbg < 0) {
    const char* e = getenv("JD_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_clusters * CLUSTER);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = Layout<R>::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
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
  const int marg = (marginalize ? 1 : 0) | ((dbg & 1) << 8);  // (the knob only exists in the JD_TCM_TRACE build)
  cudaError_t le;
  if (upper_tri && zero_mean)
    le = cudaLaunchKernelEx(&cfg, gmm_fwd_tcx2_kernel<R, true, true>, flux, g, shift_yx, bt8, mw, ck, binv, K, marg,
                            p.chunk, p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  else if (upper_tri)
    le = cudaLaunchKernelEx(&cfg, gmm_fwd_tcx2_kernel<R, true, false>, flux, g, shift_yx, bt8, mw, ck, binv, K, marg,
                            p.chunk, p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  else if (zero_mean)
    le = cudaLaunchKernelEx(&cfg, gmm_fwd_tcx2_kernel<R, false, true>, flux, g, shift_yx, bt8, mw, ck, binv, K, marg,
                            p.chunk, p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  else
    le = cudaLaunchKernelEx(&cfg, gmm_fwd_tcx2_kernel<R, false, false>, flux, g, shift_yx, bt8, mw, ck, binv, K, marg,
                            p.chunk, p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("%s: launch failed: %s", who, cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH(who);
  return JD_OK;
}

}  // namespace tcx2
}  // namespace jd

using namespace jd;

extern "C" {

#if defined(JD_TCM_TRACE)
int jd_debug_tcm2_trace(long long* host_dst, int n_positions) {
  if (n_positions > tcx2::TRACE_POS) n_positions = tcx2::TRACE_POS;
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_dst, tcx2::g_trace, (size_t)n_positions * 16 * sizeof(long long)) == cudaSuccess ? 0 : 1;
}
#endif

int64_t jd_gmm_tcm2_workspace_bytes(int64_t P, int K) {
  if (P <= 0 || K <= 0) return 0;
  return (int64_t)tcx2::plan(P, K).bytes;
}

int jd_gmm_prior_forward_tcm2(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                              int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                              int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                              int32_t* argmax, float* logp, double* sum, jd_stream_t stream) {
  return tcx2::launch<0>("jd_gmm_prior_forward_tcm2", flux, fH, fW, shift_yx, stride, row_begin, row_end, Bt, binv, mw,
                         ck, K, upper_tri, zero_mean, marginalize, workspace, value, argmax, logp, sum, stream);
}

int jd_gmm_prior_forward_tcx2_on(int recipe, int clusters, const float* flux, int fH, int fW, const int32_t* shift_yx,
                                 int stride, int row_begin, int row_end, const void* Bt, const float* binv,
                                 const float* mw, const float* ck, int K, int upper_tri, int zero_mean, int marginalize,
                                 void* workspace, float* value, int32_t* argmax, float* logp, double* sum,
                                 jd_stream_t stream) {
  JD_CHECK_ARG((recipe == 0 || recipe == 1) && clusters >= 0, "jd_gmm_prior_forward_tcx2_on: recipe 0 / 1, clusters >= 0");
  if (recipe == 0)
    return tcx2::launch<0>("jd_gmm_prior_forward_tcx2_on", flux, fH, fW, shift_yx, stride, row_begin, row_end, Bt, binv,
                           mw, ck, K, upper_tri, zero_mean, marginalize, workspace, value, argmax, logp, sum, stream,
                           clusters);
  return tcx2::launch<1>("jd_gmm_prior_forward_tcx2_on", flux, fH, fW, shift_yx, stride, row_begin, row_end, Bt, binv, mw,
                         ck, K, upper_tri, zero_mean, marginalize, workspace, value, argmax, logp, sum, stream, clusters);
}

int jd_gmm_prior_forward_tc16x2(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                                int row_end, const void* Bt16, const float* binv, const float* mw, const float* ck,
                                int K, int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                                int32_t* argmax, float* logp, double* sum, jd_stream_t stream) {
  return tcx2::launch<1>("jd_gmm_prior_forward_tc16x2", flux, fH, fW, shift_yx, stride, row_begin, row_end, Bt16, binv,
                         mw, ck, K, upper_tri, zero_mean, marginalize, workspace, value, argmax, logp, sum, stream);
}

}  // extern "C"
