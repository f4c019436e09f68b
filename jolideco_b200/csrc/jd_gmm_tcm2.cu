// GMM patch prior forward on tcgen05, fourth generation: the mixed TF32 / FP16 split of jd_gmm_tcm.cu with TWO patch
// tiles per CTA and B stage.
//
// Why.  ncu on the one-tile kernel at 1024^2 / K = 256 (profiles/r02_ncu_joint1024_step.csv): 3.75 GB cross the
// crossbar into the SMs per launch (l1tex__m_xbar2l1tex_read_bytes) = 41 B/clk/SM for 314 us, the tensor pipe is 65 %
// busy, and switching the epilogue's TMEM loads or all but one MMA off (JD_TC_DEBUG) barely moves the time: every SM has
// to take in the whole 28 KB operand image of a component for 128 patches of work, and a CTA-pair multicast does not
// change what ONE SM must receive.  tools/ubench.cu: TMEM reads sustain > 400 B/clk/SM (not the limiter).  Processing
// two tiles (256 patches) against each staged image halves the bytes per unit of work and the number of
// producer / issuer / epilogue hand-overs.
//
// What changes against jd_gmm_tcm.cu:
//   * work unit = tile group of CLUSTER x TPC = 4 tiles; the stream-K space is (tile group, component);
//   * TMEM (512 columns): [0,128) A of tile 0 | [128,256) A of tile 1 | two accumulator slots of 2 x 64 columns.  The A
//     operand is single-buffered: at a segment boundary the gather warps (which prefetch tile 0 into registers while
//     the last MMAs of the previous segment run) store once the issuers' last commit has fired;
//   * epilogue group A (warps 8-11) owns tile 0, group B (12-15) tile 1, both take every position: no merge between the
//     groups, each finishes (or stream-K-merges) its own tile.
// Precision recipe, operand image (jd_gmm_tcm_pack), barriers-on-position-counter scheme: unchanged.
//
// Warps (512 threads, one CTA per SM): 0-1 bulk-TMA producers | 2-3 MMA issuers | 4-7 gather | 8-11 epilogue of tile 0
// | 12-15 epilogue of tile 1.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_tc_ptx.cuh"

namespace jd {
namespace tcm2 {

using namespace tcx;

constexpr int TM = 128;
constexpr int TPC = 2;                       // patch tiles per CTA and B stage
constexpr int NSTAGE = 6;
constexpr int NSLOT = 2;                     // accumulator slots (positions in flight between issuers and epilogue)
constexpr int ACC_COLS = 64;
constexpr int SLOT_COLS = TPC * ACC_COLS;    // both tiles' accumulators of one position
constexpr int A_COLS = 128;                  // per tile
constexpr int ACC0 = TPC * A_COLS;
constexpr int TMEM_COLS = 512;
constexpr int KBLOCK_BYTES = 64 * 128;           // 8 KB: 64 rows x 128 B
constexpr int B_TF32_BYTES = 2 * KBLOCK_BYTES;   // 64 x 64 tf32
constexpr int B_F16_BYTES = KBLOCK_BYTES;        // 64 x 64 half
constexpr int B_BYTES = B_TF32_BYTES + 2 * B_F16_BYTES;
constexpr int CLUSTER = 2;
constexpr int NPROD = 2;
constexpr int NMMA = 2;
static_assert(NSTAGE % NPROD == 0 && NSLOT % NMMA == 0, "fixed barrier ownership (jd_gmm_tc.cu)");
static_assert(ACC0 + NSLOT * SLOT_COLS <= 512, "TMEM budget");
constexpr int M0 = NPROD;      // first MMA warp (owns the TMEM allocation)
constexpr int G0 = 4;          // first gather warp
constexpr int E0 = 8;          // first epilogue warp
constexpr int NTHREADS = 512;
constexpr int MW_BYTES = 64 * 4;
constexpr int NBAR = 2 * NSTAGE + 3 * NSLOT + TPC + 1;
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

__device__ __forceinline__ int seg_rotation(int cl, int len, unsigned mul) {
  return (int)(((unsigned)cl * mul) % (unsigned)len);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  return (uint32_t)__half_as_ushort(__float2half_rn(a)) | ((uint32_t)__half_as_ushort(__float2half_rn(b)) << 16);
}

// ---------------------------------------------------------------- the forward kernel
template <bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tcm2_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
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
  int* s_flag = reinterpret_cast<int*>(s_tmem + 2);           // one per epilogue group
  // per-row flags of the gather warps, double-buffered on the segment parity: [parity][tile][row]
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);
  float* s_rinv = reinterpret_cast<float*>(s_valid + 2 * TPC * TM);  // 1 / (row scale)

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
  auto afull_bar = [&](int h) { return bar0 + 8u * (2 * NSTAGE + 3 * NSLOT + h); };  // A operand of tile h stored
  const uint32_t aempty_bar = bar0 + 8u * (2 * NSTAGE + 3 * NSLOT + TPC);             // last MMAs of the segment done
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
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4 * TPC);  // one arrive per epilogue warp (both tiles' accumulators share the slot)
      mbar_init(mwfull_bar(s), 1);
    }
    for (int h = 0; h < TPC; ++h) mbar_init(afull_bar(h), 4);  // one arrive per gather warp
    mbar_init(aempty_bar, NMMA);                                // last MMAs of the segment, both issuers
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
#pragma unroll
      for (int h = 0; h < TPC; ++h) mbar_wait(afull_bar(h), par);  // the gather warps have stored this segment's A operands
      tc_fence_after();
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
#pragma unroll
          for (int h = 0; h < TPC; ++h) {  // both tiles against the same staged image
            const uint32_t a_base = tmem_u + h * A_COLS;
            const uint32_t d = tmem_u + ACC0 + t * SLOT_COLS + h * ACC_COLS;
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
          }
          umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));  // stage free in both CTAs of the pair
          umma_commit(tfull_bar(t));                                       // both accumulators of the slot complete
          if (last) umma_commit(aempty_bar);  // this warp's last read of the A operands (same thread as its MMAs)
        }
        released = released || last;
        __syncwarp();
      }
      if (!released) {  // no position of this segment fell to this warp
        if (elect_one()) mbar_arrive(aempty_bar);
        __syncwarp();
      }
    }
  } else if (warp < E0) {
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
      s_valid[(par * TPC + h) * TM + row] = ok ? 1 : 0;
      s_rinv[(par * TPC + h) * TM + row] = ldexpf(1.f, e - 13);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(afull_bar(h));
      }
    }
  } else {
    // ===================== epilogue: group A (warps 8-11) tile 0, group B (12-15) tile 1, every position ========
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
        if (!ZERO_MEAN) mbar_wait(mwfull_bar(t), (pos / NSLOT) & 1);
        mbar_wait(tfull_bar(t), (pos / NSLOT) & 1);
        tc_fence_after();
        if (pos == pos_lo) {  // written by the gather warps before the first MMA of the segment
          row_inv = s_rinv[(par * TPC + grp) * TM + row];
          ok = s_valid[(par * TPC + grp) * TM + row] != 0;
        }
        const float inv = row_inv * b_inv;  // undo the row and component scales
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ACC0 + t * SLOT_COLS + grp * ACC_COLS;
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
static Plan plan(int64_t P, int K) {
  Plan p;
  const int64_t n_tiles = (P + TM - 1) / TM, n_groups = (n_tiles + CLUSTER * TPC - 1) / (CLUSTER * TPC);
  const int64_t w_tot = n_groups * K;
  int C = num_sms() / CLUSTER;  // one CTA per SM, two SMs per pair
  if (const char* e = getenv("JD_TCM_CLUSTERS")) C = atoi(e) > 0 ? atoi(e) : C;
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

}  // namespace tcm2
}  // namespace jd

using namespace jd;

extern "C" {

int64_t jd_gmm_tcm2_workspace_bytes(int64_t P, int K) {
  if (P <= 0 || K <= 0) return 0;
  return (int64_t)tcm2::plan(P, K).bytes;
}

int jd_gmm_prior_forward_tcm2(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                             int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                             int upper_tri, int zero_mean, int marginalize, void* workspace, float* value,
                             int32_t* argmax, float* logp, double* sum, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && binv && mw && ck && workspace && K > 0, "jd_gmm_prior_forward_tcm2: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tcm2: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tcm2: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "jd_gmm_prior_forward_tcm2: Bt and mw must be 16-byte aligned, the workspace 256-byte aligned");
  tcx::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};  // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev & 63]) {
    cudaError_t e = cudaSuccess;
    const void* kerns[4] = {(const void*)tcm2::gmm_fwd_tcm2_kernel<false, false>, (const void*)tcm2::gmm_fwd_tcm2_kernel<false, true>,
                            (const void*)tcm2::gmm_fwd_tcm2_kernel<true, false>, (const void*)tcm2::gmm_fwd_tcm2_kernel<true, true>};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcm2::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tcm2: cannot reserve %zu B of shared memory: %s", tcm2::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
    attr_set[dev & 63] = true;
  }
  const tcm2::Plan p = tcm2::plan(g.P, K);
  static int rot_env = -1;
  if (rot_env < 0) {
    const char* e = getenv("JD_TC_SK_ROT");  // 0: visit the components of a segment in ascending order
    rot_env = e ? atoi(e) : 40503;
  }
  auto kern = upper_tri ? (zero_mean ? tcm2::gmm_fwd_tcm2_kernel<true, true> : tcm2::gmm_fwd_tcm2_kernel<true, false>)
                        : (zero_mean ? tcm2::gmm_fwd_tcm2_kernel<false, true> : tcm2::gmm_fwd_tcm2_kernel<false, false>);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_clusters * tcm2::CLUSTER);
  cfg.blockDim = dim3(tcm2::NTHREADS);
  cfg.dynamicSmemBytes = tcm2::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tcm2::CLUSTER;
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
  if (dbg & 16) kern = zero_mean ? tcm2::gmm_fwd_tcm2_kernel<false, true> : tcm2::gmm_fwd_tcm2_kernel<false, false>;  // dense
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, flux, g, shift_yx, bt8, mw, ck, binv, K, (marginalize ? 1 : 0) | (dbg << 8), p.chunk,
                                      p.smax, (unsigned)rot_env, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_forward_tcm2: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tcm2");
  return JD_OK;
}

}  // extern "C"
