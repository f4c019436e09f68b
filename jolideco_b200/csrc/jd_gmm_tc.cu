// GMM patch prior forward on the 5th-generation tensor cores (tcgen05 / TMEM / bulk-TMA), sm_100a.
//
// The K-component Mahalanobis evaluation is one dense contraction (patches x 64) . (64 x 64 K).  It
// runs as split-TF32 ("3xTF32": x = hi + lo with both halves exactly representable in TF32,
// D += lo.hi + hi.lo + hi.hi, FP32 accumulation in TMEM) so that the result keeps FP32 accuracy
// (the per-iteration parity bar is 1e-5, single-pass TF32 gives 1e-3).
//
// CTA = 128 patches (UMMA M = 128), 6 warps:
//   warp 0      bulk-TMA producer: streams the pre-packed B image of component k
//               (Lw_k^T split hi/lo, 128B-swizzled K-major, 32 KB) into a 4-stage smem ring;
//   warp 1      TMEM allocator + MMA issuer: 24 tcgen05.mma (M128 N<=64 K8, kind::tf32, A from TMEM,
//               B from smem) per component into one of 6 accumulator slots (64 TMEM columns each).
//               Lw_k is upper triangular (a Cholesky factor), so k-step kk only feeds whitened
//               features j >= 8 kk: the MMA N extent shrinks 64,64,48,48,32,32,16,16 (62.5 % of
//               the dense tensor work);
//   warps 2..5  gather the 128 patches from the flux image at rolled coordinates, subtract the
//               patch mean, split hi/lo and store the A operand straight into TMEM (tcgen05.st,
//               thread = patch row = TMEM lane; 128 columns) so the MMA never re-reads A from shared
//               memory; then act as the epilogue: tcgen05.ld the 128x64 accumulator of each component
//               (thread = patch row), subtract mw_k, square, reduce over the 64 whitened features,
//               add ck_k and fold into a running max/argmax or online logsumexp.  Y never leaves
//               the SM; only value/argmax (and optionally logp) are written.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_tc_ptx.cuh"

namespace jd {

int gmm_prior_forward_simt(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* Lw, const float* mw, const float* ck, int K, int marginalize,
                           float* value, int32_t* argmax, float* logp, double* sum, cudaStream_t st);

namespace tc {

constexpr int TM = 128;                 // patches per CTA
constexpr int NSTAGE = 6;               // B ring depth (multiple of NPROD)
constexpr int NSLOT = 6;                // TMEM accumulator slots (multiple of NMMA and of the 2 epilogue groups)
constexpr int SLOT_COLS = 64;
constexpr int A_COLS = 128;             // TMEM columns [0,64) = A hi, [64,128) = A lo
constexpr int TMEM_COLS = 512;          // A_COLS + NSLOT * SLOT_COLS
constexpr int KBLOCK_BYTES_B = 64 * 128;      // 64 rows x 128 B = 8 KB
constexpr int B_BYTES = 4 * KBLOCK_BYTES_B;   // hi(kb0,kb1) lo(kb0,kb1) = 32 KB per component
constexpr int CLUSTER = 2;              // CTA pair: each CTA fetches half of every B image and multicasts it
constexpr int NMMA = 3;                 // MMA-issuing warps: component position k is issued by warp M0 + k % NMMA
constexpr int NPROD = 2;                // bulk-TMA producer warps: position k is loaded by warp k % NPROD
// Every mbarrier is waited on by ONE fixed warp (group) across its successive uses: a waiter may be at most one
// phase ahead of the barrier (parity aliasing otherwise), which in-order processing by a single owner guarantees.
static_assert(NSTAGE % NPROD == 0, "a smem stage must always be refilled by the same producer warp");
static_assert(NSLOT % NMMA == 0 && NSLOT % 2 == 0, "a TMEM slot must always belong to the same MMA warp / epilogue group");
constexpr int M0 = NPROD;                // first MMA-issuer warp (also owns the TMEM allocation)
constexpr int E0 = NPROD + NMMA;         // first epilogue warp
constexpr int NTHREADS = 32 * (NPROD + NMMA + 8);  // producers, MMA issuers, 2 x 4 epilogue warps
constexpr int MW_BYTES = 64 * 4;        // mw_k staged per accumulator slot
constexpr int NSTAGE_BWD = 4;           // B ring depth of the logsumexp backward (room for the merge buffer)
constexpr int MERGE_BYTES = 64 * TM * 4; // gradient rows of epilogue group B
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + NSTAGE * B_BYTES + NSLOT * MW_BYTES + 4096 /*barriers etc.*/;
constexpr size_t SMEM_BYTES_BWD = 1024 + NSTAGE_BWD * B_BYTES + NSLOT * MW_BYTES + MERGE_BYTES + 4096;

using namespace tcx;

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __host__ constexpr uint32_t idesc_n(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// byte offset of element (row, d) inside a [rows x 64 tf32] operand stored as two 128B-swizzled k-blocks
__device__ __host__ __forceinline__ uint32_t sw128_offset(int row, int d, int kblock_bytes) {
  int kb = d >> 5, c = (d & 31) >> 2, e = d & 3;
  return kb * kblock_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4) + e * 4;
}

// ---------------------------------------------------------------- setup: pack Lw_k^T into the smem image
// Bt[k] (32 KB): hi kb0, hi kb1, lo kb0, lo kb1; row n = whitened feature j, K index = input feature i.
__global__ void pack_b_kernel(const float* __restrict__ Lw, int K, uint8_t* __restrict__ out) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)K * 4096) return;
  int k = (int)(idx >> 12), rem = (int)(idx & 4095), i = rem >> 6, j = rem & 63;
  float v = Lw[idx];  // Lw[k][i][j]
  float hi = tf32_rna(v);
  float lo = tf32_rna(v - hi);
  uint8_t* base = out + (size_t)k * B_BYTES;
  uint32_t off = sw128_offset(j, i, KBLOCK_BYTES_B);
  *reinterpret_cast<float*>(base + off) = hi;
  *reinterpret_cast<float*>(base + 2 * KBLOCK_BYTES_B + off) = lo;
}

// ---------------------------------------------------------------- the forward kernel
template <bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tc_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                  const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck, int K,
                  int marginalize, float* __restrict__ value, int32_t* __restrict__ argmax, float* __restrict__ logp,
                  double* __restrict__ sum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                          // NSTAGE x 32 KB
  float* sMW = reinterpret_cast<float*>(sB + NSTAGE * B_BYTES);  // NSLOT x 64 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES + NSLOT * MW_BYTES);
  // barrier indices: full[NSTAGE], empty[NSTAGE], tfull[NSLOT], tempty[NSLOT], mwfull[NSLOT]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3 * NSLOT);
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);       // 128 ints
  double* s_red = reinterpret_cast<double*>(s_valid + TM);  // 4 doubles
  float* s_mm = reinterpret_cast<float*>(s_red + 4);        // merge buffers of epilogue group B: max,
  float* s_ms = s_mm + TM;                                  //   sum-exp,
  int* s_mk = reinterpret_cast<int*>(s_ms + TM);            //   argmax

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
  const int dbg = marginalize >> 8;  // profiling knobs (JD_TC_DEBUG): 1 = no epilogue TMEM loads, 2 = one MMA per component
  const bool trim8 = (dbg & 8) != 0;  // triangular trim per k-step (N = 64, 56, .., 8) instead of per pair of k-steps
  marginalize &= 1;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NSLOT + s); };
  auto mwfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + s); };
  // Every CTA walks the components in a different cyclic order (start k0): at any instant the
  // CTAs stream different B images / mw rows, which spreads the L2 reads over the slices instead
  // of 148 SMs hammering the same lines in lockstep.  max / logsumexp do not depend on the order
  // (ties in max resolve to the lowest component index, as torch.max does).
  // The two CTAs of a cluster share every B image (each loads one half and multicasts it to both),
  // which halves the L2 -> SM traffic; they therefore walk the components in the same order.
  const uint32_t crank = cluster_ctarank();
  const int k0 = (int)(((long long)(blockIdx.x / CLUSTER) * K) / (gridDim.x / CLUSTER));

  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int64_t p0 = (int64_t)blockIdx.x * TM;

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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M0) tmem_alloc(smem_u32(s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // TMEM lane quarter of an epilogue/gather warp, and the patch row (= TMEM lane) of its threads
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int64_t p = p0 + row;

  if (warp < NPROD) {
    // ===================== bulk-TMA producers (whole warp waits, one elected lane issues) ==========
    int kc = k0 + warp;  // component handled at position k
    kc = kc >= K ? kc - K : kc;
    for (int k = warp; k < K; k += NPROD) {
      const int s = k % NSTAGE, t = k % NSLOT;
      mbar_wait(empty_bar(s), ((k / NSTAGE) & 1) ^ 1);
      // mw_k rides with accumulator slot t: free once the epilogue of component k - NSLOT is done
      if (!ZERO_MEAN) mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
      if (elect_one()) {
        // this CTA fetches half `crank` (hi or lo, 16 KB) of the image for both CTAs of the pair
        mbar_arrive_expect_tx(full_bar(s), B_BYTES);
        bulk_g2s_mc(smem_u32(sB + s * B_BYTES) + crank * (B_BYTES / CLUSTER),
                    Bt + (size_t)kc * B_BYTES + crank * (B_BYTES / CLUSTER), B_BYTES / CLUSTER, full_bar(s),
                    (uint16_t)((1u << CLUSTER) - 1));
        if (!ZERO_MEAN) {
          mbar_arrive_expect_tx(mwfull_bar(t), MW_BYTES);
          bulk_g2s(smem_u32(sMW + t * 64), mw + (size_t)kc * 64, MW_BYTES, mwfull_bar(t));
        }
      }
      __syncwarp();
      kc += NPROD;
      kc = kc >= K ? kc - K : kc;
    }
  } else if (warp >= E0 && warp < E0 + 4) {
    // ---- gather (epilogue group A): thread = patch row; 64 loads, mean, hi/lo split, tcgen05.st into TMEM lane `row`
    float vals[64];
    float s = 0.f;
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
          s += x;
          ok = ok && (x > -1e5f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) vals[i] = 0.f;
    }
    const float mean = s * (1.f / 64.f);
    const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = ok ? vals[h * 32 + i] - mean : 0.f;
        hi[i] = tf32_rna(x);
        lo[i] = tf32_rna(x - hi[i]);
      }
      tmem_st32(a_lane + h * 32, hi);
      tmem_st32(a_lane + 64 + h * 32, lo);
    }
    tmem_st_wait();
    s_valid[row] = ok ? 1 : 0;
    tc_fence_before();
    // the 4 gather warps -> MMA warp: named barrier 1 (128 gather threads + 32 MMA-warp threads)
    asm volatile("bar.arrive 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");
  }

  if (warp >= M0 && warp < M0 + NMMA) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =========
    asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");  // A operand is in TMEM
    tc_fence_after();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    for (int k = warp - M0; k < K; k += NMMA) {
      const int s = k % NSTAGE, t = k % NSLOT;
      mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
      mbar_wait(full_bar(s), (k / NSTAGE) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_hi = sB_lo0 + s * (B_BYTES >> 4), b_lo = b_hi + ((2 * KBLOCK_BYTES_B) >> 4);
        const uint32_t d = tmem_u + A_COLS + t * SLOT_COLS;
        uint32_t acc = 0;
        // small terms first: lo.hi, hi.lo, then hi.hi
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_col = pass == 0 ? 64u : 0u;
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if ((dbg & 2) && (pass > 0 || kk > 0)) continue;
            // upper-triangular Lw: input features [8kk, 8kk+8) only reach whitened features >= 8kk
            const uint32_t n0 = TRI ? (trim8 ? 8u * kk : 16u * (kk >> 1)) : 0u;
            const uint32_t off16 = ((kk >> 2) * KBLOCK_BYTES_B + (kk & 3) * 32 + n0 * 128) >> 4;
            umma_tf32_ts(d + n0, tmem_u + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_n(64 - n0), acc);
            acc = 1;
          }
        }
        umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));  // stage free in both CTAs of the pair
        umma_commit(tfull_bar(t));  // accumulator slot complete
      }
      __syncwarp();
    }
  } else if (warp >= E0) {
    // ===================== epilogue: group A (warps 2-5) takes even positions, group B odd =========
    const int grp = warp >= E0 + 4 ? 1 : 0;
    float run_m = -CUDART_INF_F, run_s = 0.f;
    int run_k = 0x7fffffff;
    int kc = k0 + grp;
    kc = kc >= K ? kc - K : kc;
    for (int k = grp; k < K; k += 2) {
      const int t = k % NSLOT;
      const float c_k = __ldg(ck + kc);
      if (!ZERO_MEAN) mbar_wait(mwfull_bar(t), (k / NSLOT) & 1);
      mbar_wait(tfull_bar(t), (k / NSLOT) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + A_COLS + t * SLOT_COLS;
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
      } else {
        const float4* mwk = reinterpret_cast<const float4*>(sMW + t * 64);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 b0 = mwk[c4], b1 = mwk[8 + c4];
          float d0 = y0[4 * c4] - b0.x, d1 = y0[4 * c4 + 1] - b0.y, d2 = y0[4 * c4 + 2] - b0.z, d3 = y0[4 * c4 + 3] - b0.w;
          float e0 = y1[4 * c4] - b1.x, e1 = y1[4 * c4 + 1] - b1.y, e2 = y1[4 * c4 + 2] - b1.z, e3 = y1[4 * c4 + 3] - b1.w;
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
      // accumulator slot and mw row are free once both are consumed (lp depends on every load, see mbar_arrive_after)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_after(tempty_bar(t), lp, rt_zero);
      if (logp && p < g.P) logp[(size_t)kc * g.P + p] = lp;  // component-major (K x P'): coalesced over patch rows
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
      kc += 2;
      kc = kc >= K ? kc - K : kc;
    }
    // merge group B into group A (named barrier 2 over the 256 epilogue threads)
    if (grp == 1) {
      s_mm[row] = run_m;
      s_ms[row] = run_s;
      s_mk[row] = run_k;
      asm volatile("bar.arrive 2, 256;" ::: "memory");
    } else {
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float om = s_mm[row], os = s_ms[row];
      const int ok_ = s_mk[row];
      if (marginalize) {
        const float m = fmaxf(run_m, om);
        run_s = run_s * expf(run_m - m) + (om == -CUDART_INF_F ? 0.f : os * expf(om - m));
        run_k = om > run_m ? ok_ : run_k;
        run_m = m;
      } else if (om > run_m || (om == run_m && ok_ < run_k)) {
        run_m = om;
        run_k = ok_;
      }
      double part = 0.0;
      if (p < g.P) {
        const bool ok = s_valid[row] != 0;
        float v = marginalize ? run_m + logf(run_s) : run_m;
        v = ok ? v : 0.f;
        if (value) value[p] = v;
        if (argmax) argmax[p] = ok ? run_k : -1;
        part = (double)v;
      }
      part = warp_sum(part);
      if (lane == 0) s_red[q] = part;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until here
  if (threadIdx.x == 0 && sum) atomicAdd(sum, s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  if (warp == M0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- stream-K forward
// Same pipeline, different work decomposition: the (tile pair, component) space is linearised (tile pair major)
// and cut into equal chunks, one per CTA pair; with at most one CTA pair per two SMs every SM streams the same
// number of components whatever the patch count is (16 129 patches = 127 tiles leave 21 of 148 SMs idle in the
// one-tile-per-CTA kernel; a row-block shard of 1/8 of a 1024^2 image only fills 64).  A chunk covers one or more
// SEGMENTS (a contiguous component range of one tile pair).  Per segment: gather A -> TMEM, run the pipeline,
// merge the two epilogue groups.  A segment that covers all K components writes value/argmax directly; a partial
// one stores (max, sum-exp, argmax) per patch row in a workspace slot and bumps the tile's arrival counter; the
// last segment of a tile to arrive merges the slots in component order (deterministic) and resets the counter.
// Pipeline indices (smem stage, TMEM slot, mbarrier phase) run on the CTA's global position counter, so the fixed
// barrier-ownership rule of the kernel above carries over unchanged.  Named barriers: 1 = A operand ready
// (gather warps + MMA warps), 2/3 = epilogue-group merge hand-shake, 4 = gather group only.
// Components of a segment are visited in a cyclic order that starts at a different offset in every CTA pair
// (max / logsumexp do not depend on the order): at any instant the SMs stream different B images, which
// spreads the L2 reads over the slices.
__device__ __forceinline__ int seg_rotation(int cl, int len, unsigned mul) { return (int)(((unsigned)cl * mul) % (unsigned)len); }

template <bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tc_sk_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                     const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck, int K,
                     int marginalize, int chunk, int smax, unsigned rot_mul, unsigned* __restrict__ counters, float* __restrict__ ws_m,
                     float* __restrict__ ws_s, int* __restrict__ ws_k, float* __restrict__ value,
                     int32_t* __restrict__ argmax, float* __restrict__ logp, double* __restrict__ sum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                          // NSTAGE x 32 KB
  float* sMW = reinterpret_cast<float*>(sB + NSTAGE * B_BYTES);  // NSLOT x 64 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES + NSLOT * MW_BYTES);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3 * NSLOT);
  int* s_flag = reinterpret_cast<int*>(s_tmem + 2);
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);       // 128 ints
  double* s_red = reinterpret_cast<double*>(s_valid + TM);  // 4 doubles (unused here, keeps the layout)
  float* s_mm = reinterpret_cast<float*>(s_red + 4);        // merge buffers of epilogue group B: max,
  float* s_ms = s_mm + TM;                                  //   sum-exp,
  int* s_mk = reinterpret_cast<int*>(s_ms + TM);            //   argmax

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
  const bool trim8 = ((marginalize >> 8) & 8) != 0;
  marginalize &= 1;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NSLOT + s); };
  auto mwfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + s); };
  const uint32_t crank = cluster_ctarank();

  // this CTA pair's chunk of the linearised (tile pair, component) space
  const int cl = blockIdx.x / CLUSTER;
  const int n_tiles = (g.P + TM - 1) / TM, n_pairs = (n_tiles + CLUSTER - 1) / CLUSTER;
  const long long w_tot = (long long)n_pairs * K;
  const long long lin_begin = (long long)cl * chunk;
  const long long lin_end = lin_begin + chunk < w_tot ? lin_begin + chunk : w_tot;
  const int npos = (int)(lin_end - lin_begin);            // > 0: the host launches ceil(w_tot / chunk) pairs
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

  if (warp < NPROD) {
    // ===================== bulk-TMA producers ======================================================
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      const long long comp0 = (long long)tp * K;
      const int pos_lo = (int)((lin_begin > comp0 ? lin_begin : comp0) - lin_begin);
      const int pos_hi = (int)((lin_end < comp0 + K ? lin_end : comp0 + K) - lin_begin);
      const int ka = (int)(lin_begin + pos_lo - comp0), len = pos_hi - pos_lo, rot = seg_rotation(cl, len, rot_mul);
      int pos = pos_lo + ((pos_lo % NPROD) == warp ? 0 : (warp - (pos_lo % NPROD) + NPROD) % NPROD);
      for (; pos < pos_hi; pos += NPROD) {
        int idx = pos - pos_lo + rot;
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int s = pos % NSTAGE, t = pos % NSLOT;
        mbar_wait(empty_bar(s), ((pos / NSTAGE) & 1) ^ 1);
        if (!ZERO_MEAN) mbar_wait(tempty_bar(t), ((pos / NSLOT) & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(s), B_BYTES);
          bulk_g2s_mc(smem_u32(sB + s * B_BYTES) + crank * (B_BYTES / CLUSTER),
                      Bt + (size_t)kc * B_BYTES + crank * (B_BYTES / CLUSTER), B_BYTES / CLUSTER, full_bar(s),
                      (uint16_t)((1u << CLUSTER) - 1));
          if (!ZERO_MEAN) {
            mbar_arrive_expect_tx(mwfull_bar(t), MW_BYTES);
            bulk_g2s(smem_u32(sMW + t * 64), mw + (size_t)kc * 64, MW_BYTES, mwfull_bar(t));
          }
        }
        __syncwarp();
      }
    }
  } else if (warp < M0 + NMMA) {
    // ===================== MMA issuers =============================================================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    int seg = tp_first - 1;  // last tile pair whose A operand this warp has been told about
    for (int pos = warp - M0; pos < npos; pos += NMMA) {
      const int tp = (int)((lin_begin + pos) / K);
      while (seg < tp) {  // every MMA warp takes part in every segment's hand-over, in order
        asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");
        ++seg;
      }
      const int s = pos % NSTAGE, t = pos % NSLOT;
      mbar_wait(tempty_bar(t), ((pos / NSLOT) & 1) ^ 1);
      mbar_wait(full_bar(s), (pos / NSTAGE) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_hi = sB_lo0 + s * (B_BYTES >> 4), b_lo = b_hi + ((2 * KBLOCK_BYTES_B) >> 4);
        const uint32_t d = tmem_u + A_COLS + t * SLOT_COLS;
        uint32_t acc = 0;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {  // small terms first: lo.hi, hi.lo, then hi.hi
          const uint32_t a_col = pass == 0 ? 64u : 0u;
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t n0 = TRI ? (trim8 ? 8u * kk : 16u * (kk >> 1)) : 0u;
            const uint32_t off16 = ((kk >> 2) * KBLOCK_BYTES_B + (kk & 3) * 32 + n0 * 128) >> 4;
            umma_tf32_ts(d + n0, tmem_u + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_n(64 - n0), acc);
            acc = 1;
          }
        }
        umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));
        umma_commit(tfull_bar(t));
      }
      __syncwarp();
    }
    while (seg < tp_last) {
      asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");
      ++seg;
    }
  } else {
    // ===================== gather + epilogue: group A (first 4 warps) even positions, group B odd ==
    const int grp = warp >= E0 + 4 ? 1 : 0;
    for (int tp = tp_first; tp <= tp_last; ++tp) {
      const long long comp0 = (long long)tp * K;
      const int pos_lo = (int)((lin_begin > comp0 ? lin_begin : comp0) - lin_begin);
      const int pos_hi = (int)((lin_end < comp0 + K ? lin_end : comp0 + K) - lin_begin);
      const int tile = tp * CLUSTER + (int)crank;
      const int64_t p = (int64_t)tile * TM + row;
      if (grp == 0) {
        // ---- gather: thread = patch row; 64 loads, mean, hi/lo split, tcgen05.st into TMEM lane `row`
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
        const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float hi[32], lo[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = ok ? vals[h * 32 + i] - mean : 0.f;
            hi[i] = tf32_rna(x);
            lo[i] = tf32_rna(x - hi[i]);
          }
          tmem_st32(a_lane + h * 32, hi);
          tmem_st32(a_lane + 64 + h * 32, lo);
        }
        tmem_st_wait();
        s_valid[row] = ok ? 1 : 0;
        tc_fence_before();
        asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");
      }

      float run_m = -CUDART_INF_F, run_s = 0.f;
      int run_k = 0x7fffffff;
      int pos = pos_lo + ((pos_lo ^ grp) & 1);  // first position of this group's parity in the segment
      const int ka = (int)(lin_begin + pos_lo - comp0), len = pos_hi - pos_lo, rot = seg_rotation(cl, len, rot_mul);
      for (; pos < pos_hi; pos += 2) {
        int idx = pos - pos_lo + rot;  // same rotated component order as the producers
        idx = idx >= len ? idx - len : idx;
        const int kc = ka + idx;
        const int t = pos % NSLOT;
        const float c_k = __ldg(ck + kc);
        if (!ZERO_MEAN) mbar_wait(mwfull_bar(t), (pos / NSLOT) & 1);
        mbar_wait(tfull_bar(t), (pos / NSLOT) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + A_COLS + t * SLOT_COLS;
        float y0[32], y1[32];
        tmem_ld32(taddr, y0);
        tmem_ld32(taddr + 32, y1);
        tmem_ld_wait();
        float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
        if (ZERO_MEAN) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            qa = fmaf(y0[i], y0[i], qa);
            qb = fmaf(y1[i], y1[i], qb);
            qc = fmaf(y0[i + 1], y0[i + 1], qc);
            qd = fmaf(y1[i + 1], y1[i + 1], qd);
          }
        } else {
          const float4* mwk = reinterpret_cast<const float4*>(sMW + t * 64);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 b0 = mwk[c4], b1 = mwk[8 + c4];
            float d0 = y0[4 * c4] - b0.x, d1 = y0[4 * c4 + 1] - b0.y, d2 = y0[4 * c4 + 2] - b0.z, d3 = y0[4 * c4 + 3] - b0.w;
            float e0 = y1[4 * c4] - b1.x, e1 = y1[4 * c4 + 1] - b1.y, e2 = y1[4 * c4 + 2] - b1.z, e3 = y1[4 * c4 + 3] - b1.w;
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

      // ---- segment end: group B -> group A through shared memory; both groups are past their last
      // accumulator, i.e. every MMA of the segment has completed and the A operand may be overwritten
      if (grp == 1) {
        s_mm[row] = run_m;
        s_ms[row] = run_s;
        s_mk[row] = run_k;
        asm volatile("bar.sync 2, 256;" ::: "memory");
        asm volatile("bar.sync 3, 256;" ::: "memory");  // group A has read the buffers
        continue;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
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
      bool final_here = (pos_hi - pos_lo) == K;
      if (!final_here) {
        const int c_first = (int)(comp0 / chunk), c_last = (int)((comp0 + K - 1) / chunk);
        const int nseg = c_last - c_first + 1;
        const size_t base = ((size_t)tile * smax) * TM + row;
        const size_t idx = base + (size_t)(cl - c_first) * TM;
        ws_m[idx] = run_m;
        ws_s[idx] = run_s;
        ws_k[idx] = run_k;
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
      }
      if (final_here) {
        double part = 0.0;
        if (p < g.P) {
          const bool ok = s_valid[row] != 0;
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

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until here
  if (warp == M0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- logsumexp backward on the tensor cores
// G[p,:] = scale * sum_k r[p,k] (xc_p Lam_k - bk_k) - row mean,  r = exp(logp[p,k] - lse[p]).
// Same pipeline as the forward kernel with B = Lam_k (dense schedule: Lam is symmetric, not triangular) and
// bk staged per slot; the epilogue scales each accumulator row by its responsibility and adds it to a
// 64-register gradient row per thread; the two epilogue groups merge through shared memory.
// logpT is component-major (K x P'), as the tensor-core forward writes it.
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_bwd_lse_tc_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                      const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ logpT,
                      const float* __restrict__ lse, int K, float scale, float* __restrict__ G) {
  constexpr bool trim8 = false;
  constexpr bool TRI = false, ZERO_MEAN = false;
  constexpr int NSTAGE = NSTAGE_BWD;
  const int marginalize = 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                          // NSTAGE x 32 KB
  float* sMW = reinterpret_cast<float*>(sB + NSTAGE * B_BYTES);  // NSLOT x 64 floats
  float* s_merge = reinterpret_cast<float*>(sB + NSTAGE * B_BYTES + NSLOT * MW_BYTES);  // 64 x 128 floats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES + NSLOT * MW_BYTES + MERGE_BYTES);
  // barrier indices: full[NSTAGE], empty[NSTAGE], tfull[NSLOT], tempty[NSLOT], mwfull[NSLOT]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 3 * NSLOT);
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);       // 128 ints
  double* s_red = reinterpret_cast<double*>(s_valid + TM);  // 4 doubles
  float* s_mm = reinterpret_cast<float*>(s_red + 4);        // merge buffers of epilogue group B: max,
  float* s_ms = s_mm + TM;                                  //   sum-exp,
  int* s_mk = reinterpret_cast<int*>(s_ms + TM);            //   argmax

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
  const int dbg = 0;
  (void)marginalize;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NSLOT + s); };
  auto mwfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + 2 * NSLOT + s); };
  // Every CTA walks the components in a different cyclic order (start k0): at any instant the
  // CTAs stream different B images / mw rows, which spreads the L2 reads over the slices instead
  // of 148 SMs hammering the same lines in lockstep.  max / logsumexp do not depend on the order
  // (ties in max resolve to the lowest component index, as torch.max does).
  // The two CTAs of a cluster share every B image (each loads one half and multicasts it to both),
  // which halves the L2 -> SM traffic; they therefore walk the components in the same order.
  const uint32_t crank = cluster_ctarank();
  const int k0 = (int)(((long long)(blockIdx.x / CLUSTER) * K) / (gridDim.x / CLUSTER));

  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int64_t p0 = (int64_t)blockIdx.x * TM;

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
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == M0) tmem_alloc(smem_u32(s_tmem), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // TMEM lane quarter of an epilogue/gather warp, and the patch row (= TMEM lane) of its threads
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int64_t p = p0 + row;

  if (warp < NPROD) {
    // ===================== bulk-TMA producers (whole warp waits, one elected lane issues) ==========
    int kc = k0 + warp;  // component handled at position k
    kc = kc >= K ? kc - K : kc;
    for (int k = warp; k < K; k += NPROD) {
      const int s = k % NSTAGE, t = k % NSLOT;
      mbar_wait(empty_bar(s), ((k / NSTAGE) & 1) ^ 1);
      // mw_k rides with accumulator slot t: free once the epilogue of component k - NSLOT is done
      if (!ZERO_MEAN) mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
      if (elect_one()) {
        // this CTA fetches half `crank` (hi or lo, 16 KB) of the image for both CTAs of the pair
        mbar_arrive_expect_tx(full_bar(s), B_BYTES);
        bulk_g2s_mc(smem_u32(sB + s * B_BYTES) + crank * (B_BYTES / CLUSTER),
                    Bt + (size_t)kc * B_BYTES + crank * (B_BYTES / CLUSTER), B_BYTES / CLUSTER, full_bar(s),
                    (uint16_t)((1u << CLUSTER) - 1));
        if (!ZERO_MEAN) {
          mbar_arrive_expect_tx(mwfull_bar(t), MW_BYTES);
          bulk_g2s(smem_u32(sMW + t * 64), mw + (size_t)kc * 64, MW_BYTES, mwfull_bar(t));
        }
      }
      __syncwarp();
      kc += NPROD;
      kc = kc >= K ? kc - K : kc;
    }
  } else if (warp >= E0 && warp < E0 + 4) {
    // ---- gather (epilogue group A): thread = patch row; 64 loads, mean, hi/lo split, tcgen05.st into TMEM lane `row`
    float vals[64];
    float s = 0.f;
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
          s += x;
          ok = ok && (x > -1e5f);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 64; ++i) vals[i] = 0.f;
    }
    const float mean = s * (1.f / 64.f);
    const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float hi[32], lo[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float x = ok ? vals[h * 32 + i] - mean : 0.f;
        hi[i] = tf32_rna(x);
        lo[i] = tf32_rna(x - hi[i]);
      }
      tmem_st32(a_lane + h * 32, hi);
      tmem_st32(a_lane + 64 + h * 32, lo);
    }
    tmem_st_wait();
    s_valid[row] = ok ? 1 : 0;
    tc_fence_before();
    // the 4 gather warps -> MMA warp: named barrier 1 (128 gather threads + 32 MMA-warp threads)
    asm volatile("bar.arrive 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");
  }

  if (warp >= M0 && warp < M0 + NMMA) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =========
    asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA) : "memory");  // A operand is in TMEM
    tc_fence_after();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    for (int k = warp - M0; k < K; k += NMMA) {
      const int s = k % NSTAGE, t = k % NSLOT;
      mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
      mbar_wait(full_bar(s), (k / NSTAGE) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_hi = sB_lo0 + s * (B_BYTES >> 4), b_lo = b_hi + ((2 * KBLOCK_BYTES_B) >> 4);
        const uint32_t d = tmem_u + A_COLS + t * SLOT_COLS;
        uint32_t acc = 0;
        // small terms first: lo.hi, hi.lo, then hi.hi
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_col = pass == 0 ? 64u : 0u;
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if ((dbg & 2) && (pass > 0 || kk > 0)) continue;
            // upper-triangular Lw: input features [8kk, 8kk+8) only reach whitened features >= 8kk
            const uint32_t n0 = TRI ? (trim8 ? 8u * kk : 16u * (kk >> 1)) : 0u;
            const uint32_t off16 = ((kk >> 2) * KBLOCK_BYTES_B + (kk & 3) * 32 + n0 * 128) >> 4;
            umma_tf32_ts(d + n0, tmem_u + a_col + kk * 8, desc_from_lo(b_base + off16), idesc_n(64 - n0), acc);
            acc = 1;
          }
        }
        umma_commit_mc(empty_bar(s), (uint16_t)((1u << CLUSTER) - 1));  // stage free in both CTAs of the pair
        umma_commit(tfull_bar(t));  // accumulator slot complete
      }
      __syncwarp();
    }
  } else if (warp >= E0) {
    // ===================== epilogue: group A (warps 2-5) takes even positions, group B odd =========
    const int grp = warp >= E0 + 4 ? 1 : 0;
    float gacc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) gacc[i] = 0.f;
    const float lse_p = p < g.P ? __ldg(lse + p) : 0.f;
    int kc = k0 + grp;
    kc = kc >= K ? kc - K : kc;
    for (int k = grp; k < K; k += 2) {
      const int t = k % NSLOT;
      const float r = p < g.P ? expf(__ldg(logpT + (size_t)kc * g.P + p) - lse_p) : 0.f;
      mbar_wait(mwfull_bar(t), (k / NSLOT) & 1);
      mbar_wait(tfull_bar(t), (k / NSLOT) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + A_COLS + t * SLOT_COLS;
      const float4* bkk = reinterpret_cast<const float4*>(sMW + t * 64);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float y[32];
        tmem_ld32(taddr + 32 * h, y);
        tmem_ld_wait();
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b = bkk[8 * h + c4];
          gacc[32 * h + 4 * c4 + 0] = fmaf(r, y[4 * c4 + 0] - b.x, gacc[32 * h + 4 * c4 + 0]);
          gacc[32 * h + 4 * c4 + 1] = fmaf(r, y[4 * c4 + 1] - b.y, gacc[32 * h + 4 * c4 + 1]);
          gacc[32 * h + 4 * c4 + 2] = fmaf(r, y[4 * c4 + 2] - b.z, gacc[32 * h + 4 * c4 + 2]);
          gacc[32 * h + 4 * c4 + 3] = fmaf(r, y[4 * c4 + 3] - b.w, gacc[32 * h + 4 * c4 + 3]);
        }
      }
      fence_cta();  // the bk row reads above are performed before the slot is released (see mbar_arrive_after)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_after(tempty_bar(t), gacc[31] + gacc[63], rt_zero);
      kc += 2;
      kc = kc >= K ? kc - K : kc;
    }
    // merge group B into group A (feature-major buffer: conflict-free), then mean-subtract, scale, store
    if (grp == 1) {
#pragma unroll
      for (int j = 0; j < 64; ++j) s_merge[j * TM + row] = gacc[j];
      asm volatile("bar.arrive 2, 256;" ::: "memory");
    } else {
      asm volatile("bar.sync 2, 256;" ::: "memory");
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        gacc[j] += s_merge[j * TM + row];
        rs += gacc[j];
      }
      const float mean = rs * (1.f / 64.f);
      if (p < g.P) {
        const bool ok = s_valid[row] != 0;
        float4* out = reinterpret_cast<float4*>(G + (size_t)p * 64);
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          float4 o;
          o.x = ok ? scale * (gacc[4 * c4 + 0] - mean) : 0.f;
          o.y = ok ? scale * (gacc[4 * c4 + 1] - mean) : 0.f;
          o.z = ok ? scale * (gacc[4 * c4 + 2] - mean) : 0.f;
          o.w = ok ? scale * (gacc[4 * c4 + 3] - mean) : 0.f;
          out[c4] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast into / arrive on this CTA's shared memory until here
  if (warp == M0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace jd

using namespace jd;

extern "C" {

// JD_TC_TRIM8=1: per-k-step triangular trim (MMA N = 64, 56, .., 8; not a multiple of 16 for every step)
static int tc_trim8() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JD_TC_TRIM8");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v;
}

size_t jd_gmm_tc_packed_bytes(int K) { return (size_t)K * tc::B_BYTES; }

int jd_gmm_tc_pack(const float* Lw, int K, void* Bt, jd_stream_t stream) {
  JD_CHECK_ARG(Lw && Bt && K > 0, "jd_gmm_tc_pack: bad arguments");
  int64_t n = (int64_t)K * 4096;
  tc::pack_b_kernel<<<(int)((n + 255) / 256), 256, 0, to_stream(stream)>>>(Lw, K, reinterpret_cast<uint8_t*>(Bt));
  JD_CHECK_LAUNCH("jd_gmm_tc_pack");
  return JD_OK;
}

int jd_gmm_prior_forward_tc(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                            int row_end, const void* Bt, const float* mw, const float* ck, int K, int upper_tri,
                            int zero_mean, int marginalize, float* value, int32_t* argmax, float* logp, double* sum,
                            jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && mw && ck && K > 0, "jd_gmm_prior_forward_tc: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tc: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tc: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0,
               "jd_gmm_prior_forward_tc: Bt and mw must be 16-byte aligned");
  tc::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaSuccess;
    const void* kerns[4] = {(const void*)tc::gmm_fwd_tc_kernel<false, false>, (const void*)tc::gmm_fwd_tc_kernel<false, true>,
                            (const void*)tc::gmm_fwd_tc_kernel<true, false>, (const void*)tc::gmm_fwd_tc_kernel<true, true>};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tc: cannot reserve %zu B of shared memory: %s", tc::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
  }
  int grid = (g.P + tc::TM - 1) / tc::TM;
  grid = (grid + tc::CLUSTER - 1) / tc::CLUSTER * tc::CLUSTER;  // whole clusters; surplus CTAs own no patch
  auto kern = upper_tri ? (zero_mean ? tc::gmm_fwd_tc_kernel<true, true> : tc::gmm_fwd_tc_kernel<true, false>)
                        : (zero_mean ? tc::gmm_fwd_tc_kernel<false, true> : tc::gmm_fwd_tc_kernel<false, false>);
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("JD_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  if (dbg & 4) kern = zero_mean ? tc::gmm_fwd_tc_kernel<false, true> : tc::gmm_fwd_tc_kernel<false, false>;
  marginalize = (marginalize ? 1 : 0) | (dbg << 8) | (tc_trim8() << 11);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = tc::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tc::CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t* bt8 = reinterpret_cast<const uint8_t*>(Bt);
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, flux, g, shift_yx, bt8, mw, ck, K, marginalize, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_forward_tc: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tc");
  return JD_OK;
}

// ---- stream-K forward: plan (CTA pairs, chunk, workspace) + launch
namespace jd {
namespace tc {
struct SkPlan {
  int n_clusters, chunk, smax, n_tiles2;  // n_tiles2 = tiles rounded up to whole CTA pairs
  size_t off_m, off_s, off_k, bytes;
};
static int sk_max_clusters() {
  static int n = -1;
  if (n < 0) {
    const char* e = getenv("JD_TC_SK_CLUSTERS");  // tuning knob
    n = e ? atoi(e) : num_sms() / CLUSTER;         // one CTA per SM (launch bounds), two SMs per pair
    if (n < 1) n = 1;
  }
  return n;
}
static SkPlan sk_plan(int64_t P, int K) {
  SkPlan p;
  const int64_t n_tiles = (P + TM - 1) / TM, n_pairs = (n_tiles + CLUSTER - 1) / CLUSTER;
  const int64_t w_tot = n_pairs * K;
  const int C = sk_max_clusters();
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
}  // namespace tc
}  // namespace jd

int64_t jd_gmm_tc_sk_workspace_bytes(int64_t P, int K) {
  if (P <= 0 || K <= 0) return 0;
  return (int64_t)tc::sk_plan(P, K).bytes;
}

int jd_gmm_tc_sk_plan(int64_t P, int K, int* n_cta_pairs, int* chunk, int* max_segments_per_tile) {
  JD_CHECK_ARG(P > 0 && K > 0 && n_cta_pairs && chunk && max_segments_per_tile, "jd_gmm_tc_sk_plan: bad arguments");
  const tc::SkPlan p = tc::sk_plan(P, K);
  *n_cta_pairs = p.n_clusters;
  *chunk = p.chunk;
  *max_segments_per_tile = p.smax;
  return JD_OK;
}

int jd_gmm_prior_forward_tc_sk(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                               int row_end, const void* Bt, const float* mw, const float* ck, int K, int upper_tri,
                               int zero_mean, int marginalize, void* workspace, float* value, int32_t* argmax,
                               float* logp, double* sum, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && mw && ck && workspace && K > 0, "jd_gmm_prior_forward_tc_sk: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tc_sk: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tc_sk: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "jd_gmm_prior_forward_tc_sk: Bt and mw must be 16-byte aligned, the workspace 256-byte aligned");
  tc::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaSuccess;
    const void* kerns[4] = {(const void*)tc::gmm_fwd_tc_sk_kernel<false, false>, (const void*)tc::gmm_fwd_tc_sk_kernel<false, true>,
                            (const void*)tc::gmm_fwd_tc_sk_kernel<true, false>, (const void*)tc::gmm_fwd_tc_sk_kernel<true, true>};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tc_sk: cannot reserve %zu B of shared memory: %s", tc::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
  }
  const tc::SkPlan p = tc::sk_plan(g.P, K);
  static int rot_env = -1;
  if (rot_env < 0) {
    const char* e = getenv("JD_TC_SK_ROT");  // 0: visit the components of a segment in ascending order
    rot_env = e ? atoi(e) : 40503;
  }
  const unsigned rot_mul = (unsigned)rot_env;
  auto kern = upper_tri ? (zero_mean ? tc::gmm_fwd_tc_sk_kernel<true, true> : tc::gmm_fwd_tc_sk_kernel<true, false>)
                        : (zero_mean ? tc::gmm_fwd_tc_sk_kernel<false, true> : tc::gmm_fwd_tc_sk_kernel<false, false>);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_clusters * tc::CLUSTER);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = tc::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tc::CLUSTER;
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
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, flux, g, shift_yx, bt8, mw, ck, K, (marginalize ? 1 : 0) | (tc_trim8() << 11), p.chunk,
                                      p.smax, rot_mul, counters, ws_m, ws_s, ws_k, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_forward_tc_sk: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tc_sk");
  return JD_OK;
}

int jd_gmm_prior_backward_lse_tc(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                                 int row_end, const void* Bt_lam, const float* bk, int K, const float* logpT,
                                 const float* lse, float scale, float* G, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt_lam && bk && logpT && lse && G && K > 0, "jd_gmm_prior_backward_lse_tc: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_backward_lse_tc: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_backward_lse_tc: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt_lam) & 15) == 0 && (reinterpret_cast<uintptr_t>(bk) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(G) & 15) == 0,
               "jd_gmm_prior_backward_lse_tc: Bt_lam, bk and G must be 16-byte aligned");
  tcx::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(tc::gmm_bwd_lse_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc::SMEM_BYTES_BWD);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_backward_lse_tc: cannot reserve %zu B of shared memory: %s", tc::SMEM_BYTES_BWD,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
  }
  int grid = (g.P + tc::TM - 1) / tc::TM;
  grid = (grid + tc::CLUSTER - 1) / tc::CLUSTER * tc::CLUSTER;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc::NTHREADS);
  cfg.dynamicSmemBytes = tc::SMEM_BYTES_BWD;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tc::CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t* bt8 = reinterpret_cast<const uint8_t*>(Bt_lam);
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc::gmm_bwd_lse_tc_kernel, flux, g, shift_yx, bt8, bk, logpT, lse, K, scale, G);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_backward_lse_tc: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_backward_lse_tc");
  return JD_OK;
}

int jd_gmm_prior_forward(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                         int row_end, const float* Lw, const float* mw, const float* ck, int K, int marginalize,
                         float* value, int32_t* argmax, float* logp, double* sum, int backend, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Lw && mw && ck && K > 0, "jd_gmm_prior_forward: null pointer");
  if (backend == 0)
    return gmm_prior_forward_simt(flux, fH, fW, shift_yx, stride, row_begin, row_end, Lw, mw, ck, K, marginalize,
                                  value, argmax, logp, sum, to_stream(stream));
  set_error("jd_gmm_prior_forward: backend %d takes the packed operand: call jd_gmm_prior_forward_tc", backend);
  return JD_ERR_UNSUPPORTED;
}

}  // extern "C"
