// GMM patch prior forward on tcgen05 with SPLIT-FP16 operands (kind::f16, FP32 accumulation in TMEM).
//
// Same kernel structure, pipeline and epilogue as jd_gmm_tc.cu (split-TF32); only the operand encoding
// differs.  x = s^-1 (hi + lo) with hi, lo FP16 and s a power-of-two scale (per patch row for A, per
// component for B) chosen so that max |s x| lies in [2^13, 2^14): the pair carries 22 significand bits, the
// same as a TF32 hi/lo pair, elements far below the row maximum keep an absolute error < 2^-38 max.  The three
// products lo.hi + hi.lo + hi.hi then run at the FP16 tensor rate (2x TF32) on operands half the size:
// 16 KB of B per component instead of 32 KB through the TMA ring and the MMA's shared-memory reads, which is
// what bounds the TF32 kernel.  The epilogue undoes the scales (one multiply per value,
// fused into the mean subtraction).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdlib.h>

#include "jd_common.cuh"
#include "jd_tc_ptx.cuh"

namespace jd {
namespace tc16 {

using namespace tcx;

constexpr int TM = 128;                 // patches per CTA
constexpr int NSTAGE = 8;               // B ring depth (multiple of NPROD)
constexpr int NSLOT = 6;                // TMEM accumulator slots (multiple of NMMA and of the 2 epilogue groups)
constexpr int SLOT_COLS = 64;
constexpr int A_COLS = 64;              // TMEM columns [0,32) = A hi (half2 packed), [32,64) = A lo
constexpr int TMEM_COLS = 512;          // A_COLS + NSLOT * SLOT_COLS
constexpr int MAT_BYTES = 64 * 128;     // one 64 x 64 FP16 matrix: 64 rows x 128 B, one swizzle atom wide
constexpr int B_BYTES = 2 * MAT_BYTES;  // hi, lo = 16 KB per component
constexpr int CLUSTER = 2;
constexpr int NMMA = 3;                 // MMA-issuing warps: component position k is issued by warp M0 + k % NMMA
constexpr int NPROD = 2;                // bulk-TMA producer warps: position k is loaded by warp k % NPROD
// Every mbarrier is waited on by ONE fixed warp (group) across its successive uses: a waiter may be at most one
// phase ahead of the barrier (parity aliasing otherwise), which in-order processing by a single owner guarantees.
static_assert(NSTAGE % NPROD == 0, "a smem stage must always be refilled by the same producer warp");
static_assert(NSLOT % NMMA == 0 && NSLOT % 2 == 0, "a TMEM slot must always belong to the same MMA warp / epilogue group");
constexpr int M0 = NPROD;                // first MMA-issuer warp (also owns the TMEM allocation)
constexpr int E0 = NPROD + NMMA;         // first epilogue warp
constexpr int NTHREADS = 32 * (NPROD + NMMA + 8);  // producers, MMA issuers, 2 x 4 epilogue warps
constexpr int MW_BYTES = 64 * 4;
constexpr size_t SMEM_BYTES = 1024 + NSTAGE * B_BYTES + NSLOT * MW_BYTES + 4096;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// D = F32, A = B = F16, both K-major, M = 128, N = n
__device__ __host__ constexpr uint32_t idesc16_n(uint32_t n) {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---------------------------------------------------------------- setup: pack Lw_k^T as scaled FP16 hi/lo
// One CTA per component.  Image (16 KB): hi then lo, row n = whitened feature j (128 B = 64 halfs over the input
// feature i), 128B-swizzled.  binv[k] = 1 / scale_k.
__global__ void pack_b16_kernel(const float* __restrict__ Lw, int K, uint8_t* __restrict__ out,
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
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const uint32_t off = (j >> 3) * 1024 + (j & 7) * 128 + (((i >> 3) ^ (j & 7)) << 4) + (i & 7) * 2;
    *reinterpret_cast<__half*>(base + off) = h;
    *reinterpret_cast<__half*>(base + MAT_BYTES + off) = l;
  }
}

// ---------------------------------------------------------------- the forward kernel
template <bool TRI, bool ZERO_MEAN>
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tc16_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                  const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck, const float* __restrict__ binv, int K,
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
  float* s_rinv = reinterpret_cast<float*>(s_mk + TM);      // 1 / (row scale) per patch row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt_zero = (uint32_t)K >> 30;  // 0 at run time (K < 2^30), opaque to the compiler
  const int dbg = marginalize >> 8;  // profiling knobs (JD_TC_DEBUG): 1 = no epilogue TMEM loads, 2 = one MMA per component
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
  float row_inv = 1.f;  // 1 / (row scale): written to s_rinv by the gather, read by both epilogue groups

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
        bulk_g2s_mc(smem_u32(sB + s * B_BYTES) + crank * ((B_BYTES / CLUSTER)),
                    Bt + (size_t)kc * B_BYTES + crank * ((B_BYTES / CLUSTER)), (B_BYTES / CLUSTER), full_bar(s),
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
    // per-row power-of-two scale: max |x| lands in [2^13, 2^14) so that the FP16 hi/lo pair keeps 22 bits
    float amax = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      vals[i] = ok ? vals[i] - mean : 0.f;
      amax = fmaxf(amax, fabsf(vals[i]));
    }
    int e = amax > 0.f ? ilogbf(amax) : 13;
    e = max(-100, min(100, e));
    const float sA = ldexpf(1.f, 13 - e);
    s_rinv[row] = ldexpf(1.f, e - 13);
    const uint32_t a_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    {
      float hi[32], lo[32];  // 32 packed half2 words each (64 features)
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x0 = vals[2 * c] * sA, x1 = vals[2 * c + 1] * sA;
        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
        const __half l0 = __float2half_rn(x0 - __half2float(h0)), l1 = __float2half_rn(x1 - __half2float(h1));
        hi[c] = __uint_as_float((uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16));
        lo[c] = __uint_as_float((uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16));
      }
      tmem_st32(a_lane, hi);
      tmem_st32(a_lane + 32, lo);
    }
    tmem_st_wait();
    s_valid[row] = ok ? 1 : 0;
    tc_fence_before();
    // the 4 gather warps -> MMA warp: named barrier 1 (128 gather threads + 32 MMA-warp threads)
    asm volatile("bar.arrive 1, %0;" ::"n"(128 + 32 * NMMA + 128) : "memory");
  }

  if (warp >= M0 && warp < M0 + NMMA) {
    // ===================== MMA issuers (warp-uniform control flow, one elected lane issues) =========
    asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA + 128) : "memory");  // A operand is in TMEM
    tc_fence_after();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sB_lo0 = desc_lo(smem_u32(sB));
    for (int k = warp - M0; k < K; k += NMMA) {
      const int s = k % NSTAGE, t = k % NSLOT;
      mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
      mbar_wait(full_bar(s), (k / NSTAGE) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t b_hi = sB_lo0 + s * (B_BYTES >> 4), b_lo = b_hi + (MAT_BYTES >> 4);
        const uint32_t d = tmem_u + A_COLS + t * SLOT_COLS;
        uint32_t acc = 0;
        // small terms first: lo.hi, hi.lo, then hi.hi
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_col = pass == 0 ? 32u : 0u;
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if ((dbg & 2) && (pass > 0 || kk > 0)) continue;
            // upper-triangular Lw: input features [16kk, 16kk+16) only reach whitened features >= 16kk
            const uint32_t n0 = TRI ? 16u * kk : 0u;
            const uint32_t off16 = (kk * 32 + n0 * 128) >> 4;
            umma_f16_ts(d + n0, tmem_u + a_col + kk * 8, desc_from_lo(b_base + off16), idesc16_n(64 - n0), acc);
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
    if (grp == 1) asm volatile("bar.sync 1, %0;" ::"n"(128 + 32 * NMMA + 128) : "memory");  // row scales written by the gather warps
    row_inv = s_rinv[row];
    float run_m = -CUDART_INF_F, run_s = 0.f;
    int run_k = 0x7fffffff;
    int kc = k0 + grp;
    kc = kc >= K ? kc - K : kc;
    for (int k = grp; k < K; k += 2) {
      const int t = k % NSLOT;
      const float c_k = __ldg(ck + kc);
      const float inv = row_inv * __ldg(binv + kc);  // undo the row and component scales
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
      const float qsum = ZERO_MEAN ? ((qa + qb) + (qc + qd)) * (inv * inv) : (qa + qb) + (qc + qd);
      const float lp = fmaf(-0.5f, qsum, c_k);
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

}  // namespace tc16
}  // namespace jd

using namespace jd;

extern "C" {

size_t jd_gmm_tc16_packed_bytes(int K) { return (size_t)K * tc16::B_BYTES; }

int jd_gmm_tc16_pack(const float* Lw, int K, void* Bt, float* binv, jd_stream_t stream) {
  JD_CHECK_ARG(Lw && Bt && binv && K > 0, "jd_gmm_tc16_pack: bad arguments");
  tc16::pack_b16_kernel<<<K, 256, 0, to_stream(stream)>>>(Lw, K, reinterpret_cast<uint8_t*>(Bt), binv);
  JD_CHECK_LAUNCH("jd_gmm_tc16_pack");
  return JD_OK;
}

int jd_gmm_prior_forward_tc16(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                              int row_end, const void* Bt, const float* binv, const float* mw, const float* ck, int K,
                              int upper_tri, int zero_mean, int marginalize, float* value, int32_t* argmax,
                              float* logp, double* sum, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && binv && mw && ck && K > 0, "jd_gmm_prior_forward_tc16: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tc16: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tc16: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0 && (reinterpret_cast<uintptr_t>(mw) & 15) == 0,
               "jd_gmm_prior_forward_tc16: Bt and mw must be 16-byte aligned");
  tcx::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set[64] = {};
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaSuccess;
    const void* kerns[4] = {(const void*)tc16::gmm_fwd_tc16_kernel<false, false>,
                            (const void*)tc16::gmm_fwd_tc16_kernel<false, true>,
                            (const void*)tc16::gmm_fwd_tc16_kernel<true, false>,
                            (const void*)tc16::gmm_fwd_tc16_kernel<true, true>};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
      e = cudaFuncSetAttribute(kerns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc16::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tc16: cannot reserve %zu B of shared memory: %s", tc16::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
  }
  int grid = (g.P + tc16::TM - 1) / tc16::TM;
  grid = (grid + tc16::CLUSTER - 1) / tc16::CLUSTER * tc16::CLUSTER;
  auto kern = upper_tri ? (zero_mean ? tc16::gmm_fwd_tc16_kernel<true, true> : tc16::gmm_fwd_tc16_kernel<true, false>)
                        : (zero_mean ? tc16::gmm_fwd_tc16_kernel<false, true> : tc16::gmm_fwd_tc16_kernel<false, false>);
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("JD_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  marginalize = (marginalize ? 1 : 0) | (dbg << 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(tc16::NTHREADS);
  cfg.dynamicSmemBytes = tc16::SMEM_BYTES;
  cfg.stream = to_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = tc16::CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const uint8_t* bt8 = reinterpret_cast<const uint8_t*>(Bt);
  cudaError_t le =
      cudaLaunchKernelEx(&cfg, kern, flux, g, shift_yx, bt8, mw, ck, binv, K, marginalize, value, argmax, logp, sum);
  if (le != cudaSuccess) {
    set_error("jd_gmm_prior_forward_tc16: launch failed: %s", cudaGetErrorString(le));
    cudaGetLastError();
    return JD_ERR_CUDA;
  }
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tc16");
  return JD_OK;
}

}  // extern "C"
