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
//   warp 1      TMEM allocator + MMA issuer: 24 tcgen05.mma (M128 N64 K8, kind::tf32) per component
//               into one of 8 accumulator slots (64 TMEM columns each);
//   warps 2..5  gather the 128 patches from the flux image at rolled coordinates, subtract the
//               patch mean, split hi/lo and write the A operand (swizzled K-major) to smem once;
//               then act as the epilogue: tcgen05.ld the 128x64 accumulator of each component
//               (thread = patch row), subtract mw_k, square, reduce over the 64 whitened features,
//               add ck_k and fold into a running max/argmax or online logsumexp.  Y never leaves
//               the SM; only value/argmax (and optionally logp) are written.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "jd_common.cuh"

namespace jd {

int gmm_prior_forward_simt(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* Lw, const float* mw, const float* ck, int K, int marginalize,
                           float* value, int32_t* argmax, float* logp, double* sum, cudaStream_t st);

namespace tc {

constexpr int TM = 128;                 // patches per CTA
constexpr int NSTAGE = 4;               // B ring depth
constexpr int NSLOT = 8;                // TMEM accumulator slots
constexpr int SLOT_COLS = 64;
constexpr int KBLOCK_BYTES_A = TM * 128;      // 128 rows x 128 B (32 tf32) = 16 KB
constexpr int KBLOCK_BYTES_B = 64 * 128;      // 64 rows x 128 B = 8 KB
constexpr int A_BYTES = 4 * KBLOCK_BYTES_A;   // hi(kb0,kb1) lo(kb0,kb1) = 64 KB
constexpr int B_BYTES = 4 * KBLOCK_BYTES_B;   // hi(kb0,kb1) lo(kb0,kb1) = 32 KB per component
constexpr int NTHREADS = 192;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + A_BYTES + NSTAGE * B_BYTES + 1024 /*barriers etc.*/;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 columns of 32-bit: thread i of the warp receives row (lane base + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major), [32,46) SBO >> 4
//   (8 rows x 128 B = 1024 B between 8-row groups), [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, M = 128, N = 64
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of element (row, d) inside a [rows x 64 tf32] operand stored as two 128B-swizzled k-blocks
__device__ __host__ __forceinline__ uint32_t sw128_offset(int row, int d, int kblock_bytes) {
  int kb = d >> 5, c = (d & 31) >> 2, e = d & 3;
  return kb * kblock_bytes + (row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4) + e * 4;
}

struct Geom {
  int fH, fW, sy, sx, stride, nx, row_begin, P;
};

__device__ __forceinline__ int src_row(const Geom& g, int iy, int u) { return wrap(iy * g.stride + u - g.sy, g.fH); }
__device__ __forceinline__ int src_col(const Geom& g, int ix, int v) { return wrap(ix * g.stride + v - g.sx, g.fW); }

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
__global__ void __launch_bounds__(NTHREADS, 1)
gmm_fwd_tc_kernel(const float* __restrict__ flux, Geom g, const int32_t* __restrict__ shift_yx,
                  const uint8_t* __restrict__ Bt, const float* __restrict__ mw, const float* __restrict__ ck, int K,
                  int marginalize, float* __restrict__ value, int32_t* __restrict__ argmax, float* __restrict__ logp,
                  double* __restrict__ sum) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                          // 64 KB
  uint8_t* sB = smem + A_BYTES;                // NSTAGE x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + NSTAGE * B_BYTES);
  // barrier indices: full[NSTAGE], empty[NSTAGE], tfull[NSLOT], tempty[NSLOT]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 2 * NSLOT);
  int* s_valid = reinterpret_cast<int*>(s_tmem + 4);       // 128 ints
  double* s_red = reinterpret_cast<double*>(s_valid + TM);  // 4 doubles

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * NSTAGE + NSLOT + s); };

  if (shift_yx) {
    g.sy = shift_yx[0];
    g.sx = shift_yx[1];
  }
  const int64_t p0 = (int64_t)blockIdx.x * TM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), NSLOT * SLOT_COLS);

  if (warp >= 2) {
    // ---- gather: thread = patch row; 64 loads, mean, hi/lo split, swizzled stores
    const int row = threadIdx.x - 64;
    const int64_t p = p0 + row;
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
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      float4 hi, lo;
      float x0 = ok ? vals[4 * c + 0] - mean : 0.f, x1 = ok ? vals[4 * c + 1] - mean : 0.f;
      float x2 = ok ? vals[4 * c + 2] - mean : 0.f, x3 = ok ? vals[4 * c + 3] - mean : 0.f;
      hi.x = tf32_rna(x0), hi.y = tf32_rna(x1), hi.z = tf32_rna(x2), hi.w = tf32_rna(x3);
      lo.x = tf32_rna(x0 - hi.x), lo.y = tf32_rna(x1 - hi.y), lo.z = tf32_rna(x2 - hi.z), lo.w = tf32_rna(x3 - hi.w);
      uint32_t off = sw128_offset(row, 4 * c, KBLOCK_BYTES_A);
      *reinterpret_cast<float4*>(sA + off) = hi;
      *reinterpret_cast<float4*>(sA + 2 * KBLOCK_BYTES_A + off) = lo;
    }
    s_valid[row] = ok ? 1 : 0;
    // make the generic-proxy writes of A visible to the async proxy (UMMA operand reads)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== bulk-TMA producer =====================
    if (lane == 0) {
      for (int k = 0; k < K; ++k) {
        const int s = k % NSTAGE;
        const uint32_t ph = (k / NSTAGE) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), B_BYTES);
        bulk_g2s(smem_u32(sB + s * B_BYTES), Bt + (size_t)k * B_BYTES, B_BYTES, full_bar(s));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t a_hi = smem_u32(sA), a_lo = a_hi + 2 * KBLOCK_BYTES_A;
      for (int k = 0; k < K; ++k) {
        const int s = k % NSTAGE, t = k % NSLOT;
        mbar_wait(tempty_bar(t), ((k / NSLOT) & 1) ^ 1);
        mbar_wait(full_bar(s), (k / NSTAGE) & 1);
        tc_fence_after();
        const uint32_t b_hi = smem_u32(sB + s * B_BYTES), b_lo = b_hi + 2 * KBLOCK_BYTES_B;
        const uint32_t d = tmem_base + t * SLOT_COLS;
        uint32_t acc = 0;
        // small terms first: lo.hi, hi.lo, then hi.hi
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a_base = pass == 0 ? a_lo : a_hi;
          const uint32_t b_base = pass == 1 ? b_lo : b_hi;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t a_addr = a_base + (kk >> 2) * KBLOCK_BYTES_A + (kk & 3) * 32;
            const uint32_t b_addr = b_base + (kk >> 2) * KBLOCK_BYTES_B + (kk & 3) * 32;
            umma_tf32(d, make_desc(a_addr), make_desc(b_addr), IDESC, acc);
            acc = 1;
          }
        }
        umma_commit(empty_bar(s));  // smem stage reusable once these MMAs have read it
        umma_commit(tfull_bar(t));  // accumulator slot complete
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int64_t p = p0 + row;
    float run_m = -CUDART_INF_F, run_s = 0.f;
    int run_k = 0;
    for (int k = 0; k < K; ++k) {
      const int t = k % NSLOT;
      mbar_wait(tfull_bar(t), (k / NSLOT) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + t * SLOT_COLS;
      const float4* mwk = reinterpret_cast<const float4*>(mw + (size_t)k * 64);
      float qv = 0.f;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float y[32];
        tmem_ld32(taddr + h * 32, y);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 b = __ldg(mwk + h * 8 + c4);
          float d0 = y[4 * c4] - b.x, d1 = y[4 * c4 + 1] - b.y, d2 = y[4 * c4 + 2] - b.z, d3 = y[4 * c4 + 3] - b.w;
          qv = fmaf(d0, d0, qv);
          qv = fmaf(d1, d1, qv);
          qv = fmaf(d2, d2, qv);
          qv = fmaf(d3, d3, qv);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(t));
      const float lp = fmaf(-0.5f, qv, __ldg(ck + k));
      if (logp && p < g.P) logp[p * K + k] = lp;
      if (marginalize) {
        if (lp > run_m) {
          run_s = run_s * expf(run_m - lp) + 1.f;
          run_m = lp;
          run_k = k;
        } else {
          run_s += expf(lp - run_m);
        }
      } else if (lp > run_m) {
        run_m = lp;
        run_k = k;
      }
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

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && sum) atomicAdd(sum, s_red[0] + s_red[1] + s_red[2] + s_red[3]);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, NSLOT * SLOT_COLS);
  }
}

}  // namespace tc
}  // namespace jd

using namespace jd;

extern "C" {

size_t jd_gmm_tc_packed_bytes(int K) { return (size_t)K * tc::B_BYTES; }

int jd_gmm_tc_pack(const float* Lw, int K, void* Bt, jd_stream_t stream) {
  JD_CHECK_ARG(Lw && Bt && K > 0, "jd_gmm_tc_pack: bad arguments");
  int64_t n = (int64_t)K * 4096;
  tc::pack_b_kernel<<<(int)((n + 255) / 256), 256, 0, to_stream(stream)>>>(Lw, K, reinterpret_cast<uint8_t*>(Bt));
  JD_CHECK_LAUNCH("jd_gmm_tc_pack");
  return JD_OK;
}

int jd_gmm_prior_forward_tc(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                            int row_end, const void* Bt, const float* mw, const float* ck, int K, int marginalize,
                            float* value, int32_t* argmax, float* logp, double* sum, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Bt && mw && ck && K > 0, "jd_gmm_prior_forward_tc: null pointer");
  JD_CHECK_ARG(fH >= PATCH && fW >= PATCH && stride >= 1 && stride <= PATCH, "jd_gmm_prior_forward_tc: bad geometry");
  int ny = (fH - PATCH) / stride + 1, nx = (fW - PATCH) / stride + 1;
  JD_CHECK_ARG(row_begin >= 0 && row_end <= ny && row_begin < row_end,
               "jd_gmm_prior_forward_tc: bad patch-row block [%d,%d) of %d", row_begin, row_end, ny);
  JD_CHECK_ARG((reinterpret_cast<uintptr_t>(Bt) & 15) == 0, "jd_gmm_prior_forward_tc: Bt must be 16-byte aligned");
  tc::Geom g{fH, fW, 0, 0, stride, nx, row_begin, (row_end - row_begin) * nx};
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(tc::gmm_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)tc::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("jd_gmm_prior_forward_tc: cannot reserve %zu B of shared memory: %s", tc::SMEM_BYTES,
                cudaGetErrorString(e));
      return JD_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = (g.P + tc::TM - 1) / tc::TM;
  tc::gmm_fwd_tc_kernel<<<grid, tc::NTHREADS, tc::SMEM_BYTES, to_stream(stream)>>>(
      flux, g, shift_yx, reinterpret_cast<const uint8_t*>(Bt), mw, ck, K, marginalize, value, argmax, logp, sum);
  JD_CHECK_LAUNCH("jd_gmm_prior_forward_tc");
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
