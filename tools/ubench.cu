// Micro-benchmarks behind the design of the tcgen05 prior kernels: what one SM can ingest from L2
// through bulk-TMA while every SM streams the same operand image, and how fast the epilogue warps can read FP32
// accumulators out of TMEM.  Build (in-tree, travels to the GPU box):
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jolideco_b200/csrc -I include \
//          tools/ubench.cu -o tools/ubench
//     tools/ubench
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "jd_tc_ptx.cuh"

using namespace jd::tcx;

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e_ = (x);                                                            \
    if (e_ != cudaSuccess) {                                                         \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                       \
    }                                                                                \
  } while (0)

// ------------------------------------------------------------------ TMEM read throughput
// nw warps (multiple of 4) of one CTA per SM; warp w reads lanes [32 (w & 3), +32), `cols` columns per iteration
template <int COLS>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * COLS % 512;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float y[32];
#pragma unroll
    for (int c = 0; c < COLS; c += 32) {
      tmem_ld32(base + ((it * COLS + c) & 255), y);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc = fmaf(y[i], y[i], acc);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(s_tmem, 512);
  }
}

// same, but both loads of a 64-column accumulator in flight before the wait (the shape of the epilogue)
__global__ void __launch_bounds__(512, 1) tmem_ld64_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float y0[32], y1[32];
    const uint32_t col = ((it + (warp >> 2)) * 64) & 511;
    tmem_ld32(base + col, y0);
    tmem_ld32(base + col + 32, y1);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) acc = fmaf(y0[i], y0[i], fmaf(y1[i], y1[i], acc));
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(s_tmem, 512);
  }
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, float (&v)[64]) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,"
      "%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// MODE 0: one 32x32b.x64 load + wait per accumulator;  MODE 1: software pipeline - the load of accumulator i+1 is in
// flight while the squares of accumulator i are summed (two register buffers of 32 columns, x32 loads);
// MODE 2: 16 columns per load (x16), four in flight before the wait
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_ld_var_kernel(int iters, long long* cycles, float* sink) {
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    for (int it = 0; it < iters; ++it) {
      float y[64];
      tmem_ld64(base + (((it + (warp >> 2)) * 64) & 511), y);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; ++i) q[i & 7] = fmaf(y[i], y[i], q[i & 7]);
    }
  } else if (MODE == 1) {
    float ya[32], yb[32];
    tmem_ld32(base, ya);
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = ((it + (warp >> 2)) * 64) & 511;
      tmem_ld_wait();
      tmem_ld32(base + col + 32, yb);
#pragma unroll
      for (int i = 0; i < 32; ++i) q[i & 7] = fmaf(ya[i], ya[i], q[i & 7]);
      tmem_ld_wait();
      tmem_ld32(base + ((col + 64) & 511), ya);
#pragma unroll
      for (int i = 0; i < 32; ++i) q[i & 7] = fmaf(yb[i], yb[i], q[i & 7]);
    }
    tmem_ld_wait();
  } else {
    for (int it = 0; it < iters; ++it) {
      const uint32_t col = ((it + (warp >> 2)) * 64) & 511;
      uint32_t r[64];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[16 * c]), "=r"(r[16 * c + 1]), "=r"(r[16 * c + 2]), "=r"(r[16 * c + 3]), "=r"(r[16 * c + 4]),
              "=r"(r[16 * c + 5]), "=r"(r[16 * c + 6]), "=r"(r[16 * c + 7]), "=r"(r[16 * c + 8]), "=r"(r[16 * c + 9]),
              "=r"(r[16 * c + 10]), "=r"(r[16 * c + 11]), "=r"(r[16 * c + 12]), "=r"(r[16 * c + 13]),
              "=r"(r[16 * c + 14]), "=r"(r[16 * c + 15])
            : "r"(base + col + 16 * c));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        const float a = __uint_as_float(r[i]);
        q[i & 7] = fmaf(a, a, q[i & 7]);
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (q[0] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + q[7] == 123.456f) sink[0] = q[0];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(s_tmem, 512);
  }
}


// ------------------------------------------------------------------ tcgen05.mma rate vs N (A operand in TMEM)
// one elected lane issues `iters` MMAs  D[128 x N] += A[128 x 32 B] . B[N x 32 B]^T  back to back into the same
// accumulator, then commits and waits: clk per MMA as a function of N, for kind::f16 (K = 16) and kind::tf32 (K = 8)
template <bool F16>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, int n, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | (((uint32_t)n >> 3) << 17) |
                         ((128u >> 4) << 24);
  if (warp == 1) {
    const uint64_t desc = make_desc(smem_u32(smem));
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        if (F16)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + 256),
                       "r"(tmem + (uint32_t)((i & 7) * 8)), "l"(desc), "r"(idesc)
                       : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem + 256),
                       "r"(tmem + (uint32_t)((i & 7) * 8)), "l"(desc), "r"(idesc)
                       : "memory");
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    while (!mbar_try(smem_u32(&bar), 0)) {
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// the prior kernels' real MMA sequence (recipe 1: three passes of N = 64, 48, 32, 16 at row offsets 0, 16, 32, 48 into
// one 64-column accumulator, first MMA overwrites) issued by `nw` warps concurrently, each warp into its own accumulator
// slot(s), optionally with `nld` more warps streaming accumulators out of TMEM (tcgen05.ld) at the same time
__global__ void __launch_bounds__(512, 1) mma_pattern_kernel(int iters, int nw, int nld, int dense, int ncommit, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t s_tmem;
  __shared__ uint64_t bar[4];
  __shared__ uint64_t sink_bar[4];  // arrival count never reached: per-position commits land here
  __shared__ volatile int s_stop;
  const int warp = threadIdx.x >> 5;
  const bool rnd = (dense & 2) != 0;  // random FP16 operands (|x| < 2) instead of zeros: data-dependent power
  dense &= 1;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) {
    uint32_t r = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    r ^= r >> 13, r *= 2246822519u, r ^= r >> 16;
    reinterpret_cast<uint32_t*>(smem)[i] = rnd ? (r & 0xBFFFBFFFu) : 0u;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&sink_bar[i]), (1u << 20) - 1);
    s_stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t sb = desc_lo(smem_u32(smem));
  if (rnd && warp < 4) {  // random A operand in TMEM columns [0, 64)
    float v[32];
    for (int c = 0; c < 64; c += 32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        uint32_t r = (uint32_t)(threadIdx.x * 64 + c + i) * 2654435761u;
        r ^= r >> 13, r *= 2246822519u, r ^= r >> 16;
        v[i] = __uint_as_float(r & 0xBFFFBFFFu);
      }
      tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < nw) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const uint32_t d = tmem + 128 + ((uint32_t)(warp * 2 + (it & 1)) * 64) % 384;
        const uint32_t b = sb + (uint32_t)((it % 4) * (16384 >> 4));
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t n0 = dense ? 0u : 16u * kk;
            const uint32_t idesc = (1u << 4) | (((64u - n0) >> 3) << 17) | ((128u >> 4) << 24);
            const uint64_t desc = desc_from_lo(b + ((pass == 1 ? 8192u : 0u) >> 4) + ((kk * 32 + n0 * 128) >> 4));
            const uint32_t a = tmem + (pass == 0 ? 32u : 0u) + kk * 8;
            if (pass == 0 && kk == 0)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d + n0),
                           "r"(a), "l"(desc), "r"(idesc)
                           : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d + n0),
                           "r"(a), "l"(desc), "r"(idesc)
                           : "memory");
          }
        for (int c = 0; c < ncommit; ++c) umma_commit(smem_u32(&sink_bar[(warp + c) & 3]));
      }
      umma_commit(smem_u32(&bar[warp]));
    }
    __syncwarp();
    while (!mbar_try(smem_u32(&bar[warp]), 0)) {
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 4 + warp] = t1 - t0;
    __syncwarp();
    if ((threadIdx.x & 31) == 0) atomicAdd((int*)&s_stop, 1);
  } else if (warp >= 4 && warp < 4 + nld) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 128;
    float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int it = 0;
    while (s_stop < nw) {
      float y0[32], y1[32];
      tmem_ld32(base + ((it * 64) % 384), y0);
      tmem_ld32(base + ((it * 64) % 384) + 32, y1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) q[i & 7] = fmaf(y0[i], y0[i], fmaf(y1[i], y1[i], q[i & 7]));
      ++it;
    }
    if (q[0] + q[1] + q[2] + q[3] + q[4] + q[5] + q[6] + q[7] == 123.456f) cycles[1022] = it;
    if (blockIdx.x == 0 && threadIdx.x == 128) cycles[1023] = it;  // accumulators streamed by one reader warp
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------ mbarrier wait flavours
// WAIT 0: mbarrier.try_wait (may suspend the thread), 1: mbarrier.test_wait spin, 2: try_wait with a 32 ns suspend hint
template <int WAIT>
__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  if (WAIT == 0) {
    while (!mbar_try(bar, parity)) {
    }
  } else if (WAIT == 1) {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity)
          : "memory");
    }
  } else {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar), "r"(parity), "r"(32)
          : "memory");
    }
  }
}

// two warps hand a token back and forth through two mbarriers: cycles per one-way handoff
template <int WAIT>
__global__ void pingpong_kernel(int iters, long long* cycles) {
  __shared__ uint64_t bars[2];
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  if (threadIdx.x == 0) {
    mbar_init(b0, 1);
    mbar_init(b1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
      mbar_arrive(b0);
      wait_bar<WAIT>(b1, it & 1);
    }
  } else if (threadIdx.x == 32) {
    for (int it = 0; it < iters; ++it) {
      wait_bar<WAIT>(b0, it & 1);
      mbar_arrive(b1);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

// ------------------------------------------------------------------ bulk-TMA ingest per SM
// Every CTA walks the `n_img` operand images of `img_bytes` (L2 resident) `rounds` times through an NST-stage ring.
// mode 0: unicast, the CTA copies `bytes` of every image itself;  mode 1: 2-CTA cluster, each CTA copies one half of
// `bytes` and multicasts it to both (the prior kernels' scheme).  One producer thread, one consumer thread that
// frees the stage as soon as it is full.
template <int WAIT>
__global__ void __launch_bounds__(128, 1)
tma_ingest_kernel(const uint8_t* __restrict__ img, int n_img, int img_bytes, int bytes, int rounds, int nst, int mode,
                  long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t crank = mode ? cluster_ctarank() : 0;
  const int ncta = mode ? 2 : 1;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nst; ++s) {
      mbar_init(bar0 + 8 * s, 1);
      mbar_init(bar0 + 8 * (nst + s), ncta);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (mode) cluster_sync_all();
  const int total = n_img * rounds;
  const int rot = (blockIdx.x / ncta) * 37 % n_img;
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int pos = 0; pos < total; ++pos) {
      const int s = pos % nst;
      wait_bar<WAIT>(bar0 + 8 * (nst + s), ((pos / nst) & 1) ^ 1);
      const int k = (pos + rot) % n_img;
      mbar_arrive_expect_tx(bar0 + 8 * s, bytes);
      const uint32_t dst = smem_u32(ring + (size_t)s * bytes);
      if (mode) {
        const int half = bytes / 2;
        bulk_g2s_mc(dst + crank * half, img + (size_t)k * img_bytes + crank * half, half, bar0 + 8 * s, 3);
      } else {
        bulk_g2s(dst, img + (size_t)k * img_bytes, bytes, bar0 + 8 * s);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int pos = 0; pos < total; ++pos) {
      const int s = pos % nst;
      wait_bar<WAIT>(bar0 + 8 * s, (pos / nst) & 1);
      if (mode) {  // free the stage in both CTAs of the pair
        for (int c = 0; c < 2; ++c) {
          uint32_t remote;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar0 + 8 * (nst + s)), "r"(c));
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
        }
      } else {
        mbar_arrive(bar0 + 8 * (nst + s));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (mode) cluster_sync_all();
}

static double mean_cycles(long long* d_cyc, int n) {
  long long* h = (long long*)malloc(n * sizeof(long long));
  CK(cudaMemcpy(h, d_cyc, n * sizeof(long long), cudaMemcpyDeviceToHost));
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)h[i];
  free(h);
  return s / n;
}

int main() {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  printf("SMs %d\n", sms);
  long long* d_cyc;
  float* d_sink;
  CK(cudaMalloc(&d_cyc, 1024 * sizeof(long long)));
  CK(cudaMalloc(&d_sink, 16));

  // ---- TMEM read
  const int iters = 4096;
  for (int nw = 4; nw <= 16; nw *= 2) {
    for (int rep = 0; rep < 2; ++rep) tmem_ld_kernel<32><<<sms, nw * 32>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    double c = mean_cycles(d_cyc, sms);
    printf("tmem_ld 32x32b.x32 + wait, %2d warps: %.1f clk per 32-col load per warp, %.1f B/clk/SM\n", nw, c / iters,
           (double)nw * 32 * 32 * 4 * iters / c);
    for (int rep = 0; rep < 2; ++rep) tmem_ld64_kernel<<<sms, nw * 32>>>(iters, d_cyc, d_sink);
    CK(cudaDeviceSynchronize());
    c = mean_cycles(d_cyc, sms);
    printf("tmem_ld 2 x (32x32b.x32) + wait, %2d warps: %.1f clk per 64-col accumulator per warp, %.1f B/clk/SM\n", nw,
           c / iters, (double)nw * 32 * 64 * 4 * iters / c);
  }

  for (int nw = 4; nw <= 16; nw *= 2) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_ld_var_kernel<0><<<sms, nw * 32>>>(iters, d_cyc, d_sink);
        if (mode == 1) tmem_ld_var_kernel<1><<<sms, nw * 32>>>(iters, d_cyc, d_sink);
        if (mode == 2) tmem_ld_var_kernel<2><<<sms, nw * 32>>>(iters, d_cyc, d_sink);
      }
      CK(cudaDeviceSynchronize());
      const double c = mean_cycles(d_cyc, sms);
      const char* names[3] = {"one 32x32b.x64 + wait + 64 FFMA", "pipelined x32 loads + 64 FFMA", "4 x (32x32b.x16) + wait + 64 FFMA"};
      printf("tmem_ld %s, %2d warps: %.1f clk per 64-col accumulator per warp, %.1f B/clk/SM\n", names[mode], nw,
             c / iters, (double)nw * 32 * 64 * 4 * iters / c);
    }
  }


  // ---- MMA rate vs N
  {
    CK(cudaFuncSetAttribute(mma_rate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    CK(cudaFuncSetAttribute(mma_rate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    const int ns[] = {16, 32, 48, 64, 96, 128, 192, 256};
    const int it3 = 2048;
    for (int f16 = 1; f16 >= 0; --f16)
      for (int n : ns) {
        for (int rep = 0; rep < 2; ++rep) {
          if (f16) mma_rate_kernel<true><<<sms, 128, 34 * 1024>>>(it3, n, d_cyc);
          else mma_rate_kernel<false><<<sms, 128, 34 * 1024>>>(it3, n, d_cyc);
        }
        CK(cudaDeviceSynchronize());
        const double c = mean_cycles(d_cyc, sms);
        printf("tcgen05.mma kind::%s M=128 N=%3d K=%2d, A in TMEM: %.1f clk per MMA (N/2 = %d)\n", f16 ? "f16 " : "tf32", n,
               f16 ? 16 : 8, c / it3, n / 2);
      }
  }


  // ---- the kernels' MMA pattern: issuing warps x concurrent TMEM readers
  {
    CK(cudaFuncSetAttribute(mma_pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));
    const int it4 = 512;
    const int cfgs[][4] = {{1, 0, 0, 0}, {1, 0, 1, 0}, {3, 0, 0, 0}, {3, 8, 0, 0}, {1, 0, 0, 1}, {1, 0, 0, 2}, {1, 0, 0, 3},
                            {3, 0, 0, 1}, {3, 0, 0, 3}, {3, 8, 0, 3}, {3, 8, 2, 3}, {1, 0, 2, 0}, {3, 8, 3, 3}};
    for (auto& c : cfgs) {
      for (int rep = 0; rep < 2; ++rep) mma_pattern_kernel<<<sms, 512, 66 * 1024>>>(it4, c[0], c[1], c[2], c[3], d_cyc);
      CK(cudaDeviceSynchronize());
      long long h[1024];
      CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost));
      double mx = 0;
      for (int b = 0; b < 8; ++b)
        for (int w = 0; w < c[0]; ++w) mx = h[b * 4 + w] > mx ? (double)h[b * 4 + w] : mx;
      printf("mma pattern (12 MMAs per tile-position, %s), %d issuing warps, %d tcgen05.ld warps, %d commits per "
             "tile-position: %.0f clk per tile-position per SM (%.1f clk per MMA)\n",
             (c[2] & 1) ? ((c[2] & 2) ? "dense N=64, random operands" : "dense N=64") : ((c[2] & 2) ? "trimmed, random operands" : "trimmed N=64,48,32,16"), c[0], c[1], c[3], mx / (it4 * c[0]),
             mx / (it4 * c[0] * 12.0));
      if (c[1]) printf("    a reader warp took %.0f clk per 64-column accumulator (2 x tcgen05.ld.x32 + wait + 64 FFMA)\n", mx / (double)h[1023]);
    }
  }

  // ---- mbarrier handoff latency
  {
    const int it2 = 20000;
    pingpong_kernel<0><<<1, 64>>>(it2, d_cyc);
    CK(cudaDeviceSynchronize());
    printf("mbarrier ping-pong, try_wait          : %.1f clk per one-way handoff\n", mean_cycles(d_cyc, 1) / it2 / 2);
    pingpong_kernel<1><<<1, 64>>>(it2, d_cyc);
    CK(cudaDeviceSynchronize());
    printf("mbarrier ping-pong, test_wait spin    : %.1f clk per one-way handoff\n", mean_cycles(d_cyc, 1) / it2 / 2);
    pingpong_kernel<2><<<1, 64>>>(it2, d_cyc);
    CK(cudaDeviceSynchronize());
    printf("mbarrier ping-pong, try_wait hint 32ns: %.1f clk per one-way handoff\n", mean_cycles(d_cyc, 1) / it2 / 2);
  }

  // ---- TMA ingest
  const int n_img = 256, img_bytes = 32768;
  uint8_t* d_img;
  CK(cudaMalloc(&d_img, (size_t)n_img * img_bytes));
  CK(cudaMemset(d_img, 1, (size_t)n_img * img_bytes));
  CK(cudaFuncSetAttribute(tma_ingest_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(tma_ingest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(tma_ingest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Case {
    int mode, bytes, nst, grid, wait;
  };
  const Case cases[] = {
      {0, 32768, 6, sms, 0},  {0, 32768, 6, sms, 1},  {0, 32768, 6, sms, 2},  {1, 32768, 6, sms, 0},
      {1, 32768, 6, sms, 1},  {0, 16384, 6, sms, 1},  {0, 16384, 12, sms, 1}, {1, 16384, 12, sms, 1},
      {0, 8192, 12, sms, 1},  {0, 8192, 24, sms, 1},  {0, 28672, 6, sms, 1},  {0, 10240, 16, sms, 1},
      {0, 4096, 24, sms, 1},  {0, 4096, 24, sms, 0},  {0, 32768, 6, sms / 2, 1}, {0, 32768, 3, sms, 1},
      {0, 32768, 2, sms, 1},
  };
  for (const Case& c : cases) {
    const int rounds = 4;
    const size_t smem = 2048 + (size_t)c.nst * c.bytes;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c.grid / (c.mode ? 2 : 1) * (c.mode ? 2 : 1));
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.mode ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      auto kern = c.wait == 0 ? tma_ingest_kernel<0> : (c.wait == 1 ? tma_ingest_kernel<1> : tma_ingest_kernel<2>);
      CK(cudaLaunchKernelEx(&cfg, kern, (const uint8_t*)d_img, n_img, img_bytes, c.bytes, rounds, c.nst, c.mode, d_cyc));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = mean_cycles(d_cyc, cfg.gridDim.x);
    const double per_pos = cyc / (n_img * rounds);
    printf("tma ingest %s, wait %d, %5d B per position and SM, %2d stages, %3d CTAs: %.0f clk per position, %.1f B/clk/SM "
           "landed, L2 reads %.2f TB/s, %.1f us\n",
           c.mode ? "pair-multicast" : "unicast       ", c.wait, c.bytes, c.nst, cfg.gridDim.x, per_pos, c.bytes / per_pos,
           (double)cfg.gridDim.x * (c.mode ? c.bytes / 2 : c.bytes) * n_img * rounds / (ms * 1e-3) / 1e12, ms * 1e3);
  }
  return 0;
}
