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

// ------------------------------------------------------------------ bulk-TMA ingest per SM
// Every CTA walks the `n_img` operand images of `img_bytes` (L2 resident) `rounds` times through an NST-stage ring.
// mode 0: unicast, the CTA copies `bytes` of every image itself;  mode 1: 2-CTA cluster, each CTA copies one half of
// `bytes` and multicasts it to both (the prior kernels' scheme).  One producer thread, one consumer thread that
// frees the stage as soon as it is full.
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
      mbar_wait(bar0 + 8 * (nst + s), ((pos / nst) & 1) ^ 1);
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
      mbar_wait(bar0 + 8 * s, (pos / nst) & 1);
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

  // ---- TMA ingest
  const int n_img = 256, img_bytes = 32768;
  uint8_t* d_img;
  CK(cudaMalloc(&d_img, (size_t)n_img * img_bytes));
  CK(cudaMemset(d_img, 1, (size_t)n_img * img_bytes));
  CK(cudaFuncSetAttribute(tma_ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct Case {
    int mode, bytes, nst, grid;
  };
  const Case cases[] = {
      {0, 32768, 6, sms},  {1, 32768, 6, sms},  {0, 16384, 6, sms},     {0, 16384, 12, sms}, {1, 16384, 12, sms},
      {0, 8192, 12, sms},  {0, 8192, 24, sms},  {1, 32768, 6, sms / 2}, {0, 32768, 6, sms / 2},
      {0, 28672, 6, sms},  {0, 10240, 16, sms}, {0, 4096, 24, sms},
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
      CK(cudaLaunchKernelEx(&cfg, tma_ingest_kernel, (const uint8_t*)d_img, n_img, img_bytes, c.bytes, rounds, c.nst,
                            c.mode, d_cyc));
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
    }
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double cyc = mean_cycles(d_cyc, cfg.gridDim.x);
    const double per_pos = cyc / (n_img * rounds);
    printf("tma ingest %s, %5d B per position and SM, %2d stages, %3d CTAs: %.0f clk per position, %.1f B/clk/SM "
           "landed, L2 reads %.2f TB/s, %.1f us\n",
           c.mode ? "pair-multicast" : "unicast       ", c.bytes, c.nst, cfg.gridDim.x, per_pos, c.bytes / per_pos,
           (double)cfg.gridDim.x * (c.mode ? c.bytes / 2 : c.bytes) * n_img * rounds / (ms * 1e-3) / 1e12, ms * 1e3);
  }
  return 0;
}
