// Fused gradient all-reduce + Adam over NVLink peer memory (joint multi-GPU deconvolution).
//
// After the local gradient kernels every rank holds a partial flux gradient (its datasets' likelihood terms and
// its block of prior patch rows) in a symmetric-memory buffer that all peers can address.  Instead of
// ncclAllReduce followed by N identical Adam kernels, rank r owns the pixel slice [n r / G, n (r+1) / G):
// it loads the G partial gradients of that slice straight from the peers' buffers (P2P loads over
// NVLink / NVSwitch, fixed summation order), runs Adam on the slice (m, v only ever touched by the owner)
// and stores the updated theta slice into every replica (P2P stores).  Same bytes over the wire as a ring
// reduce-scatter + all-gather, one kernel, and the replicas are identical by construction.
// The caller brackets the kernel with cross-rank barriers (symmetric-memory signal pads).
#include "jd_common.cuh"

namespace jd {

__global__ void __launch_bounds__(256)
adam_allreduce_peer_kernel(const float* const* __restrict__ grad_ptrs, float* const* __restrict__ theta_ptrs, int rank,
                           int world, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ flux,
                           const uint8_t* __restrict__ mask, int use_log, int64_t lo4, int64_t hi4,
                           const float* __restrict__ scalars, float b1, float b2, float eps) {
  const float lr_over_bc1 = scalars[0], sqrt_bc2 = scalars[1];
  float* theta_own = theta_ptrs[rank];
  for (int64_t i4 = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < hi4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i4 * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {  // fixed order: bit-reproducible
      const float4 gq = *reinterpret_cast<const float4*>(grad_ptrs[q] + i);
      g.x += gq.x, g.y += gq.y, g.z += gq.z, g.w += gq.w;
    }
    float gg[4] = {g.x, g.y, g.z, g.w};
    const float4 t4 = *reinterpret_cast<const float4*>(theta_own + i);
    float th[4] = {t4.x, t4.y, t4.z, t4.w};
    const float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float4 f4 = use_log ? *reinterpret_cast<const float4*>(flux + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    float ff[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float gc = gg[c] * (use_log ? ff[c] : (mask ? (float)mask[i + c] : 1.f));
      mm[c] = mm[c] + (gc - mm[c]) * (1.f - b1);
      vv[c] = vv[c] * b2 + (1.f - b2) * gc * gc;
      const float denom = sqrtf(vv[c]) / sqrt_bc2 + eps;
      th[c] = th[c] - lr_over_bc1 * (mm[c] / denom);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const float4 tn = make_float4(th[0], th[1], th[2], th[3]);
    for (int q = 0; q < world; ++q) *reinterpret_cast<float4*>(theta_ptrs[q] + i) = tn;
  }
}

// ---- the same kernel with the two cross-rank barriers built in, so that a multi-rank joint step is ONE CUDA graph
// (no host-launched signal-pad barriers between graph replays).
//   sig_ptrs[q] -> rank q's flag block in symmetric memory: 2 x 32 uint32 (entry flags, exit flags), one slot per
//   peer rank.  `epoch` is a device counter that lives outside the restored optimiser state: the kernel bumps it at
//   its end, every rank launches the kernel the same number of times, flags only grow.
//   entry: thread 0 of CTA 0 publishes "my partial gradient is complete" (it is: the producing kernels precede this
//          one in stream order) to every peer; every CTA then waits until all peers have published this epoch.
//   exit : the last CTA of the grid to finish its theta stores publishes "my slice is stored in every replica" and
//          waits for the same from all peers; the kernel (and with it the stream) only completes after that, so
//          the next kernel of every rank sees the complete new theta.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// WORLD > 0: compile-time rank count - the WORLD peer loads of a pixel group are all in flight before the first add
// (a run-time loop issues them one NVLink round trip after the other); WORLD = 0: any rank count.
// `stamps` (optional, 4 x int64 of %globaltimer ns): kernel entry, entry barrier passed, last CTA's theta stores
// fenced, exit barrier passed - bench.py --breakdown turns them into wait / transfer shares.
template <int WORLD>
__global__ void __launch_bounds__(256)
adam_allreduce_peer_sync_kernel(const float* const* __restrict__ grad_ptrs, float* const* __restrict__ theta_ptrs,
                                uint32_t* const* __restrict__ sig_ptrs, uint32_t* __restrict__ epoch_ctr,
                                unsigned* __restrict__ done_ctr, int rank, int world, float* __restrict__ m,
                                float* __restrict__ v, const float* __restrict__ flux, const uint8_t* __restrict__ mask,
                                int use_log, int64_t lo4, int64_t hi4, const float* __restrict__ scalars, float b1,
                                float b2, float eps, long long* __restrict__ stamps) {
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(epoch_ctr) + 1u;
  uint32_t* my_sig = sig_ptrs[rank];
  if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[0] = global_timer_ns();
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(sig_ptrs[threadIdx.x] + rank, epoch);
  }
  if (threadIdx.x < world) {
    while ((int32_t)(ld_acquire_sys(my_sig + threadIdx.x) - epoch) < 0) __nanosleep(20);
  }
  __syncthreads();
  if (stamps && blockIdx.x == 0 && threadIdx.x == 0) stamps[1] = global_timer_ns();

  const float lr_over_bc1 = scalars[0], sqrt_bc2 = scalars[1];
  float* theta_own = theta_ptrs[rank];
  const float* gp[WORLD > 0 ? WORLD : 1];
  float* tp[WORLD > 0 ? WORLD : 1];
  if (WORLD > 0) {
#pragma unroll
    for (int q = 0; q < WORLD; ++q) gp[q] = grad_ptrs[q], tp[q] = theta_ptrs[q];
  }
  for (int64_t i4 = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i4 < hi4; i4 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i4 * 4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (WORLD > 0) {
      float4 gq[WORLD > 0 ? WORLD : 1];
#pragma unroll
      for (int q = 0; q < WORLD; ++q) gq[q] = __ldcg(reinterpret_cast<const float4*>(gp[q] + i));
#pragma unroll
      for (int q = 0; q < WORLD; ++q) g.x += gq[q].x, g.y += gq[q].y, g.z += gq[q].z, g.w += gq[q].w;  // fixed order
    } else {
      for (int q = 0; q < world; ++q) {  // fixed order: bit-reproducible
        const float4 gq = __ldcg(reinterpret_cast<const float4*>(grad_ptrs[q] + i));
        g.x += gq.x, g.y += gq.y, g.z += gq.z, g.w += gq.w;
      }
    }
    float gg[4] = {g.x, g.y, g.z, g.w};
    const float4 t4 = *reinterpret_cast<const float4*>(theta_own + i);
    float th[4] = {t4.x, t4.y, t4.z, t4.w};
    const float4 m4 = *reinterpret_cast<const float4*>(m + i), v4 = *reinterpret_cast<const float4*>(v + i);
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
    const float4 f4 = use_log ? *reinterpret_cast<const float4*>(flux + i) : make_float4(1.f, 1.f, 1.f, 1.f);
    float ff[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float gc = gg[c] * (use_log ? ff[c] : (mask ? (float)mask[i + c] : 1.f));
      mm[c] = mm[c] + (gc - mm[c]) * (1.f - b1);
      vv[c] = vv[c] * b2 + (1.f - b2) * gc * gc;
      const float denom = sqrtf(vv[c]) / sqrt_bc2 + eps;
      th[c] = th[c] - lr_over_bc1 * (mm[c] / denom);
    }
    *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    const float4 tn = make_float4(th[0], th[1], th[2], th[3]);
    if (WORLD > 0) {
#pragma unroll
      for (int q = 0; q < WORLD; ++q) *reinterpret_cast<float4*>(tp[q] + i) = tn;
    } else {
      for (int q = 0; q < world; ++q) *reinterpret_cast<float4*>(theta_ptrs[q] + i) = tn;
    }
  }

  // ---- exit barrier
  __shared__ int s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(done_ctr, 1u) + 1u == gridDim.x) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  if (stamps && threadIdx.x == 0) stamps[2] = global_timer_ns();
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(sig_ptrs[threadIdx.x] + 32 + rank, epoch);
    while ((int32_t)(ld_acquire_sys(my_sig + 32 + threadIdx.x) - epoch) < 0) __nanosleep(20);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (stamps) stamps[3] = global_timer_ns();
    *done_ctr = 0;
    *epoch_ctr = epoch;
    __threadfence();
  }
}

}  // namespace jd

using namespace jd;

extern "C" int jd_adam_allreduce_peer(const void* grad_ptrs_dev, const void* theta_ptrs_dev, int rank, int world,
                                      float* m, float* v, const float* flux, const uint8_t* mask, int use_log_flux,
                                      int64_t n, const float* adam_scalars, float beta1, float beta2, float eps,
                                      jd_stream_t stream) {
  JD_CHECK_ARG(grad_ptrs_dev && theta_ptrs_dev && m && v && adam_scalars, "jd_adam_allreduce_peer: null pointer");
  JD_CHECK_ARG(world >= 1 && rank >= 0 && rank < world && n > 0 && (n & 3) == 0,
               "jd_adam_allreduce_peer: bad rank/world or n not a multiple of 4 (n=%lld)", (long long)n);
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_allreduce_peer: flux required for the log parameterisation");
  const int64_t n4 = n / 4;
  const int64_t lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;
  if (hi4 <= lo4) return JD_OK;
  int64_t blocks = (hi4 - lo4 + 255) / 256;
  int64_t cap = (int64_t)num_sms() * 4;
  adam_allreduce_peer_kernel<<<(int)(blocks > cap ? cap : blocks), 256, 0, to_stream(stream)>>>(
      reinterpret_cast<const float* const*>(grad_ptrs_dev), reinterpret_cast<float* const*>(theta_ptrs_dev), rank, world,
      m, v, flux, mask, use_log_flux, lo4, hi4, adam_scalars, beta1, beta2, eps);
  JD_CHECK_LAUNCH("jd_adam_allreduce_peer");
  return JD_OK;
}

extern "C" int jd_adam_allreduce_peer_sync(const void* grad_ptrs_dev, const void* theta_ptrs_dev,
                                           const void* sig_ptrs_dev, uint32_t* sync_state, int rank, int world, float* m,
                                           float* v, const float* flux, const uint8_t* mask, int use_log_flux, int64_t n,
                                           const float* adam_scalars, float beta1, float beta2, float eps,
                                           jd_stream_t stream) {
  JD_CHECK_ARG(grad_ptrs_dev && theta_ptrs_dev && sig_ptrs_dev && sync_state && m && v && adam_scalars,
               "jd_adam_allreduce_peer_sync: null pointer");
  JD_CHECK_ARG(world >= 1 && world <= 32 && rank >= 0 && rank < world && n > 0 && (n & 3) == 0,
               "jd_adam_allreduce_peer_sync: bad rank/world (<= 32) or n not a multiple of 4 (n=%lld)", (long long)n);
  JD_CHECK_ARG(!use_log_flux || flux, "jd_adam_allreduce_peer_sync: flux required for the log parameterisation");
  const int64_t n4 = n / 4;
  const int64_t lo4 = n4 * rank / world, hi4 = n4 * (rank + 1) / world;
  int64_t blocks = (hi4 - lo4 + 255) / 256;
  // every CTA spins at the entry barrier: the grid must be co-resident (<= 2 CTAs of 256 threads per SM here)
  int64_t cap = (int64_t)num_sms() * 2;
  if (blocks < 1) blocks = 1;
  auto kern = world == 8   ? adam_allreduce_peer_sync_kernel<8>
              : world == 4 ? adam_allreduce_peer_sync_kernel<4>
              : world == 2 ? adam_allreduce_peer_sync_kernel<2>
                           : adam_allreduce_peer_sync_kernel<0>;
  // sync_state: [0] epoch, [1] finished CTAs, [2..3] unused, [4..11] four int64 time stamps of the last launch
  long long* stamps = reinterpret_cast<long long*>(sync_state + 4);
  kern<<<(int)(blocks > cap ? cap : blocks), 256, 0, to_stream(stream)>>>(
      reinterpret_cast<const float* const*>(grad_ptrs_dev), reinterpret_cast<float* const*>(theta_ptrs_dev),
      reinterpret_cast<uint32_t* const*>(sig_ptrs_dev), sync_state, sync_state + 1, rank, world, m, v, flux, mask,
      use_log_flux, lo4, hi4, adam_scalars, beta1, beta2, eps, stamps);
  JD_CHECK_LAUNCH("jd_adam_allreduce_peer_sync");
  return JD_OK;
}
