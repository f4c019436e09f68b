// Shared helpers for the jolideco_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "jolideco_b200.h"

namespace jd {

void set_error(const char* fmt, ...);

#define JD_CHECK_ARG(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      jd::set_error(__VA_ARGS__);    \
      return JD_ERR_INVALID;         \
    }                                \
  } while (0)

#define JD_CHECK_LAUNCH(name)                                                        \
  do {                                                                               \
    cudaError_t e_ = cudaGetLastError();                                             \
    if (e_ != cudaSuccess) {                                                         \
      jd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_));          \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? JD_ERR_NO_DEVICE \
                                                                             : JD_ERR_CUDA;    \
    }                                                                                \
  } while (0)

constexpr int PATCH = 8;
constexpr int PD = 64;  // patch elements

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of doubles; result valid in thread 0. `red` needs >= 32 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  double r = 0.0;
  if (w == 0) {
    r = lane < nw ? red[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__device__ __forceinline__ int wrap(int a, int n) {  // a mod n for a in [-n, 2n)
  a = a < 0 ? a + n : a;
  return a >= n ? a - n : a;
}

// true the first time it is called for (this call site's flag array, current device): function attributes
// (cudaFuncSetAttribute) are per device / context, so the opt-in is repeated on every device that is used
inline bool first_use_on_device(bool (&seen)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (seen[dev & 63]) return false;
  seen[dev & 63] = true;
  return true;
}

inline cudaStream_t to_stream(jd_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int num_sms();

}  // namespace jd
