// Sub-pixel shift of the flux by an NPredCalibration (utils/torch.py:196-223, npred.py:226-230): per-pixel arithmetic
// shared by the CUDA kernels (jd_elementwise.cu) and the host check (tests/shift_host.cpp, plain g++).
//
// shift_image_torch = affine_grid + grid_sample (bilinear, zeros padding, align_corners=False) with a pure translation:
// output pixel (i, j) samples the input at (i + g_y shift_y, j + g_x shift_x), g = scale up to the float32 rounding of
// the reference's `2 * scale / torch.tensor([[W], [H]])` (scalar * reciprocal, in float32) - a 4-tap stencil with
// constant weights.  Derivatives w.r.t. the shift are those of the bilinear interpolant (grid_sample's backward).
#pragma once

#if defined(__CUDACC__)
#define JD_SHIFT_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define JD_SHIFT_HD inline
#endif

namespace jd {

struct ShiftTaps {
  int fy, fx;      // integer part of the sample offset
  float wy, wx;    // fractional part: weight of the +1 tap
  float gy, gx;    // d(sample position) / d(shift)
};

JD_SHIFT_HD ShiftTaps shift_taps(float shift_x, float shift_y, int scale, int H, int W) {
  ShiftTaps t;
  t.gy = ((2.f * (float)scale) * (1.f / (float)H)) * (0.5f * (float)H);
  t.gx = ((2.f * (float)scale) * (1.f / (float)W)) * (0.5f * (float)W);
  const float dy = t.gy * shift_y, dx = t.gx * shift_x;
  const float fy = floorf(dy), fx = floorf(dx);
  t.fy = (int)fy, t.fx = (int)fx;
  t.wy = dy - fy, t.wx = dx - fx;
  return t;
}

JD_SHIFT_HD float shift_at(const float* img, int H, int W, int y, int x) {
  return (y >= 0 && y < H && x >= 0 && x < W) ? img[(long long)y * W + x] : 0.f;
}

// forward: shifted[i, j]
JD_SHIFT_HD float shift_sample(const float* img, int H, int W, int i, int j, const ShiftTaps& t) {
  const float a00 = shift_at(img, H, W, i + t.fy, j + t.fx), a01 = shift_at(img, H, W, i + t.fy, j + t.fx + 1);
  const float a10 = shift_at(img, H, W, i + t.fy + 1, j + t.fx), a11 = shift_at(img, H, W, i + t.fy + 1, j + t.fx + 1);
  return (1.f - t.wy) * ((1.f - t.wx) * a00 + t.wx * a01) + t.wy * ((1.f - t.wx) * a10 + t.wx * a11);
}

// adjoint w.r.t. the image: dimage[m, n] = sum_ab w_ab d[m - fy - a, n - fx - b]
JD_SHIFT_HD float shift_adjoint(const float* d, int H, int W, int m, int n, const ShiftTaps& t) {
  const float d00 = shift_at(d, H, W, m - t.fy, n - t.fx), d01 = shift_at(d, H, W, m - t.fy, n - t.fx - 1);
  const float d10 = shift_at(d, H, W, m - t.fy - 1, n - t.fx), d11 = shift_at(d, H, W, m - t.fy - 1, n - t.fx - 1);
  return (1.f - t.wy) * ((1.f - t.wx) * d00 + t.wx * d01) + t.wy * ((1.f - t.wx) * d10 + t.wx * d11);
}

// d shifted[i, j] / d shift_y and / d shift_x
JD_SHIFT_HD void shift_dshift(const float* img, int H, int W, int i, int j, const ShiftTaps& t, float* d_dy,
                              float* d_dx) {
  const float a00 = shift_at(img, H, W, i + t.fy, j + t.fx), a01 = shift_at(img, H, W, i + t.fy, j + t.fx + 1);
  const float a10 = shift_at(img, H, W, i + t.fy + 1, j + t.fx), a11 = shift_at(img, H, W, i + t.fy + 1, j + t.fx + 1);
  *d_dy = t.gy * ((1.f - t.wx) * (a10 - a00) + t.wx * (a11 - a01));
  *d_dx = t.gx * ((1.f - t.wy) * (a01 - a00) + t.wy * (a11 - a10));
}

}  // namespace jd
