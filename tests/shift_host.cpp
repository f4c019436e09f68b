// Host build of the per-pixel shift arithmetic the CUDA kernels use (jolideco_b200/csrc/jd_shift.cuh), exposed with a C
// ABI so that tests/test_shift_host.py can compare it with the oracle (which is pinned to the imported reference).
#include "jd_shift.cuh"

extern "C" {

void shift_forward_host(const float* img, float sx, float sy, int scale, int H, int W, float* out) {
  const jd::ShiftTaps t = jd::shift_taps(sx, sy, scale, H, W);
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j) out[i * W + j] = jd::shift_sample(img, H, W, i, j, t);
}

// dimage = shift^T d;  dshift_xy = (sum d * d_dx, sum d * d_dy)
void shift_backward_host(const float* d, const float* img, float sx, float sy, int scale, int H, int W, float* dimage,
                         double* dshift_xy) {
  const jd::ShiftTaps t = jd::shift_taps(sx, sy, scale, H, W);
  double ax = 0, ay = 0;
  for (int i = 0; i < H; ++i)
    for (int j = 0; j < W; ++j) {
      dimage[i * W + j] = jd::shift_adjoint(d, H, W, i, j, t);
      float dy, dx;
      jd::shift_dshift(img, H, W, i, j, t, &dy, &dx);
      ax += (double)d[i * W + j] * dx;
      ay += (double)d[i * W + j] * dy;
    }
  dshift_xy[0] = ax;
  dshift_xy[1] = ay;
}
}
