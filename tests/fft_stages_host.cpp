// Host check of jolideco_b200/csrc/jd_fft_stages.cuh (the exact stage code the CUDA FFT kernels run): mixed-radix
// Stockham FFT of length r0 * 2^L against a naive O(N^2) DFT in double precision, forward and inverse.
// Built and run by tests/test_fft_host.py with g++ (no CUDA needed).  Prints "OK" or the first failure.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "jd_fft_stages.cuh"

using namespace jd::fft;

template <bool INV>
static double check(int N, int L, int r0) {
  std::vector<float2> x(N), y(N), tw(N / 2), in(N);
  for (int j = 0; j < N / 2; ++j) {
    const double a = -2.0 * M_PI * j / N;
    tw[j] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  srand(N * 7 + (INV ? 1 : 0));
  for (int j = 0; j < N; ++j) in[j] = x[j] = make_float2(rand() / (float)RAND_MAX - 0.5f, rand() / (float)RAND_MAX - 0.5f);
  auto for_each = [](int items, auto f) {
    for (int t = 0; t < items; ++t) f(t);
  };
  float2* z = fft_mixed<INV>(x.data(), y.data(), tw.data(), N, L, r0, for_each);
  double err = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    double re = 0, im = 0;
    for (int j = 0; j < N; ++j) {
      const double a = (INV ? 2.0 : -2.0) * M_PI * (double)((long long)j * k % N) / N;
      re += in[j].x * std::cos(a) - in[j].y * std::sin(a);
      im += in[j].x * std::sin(a) + in[j].y * std::cos(a);
    }
    err = std::fmax(err, std::fmax(std::fabs(re - z[k].x), std::fabs(im - z[k].y)));
    scale = std::fmax(scale, std::fmax(std::fabs(re), std::fabs(im)));
  }
  return err / scale;
}

int main() {
  const int sizes[] = {2, 4, 6, 8, 10, 12, 16, 20, 24, 32, 40, 48, 64, 80, 96, 160, 192, 320, 384, 512, 640, 768, 1280};
  for (int N : sizes) {
    int r0 = (N % 3 == 0) ? 3 : (N % 5 == 0) ? 5 : 1, L = 0;
    for (int n = N / r0; n > 1; n >>= 1) ++L;
    if ((r0 << L) != N) {
      printf("bad size %d\n", N);
      return 1;
    }
    const double ef = check<false>(N, L, r0), ei = check<true>(N, L, r0);
    if (!(ef < 2e-6) || !(ei < 2e-6)) {
      printf("FAIL N=%d (r0=%d, L=%d): forward %.3g inverse %.3g\n", N, r0, L, ef, ei);
      return 1;
    }
  }
  // size selection
  int L, r0;
  if (fft_size(545, true, &L, &r0) != 640 || r0 != 5 || L != 7) return printf("fft_size(545) wrong\n"), 1;
  if (fft_size(1224, true, &L, &r0) != 1280 || r0 != 5) return printf("fft_size(1224) wrong\n"), 1;
  if (fft_size(700, true, &L, &r0) != 768 || r0 != 3) return printf("fft_size(700) wrong\n"), 1;
  if (fft_size(512, true, &L, &r0) != 512 || r0 != 1 || L != 9) return printf("fft_size(512) wrong\n"), 1;
  if (fft_size(545, false, &L, &r0) != 1024 || r0 != 1 || L != 10) return printf("fft_size(545, pow2) wrong\n"), 1;
  if (fft_size(1, true, &L, &r0) != 2) return printf("fft_size(1) wrong\n"), 1;
  printf("OK\n");
  return 0;
}
