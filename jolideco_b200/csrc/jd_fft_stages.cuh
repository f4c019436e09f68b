// Stockham autosort FFT stages shared by the CUDA kernels (jd_fft.cu) and the host check
// (tests/fft_stages_host.cpp, plain g++): one work item of one stage per call, so the kernels loop over
// `t = threadIdx.x; t < items; t += blockDim.x` and the host check over `t = 0 .. items`.
//
// Length N = r0 * 2^L with r0 in {1, 3, 5}.  Stage order: radix-4 stages while the remaining power of two allows,
// one radix-2 stage if L is odd, then ONE final radix-r0 stage.  Putting the odd radix last keeps every stride s a
// power of two (q = t & (s-1), p = t >> log2 s) and makes the odd stage twiddle-free (p = 0, like the last radix-2
// stage of a pure power of two).  General stage (radix r, current length n, stride s, m = n / r):
//     y[q + s (r p + k)] = w_n^{p k} * sum_j x[q + s (p + m j)] w_r^{j k},      p < m, q < s, then n <- m, s <- s r.
// tw[j] = exp(-2 pi i j / N) for j < N / 2; w_n^p = tw[p * (N / n)] = tw[p << done] with done = log2(N / n).
// INV conjugates every twiddle / butterfly constant (no scaling).
#pragma once

#if defined(__CUDACC__)
#define JD_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define JD_HD inline
struct float2 {
  float x, y;
};
static inline float2 make_float2(float x, float y) {
  float2 r;
  r.x = x;
  r.y = y;
  return r;
}
#endif

namespace jd {
namespace fft {

JD_HD float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
JD_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
JD_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// -i v (forward) or +i v (inverse)
template <bool INV>
JD_HD float2 rot(float2 v) {
  return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
}

// work item t < N / 4
template <bool INV>
JD_HD void stage_radix4(const float2* x, float2* y, const float2* tw, int n, int ls, int done, int t) {
  const int m = n >> 2, s = 1 << ls;
  const int q = t & (s - 1), p = t >> ls;
  float2 w1 = tw[p << done];        // exp(-2 pi i p / n)
  float2 w2 = tw[(2 * p) << done];  // exp(-2 pi i 2p / n), 2p < n/2
  if (INV) {
    w1.y = -w1.y;
    w2.y = -w2.y;
  }
  const float2 w3 = cmul(w1, w2);
  const float2 a = x[q + s * p], b = x[q + s * (p + m)], c = x[q + s * (p + 2 * m)], d = x[q + s * (p + 3 * m)];
  const float2 apc = cadd(a, c), amc = csub(a, c), bpd = cadd(b, d);
  const float2 jb = rot<INV>(csub(b, d));
  float2* o = y + q + s * (4 * p);
  o[0] = cadd(apc, bpd);
  o[s] = cmul(cadd(amc, jb), w1);
  o[2 * s] = cmul(csub(apc, bpd), w2);
  o[3 * s] = cmul(csub(amc, jb), w3);
}

// work item t < N / 2
template <bool INV>
JD_HD void stage_radix2(const float2* x, float2* y, const float2* tw, int n, int ls, int done, int t) {
  const int m = n >> 1, s = 1 << ls;
  const int q = t & (s - 1), p = t >> ls;
  float2 w = tw[p << done];
  if (INV) w.y = -w.y;
  const float2 a = x[q + s * p], b = x[q + s * (p + m)];
  float2* o = y + q + s * (2 * p);
  o[0] = cadd(a, b);
  o[s] = cmul(csub(a, b), w);
}

// final radix-3 stage (n == 3, p == 0, no twiddles): work item t < s = N / 3
template <bool INV>
JD_HD void stage_radix3_last(const float2* x, float2* y, int s, int t) {
  const float2 a = x[t], b = x[t + s], c = x[t + 2 * s];
  const float2 t1 = cadd(b, c);
  const float2 t2 = make_float2(a.x - 0.5f * t1.x, a.y - 0.5f * t1.y);
  const float2 d = csub(b, c);
  const float2 t3 = rot<INV>(make_float2(0.86602540378443865f * d.x, 0.86602540378443865f * d.y));  // -+ i sqrt(3)/2 (b - c)
  y[t] = cadd(a, t1);
  y[t + s] = cadd(t2, t3);
  y[t + 2 * s] = csub(t2, t3);
}

// final radix-5 stage (n == 5, p == 0, no twiddles): work item t < s = N / 5
template <bool INV>
JD_HD void stage_radix5_last(const float2* x, float2* y, int s, int t) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f;  // cos(2 pi / 5), cos(4 pi / 5)
  const float s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;   // sin(2 pi / 5), sin(4 pi / 5)
  const float2 a = x[t], b = x[t + s], c = x[t + 2 * s], d = x[t + 3 * s], e = x[t + 4 * s];
  const float2 t1 = cadd(b, e), t2 = cadd(c, d), t3 = csub(b, e), t4 = csub(c, d);
  const float2 m1 = make_float2(a.x + c1 * t1.x + c2 * t2.x, a.y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(a.x + c2 * t1.x + c1 * t2.x, a.y + c2 * t1.y + c1 * t2.y);
  const float2 n1 = rot<INV>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));  // -+ i n1
  const float2 n2 = rot<INV>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
  y[t] = cadd(a, cadd(t1, t2));
  y[t + s] = cadd(m1, n1);
  y[t + 2 * s] = cadd(m2, n2);
  y[t + 3 * s] = csub(m2, n2);
  y[t + 4 * s] = csub(m1, n1);
}

// Whole transform of length N = r0 * 2^L, ping-pong between x and y; returns the buffer holding the result (natural
// order).  `for_each(items, f)` runs f(t) for every work item t < items and then synchronises the workers (thread-strided
// loop + __syncthreads() in the kernels, a plain loop on the host).
template <bool INV, class ForEach>
JD_HD float2* fft_mixed(float2* x, float2* y, const float2* tw, int N, int L, int r0, ForEach&& for_each) {
  int n = N, ls = 0, done = 0;  // stride s = 1 << ls, n = N >> done
  while (L - done >= 2) {
    for_each(N / 4, [&](int t) { stage_radix4<INV>(x, y, tw, n, ls, done, t); });
    float2* tmp = x;
    x = y, y = tmp;
    n >>= 2, ls += 2, done += 2;
  }
  if (L - done == 1) {
    for_each(N / 2, [&](int t) { stage_radix2<INV>(x, y, tw, n, ls, done, t); });
    float2* tmp = x;
    x = y, y = tmp;
    n >>= 1, ls += 1, done += 1;
  }
  if (r0 == 3) {
    for_each(N / 3, [&](int t) { stage_radix3_last<INV>(x, y, N / 3, t); });
    float2* tmp = x;
    x = y, y = tmp;
  } else if (r0 == 5) {
    for_each(N / 5, [&](int t) { stage_radix5_last<INV>(x, y, N / 5, t); });
    float2* tmp = x;
    x = y, y = tmp;
  }
  return x;
}

// smallest r0 * 2^L >= v with r0 in {1, 3, 5} and L >= 1 (mixed) or the next power of two (!mixed)
inline int fft_size(int v, bool mixed, int* L, int* r0) {
  int best = 0, bl = 0, br = 1;
  const int radices[3] = {1, 3, 5};
  for (int i = 0; i < (mixed ? 3 : 1); ++i) {
    int n = 2 * radices[i], l = 1;
    while (n < v) {
      n <<= 1;
      ++l;
    }
    if (best == 0 || n < best) best = n, bl = l, br = radices[i];
  }
  *L = bl;
  *r0 = br;
  return best;
}

}  // namespace fft
}  // namespace jd
