// Instantiations of the batched likelihood kernel (jd_likelihood.cuh) with 8 x 8 outputs per thread (64 x 64 tiles) for
// f = 1 and tap rows of 17..20 taps (KG = 5), both directions: the A/B partner (JD_LIK_RT=8) of the 4 x 8 kernels, which
// measured faster at the north-star shape for one dataset (fwd 39.1 -> 29.7 us, adjoint 29.2 -> 23.0 us: 256 CTAs of 2
// warps leave the 148 SMs nearly empty) and for eight (step 505 -> 485 us: half the unrolled loop body, twice the warps).
// key = 4 * mode + KT - 1.
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f1_rt8(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                    float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 0: return launch<FWD, 1, 5, 1, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 1: return launch<FWD, 1, 5, 2, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 2: return launch<FWD, 1, 5, 3, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 3: return launch<FWD, 1, 5, 4, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 4: return launch<BWD, 1, 5, 1, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 5: return launch<BWD, 1, 5, 2, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 6: return launch<BWD, 1, 5, 3, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 7: return launch<BWD, 1, 5, 4, 8>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no 8 x 8 kernel for key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
