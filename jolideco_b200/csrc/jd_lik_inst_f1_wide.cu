// Instantiations of the batched likelihood kernel (jd_likelihood.cuh), f = 1, PSF rows of 33..40 taps (9 or 10 tap
// groups, padded: KT = 4), both directions.  key = 16 * mode + (KG - 1).
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f1_wide(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                     float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 8: return launch<FWD, 1, 9, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 9: return launch<FWD, 1, 10, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 24: return launch<BWD, 1, 9, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 25: return launch<BWD, 1, 10, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no wide f = 1 kernel for key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
