// Instantiations of the batched likelihood kernel (jd_likelihood.cuh) for upsampling factor 2: tap rows padded to
// whole groups of four (KT = 4), up to 10 groups (PSF rows of <= 37..40 taps), both directions.
// key = 16 * mode + (KG - 1).
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f2(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 0: return launch<FWD, 2, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 1: return launch<FWD, 2, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 2: return launch<FWD, 2, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 3: return launch<FWD, 2, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 4: return launch<FWD, 2, 5, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 5: return launch<FWD, 2, 6, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 6: return launch<FWD, 2, 7, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 7: return launch<FWD, 2, 8, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 8: return launch<FWD, 2, 9, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 9: return launch<FWD, 2, 10, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 16: return launch<BWD, 2, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 17: return launch<BWD, 2, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 18: return launch<BWD, 2, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 19: return launch<BWD, 2, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 20: return launch<BWD, 2, 5, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 21: return launch<BWD, 2, 6, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 22: return launch<BWD, 2, 7, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 23: return launch<BWD, 2, 8, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 24: return launch<BWD, 2, 9, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 25: return launch<BWD, 2, 10, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no f = 2 kernel for key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
