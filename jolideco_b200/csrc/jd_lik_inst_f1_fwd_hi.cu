// Instantiations of the batched likelihood kernel (jd_likelihood.cuh), 4 x 8 outputs per thread: FWD, f = 1, tap groups 5..8.
// One translation unit per slice so that the build compiles them in parallel.
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f1_fwd_hi(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                    float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 16: return launch<FWD, 1, 5, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 17: return launch<FWD, 1, 5, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 18: return launch<FWD, 1, 5, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 19: return launch<FWD, 1, 5, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 20: return launch<FWD, 1, 6, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 21: return launch<FWD, 1, 6, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 22: return launch<FWD, 1, 6, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 23: return launch<FWD, 1, 6, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 24: return launch<FWD, 1, 7, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 25: return launch<FWD, 1, 7, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 26: return launch<FWD, 1, 7, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 27: return launch<FWD, 1, 7, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 28: return launch<FWD, 1, 8, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 29: return launch<FWD, 1, 8, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 30: return launch<FWD, 1, 8, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 31: return launch<FWD, 1, 8, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no kernel for tap-group key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
