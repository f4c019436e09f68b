// Instantiations of the batched likelihood kernel (jd_likelihood.cuh), 4 x 8 outputs per thread: FWD, f = 1, tap groups 1..4.
// One translation unit per slice so that the build compiles them in parallel.
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f1_fwd_lo(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                    float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 0: return launch<FWD, 1, 1, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 1: return launch<FWD, 1, 1, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 2: return launch<FWD, 1, 1, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 3: return launch<FWD, 1, 1, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 4: return launch<FWD, 1, 2, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 5: return launch<FWD, 1, 2, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 6: return launch<FWD, 1, 2, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 7: return launch<FWD, 1, 2, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 8: return launch<FWD, 1, 3, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 9: return launch<FWD, 1, 3, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 10: return launch<FWD, 1, 3, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 11: return launch<FWD, 1, 3, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 12: return launch<FWD, 1, 4, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 13: return launch<FWD, 1, 4, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 14: return launch<FWD, 1, 4, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 15: return launch<FWD, 1, 4, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no kernel for tap-group key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
