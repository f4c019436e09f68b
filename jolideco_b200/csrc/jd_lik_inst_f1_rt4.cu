// Instantiations of the batched likelihood kernel (jd_likelihood.cuh) with 4 x 8 outputs per thread (32 x 64 tiles):
// twice the CTAs / warps of the 8 x 8 variant for launches that would leave the SMs under-filled (one 1024^2 dataset is
// 256 CTAs of 2 warps on 148 SMs).  f = 1, tap rows of 17..20 taps (KG = 5), both directions.  key = 4 * mode + KT - 1.
#include "jd_likelihood.cuh"

namespace jd {
namespace lik {

int dispatch_f1_rt4(int key, const jd_lik_dataset* table, int n_datasets, int fH, int fW, int kh, int kw, int H, int W,
                    float eps, float grad_scale, cudaStream_t st) {
  switch (key) {
    case 0: return launch<FWD, 1, 5, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 1: return launch<FWD, 1, 5, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 2: return launch<FWD, 1, 5, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 3: return launch<FWD, 1, 5, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 4: return launch<BWD, 1, 5, 1, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 5: return launch<BWD, 1, 5, 2, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 6: return launch<BWD, 1, 5, 3, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
    case 7: return launch<BWD, 1, 5, 4, 4>(table, n_datasets, fH, fW, kh, kw, H, W, eps, grad_scale, st);
  }
  set_error("jd_likelihood: no 4 x 8 kernel for key %d", key);
  return JD_ERR_UNSUPPORTED;
}

}  // namespace lik
}  // namespace jd
