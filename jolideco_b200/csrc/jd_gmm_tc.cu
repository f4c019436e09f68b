// GMM patch prior forward: backend dispatch.  backend 1 (tcgen05 split-TF32) lives here.
#include "jd_common.cuh"

namespace jd {
int gmm_prior_forward_simt(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride, int row_begin,
                           int row_end, const float* Lw, const float* mw, const float* ck, int K, int marginalize,
                           float* value, int32_t* argmax, float* logp, double* sum, cudaStream_t st);
}

using namespace jd;

extern "C" int jd_gmm_prior_forward(const float* flux, int fH, int fW, const int32_t* shift_yx, int stride,
                                    int row_begin, int row_end, const float* Lw, const float* mw, const float* ck,
                                    int K, int marginalize, float* value, int32_t* argmax, float* logp, double* sum,
                                    int backend, jd_stream_t stream) {
  JD_CHECK_ARG(flux && Lw && mw && ck && K > 0, "jd_gmm_prior_forward: null pointer");
  if (backend == 0)
    return gmm_prior_forward_simt(flux, fH, fW, shift_yx, stride, row_begin, row_end, Lw, mw, ck, K, marginalize,
                                  value, argmax, logp, sum, to_stream(stream));
  set_error("jd_gmm_prior_forward: backend %d not available", backend);
  return JD_ERR_UNSUPPORTED;
}
