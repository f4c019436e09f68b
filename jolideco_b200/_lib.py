"""ctypes binding of `libjolideco_b200.so` (the C ABI declared in include/jolideco_b200.h).

The library is the only compute path of this package: if it cannot be loaded, or a call fails,
a `JolidecoB200Error` is raised — there is no PyTorch / CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JD_LIB_PATH") or os.path.join(_HERE, "libjolideco_b200.so")

c_f32p = ctypes.c_void_p
c_i32p = ctypes.c_void_p
c_u8p = ctypes.c_void_p
c_f64p = ctypes.c_void_p
c_stream = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_float = ctypes.c_float

# name -> argtypes (restype is int for all but jd_last_error); mirrors include/jolideco_b200.h
PROTOTYPES = {
    "jd_abi_version": [],
    "jd_last_error": [],
    "jd_device_supported": [c_int],
    "jd_flux_forward": [c_f32p, c_u8p, c_f32p, c_i64, c_int, c_stream],
    "jd_conv_forward_direct": [c_f32p, c_f32p, c_f32p, c_f32p, c_int, c_int, c_int, c_int, c_stream],
    "jd_conv_backward_direct": [c_f32p, c_f32p, c_f32p, c_f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_stream],
    "jd_fftconv_sizes": [c_int, c_int, c_int, c_int, ctypes.c_void_p, ctypes.c_void_p],
    "jd_fftconv_prepare_psf": [c_f32p, c_int, c_int, c_int, c_int, c_f32p, c_f32p, c_stream],
    "jd_conv_tuning": [c_int, c_int, c_int],
    "jd_conv_forward_fft": [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_int, c_int, c_int, c_int, c_stream],
    "jd_conv_backward_fft": [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_stream],
    "jd_pool_sum": [c_f32p, c_f32p, c_int, c_int, c_int, c_int, c_stream],
    "jd_adam_step_dev": [c_f32p, c_f32p, c_f32p, c_f32p, c_u8p, c_f32p, c_f32p, c_float, c_int, c_i64, c_f32p,
                         c_float, c_float, c_float, c_stream],
    "jd_step_begin": [c_i32p, c_i32p, c_int, c_i32p, c_int, c_float, c_float, c_float, c_f32p, c_f64p, c_int,
                      c_stream],
    "jd_shift_forward": [c_f32p, c_f32p, c_int, c_int, c_int, c_f32p, c_stream],
    "jd_shift_backward": [c_f32p, c_f32p, c_f32p, c_int, c_int, c_int, c_f32p, c_int, c_f64p, c_stream],
    "jd_adam_scalar_step_dev": [c_f32p, c_f32p, c_f32p, c_f64p, c_i32p, c_int, c_float, c_float, c_float, c_float,
                                c_stream],
    "jd_adam_allreduce_peer": [ctypes.c_void_p, ctypes.c_void_p, c_int, c_int, c_f32p, c_f32p, c_f32p, c_u8p, c_int,
                               c_i64, c_f32p, c_float, c_float, c_float, c_stream],
    "jd_step_begin_flux": [c_i32p, c_i32p, c_int, c_i32p, c_int, c_float, c_float, c_float, c_f32p, c_f64p, c_int,
                           c_f32p, c_u8p, c_f32p, c_i64, c_int, c_stream],
    "jd_adam_fold_step_dev": [c_f32p, c_f32p, c_f32p, c_f32p, c_u8p, c_f32p, c_f32p, c_float, c_int, c_int, c_int,
                              c_i32p, c_int, c_int, c_int, c_f32p, c_float, c_float, c_float, c_stream],
    "jd_poisson_forward_backward": [c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f64p, c_f64p, c_int, c_int,
                                    c_int, c_int, c_float, c_float, c_stream],
    "jd_gmm_log_prob": [c_f32p, c_i64, c_int, c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_stream],
    "jd_extract_patches": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_stream],
    "jd_gmm_prior_forward": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_f32p, c_f32p, c_int,
                             c_int, c_f32p, c_i32p, c_f32p, c_f64p, c_int, c_stream],
    "jd_gmm_tc_packed_bytes": [c_int],
    "jd_gmm_tc_pack": [c_f32p, c_int, ctypes.c_void_p, c_stream],
    "jd_gmm_prior_forward_tc": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                c_int, c_int, c_int, c_int, c_f32p, c_i32p, c_f32p, c_f64p, c_stream],
    "jd_gmm_tc_sk_workspace_bytes": [c_i64, c_int],
    "jd_gmm_tc_sk_plan": [c_i64, c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p],
    "jd_gmm_prior_forward_tc_sk": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                   c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_i32p, c_f32p, c_f64p,
                                   c_stream],
    "jd_gmm_prior_backward_lse_tc": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_int,
                                     c_f32p, c_f32p, c_float, c_f32p, c_stream],
    "jd_gmm_tc16_packed_bytes": [c_int],
    "jd_gmm_tc16_pack": [c_f32p, c_int, ctypes.c_void_p, c_f32p, c_stream],
    "jd_gmm_prior_forward_tc16": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                  c_f32p, c_int, c_int, c_int, c_int, c_f32p, c_i32p, c_f32p, c_f64p, c_stream],
    "jd_likelihood_forward_fft": [ctypes.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                  c_float, c_stream],
    "jd_likelihood_backward_fft": [ctypes.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_stream],
    "jd_gmm_tcm_packed_bytes": [c_int],
    "jd_gmm_tcm_pack": [c_f32p, c_int, ctypes.c_void_p, c_f32p, c_stream],
    "jd_gmm_tcm_workspace_bytes": [c_i64, c_int],
    "jd_gmm_prior_forward_tcm": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                 c_f32p, c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_i32p, c_f32p, c_f64p,
                                 c_stream],
    "jd_gmm_tcm2_workspace_bytes": [c_i64, c_int],
    "jd_gmm_prior_forward_tcm2": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                 c_f32p, c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_i32p, c_f32p, c_f64p,
                                 c_stream],
    "jd_gmm_prior_forward_tc16x2": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_f32p,
                                 c_f32p, c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_i32p, c_f32p, c_f64p,
                                 c_stream],
    "jd_gmm_prior_forward_tcx2_on": [c_int, c_int, c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, ctypes.c_void_p, c_f32p,
                                     c_f32p, c_f32p, c_int, c_int, c_int, c_int, ctypes.c_void_p, c_f32p, c_i32p, c_f32p,
                                     c_f64p, c_stream],
    "jd_gmm_backward_workspace_elems": [c_i64, c_int],
    "jd_gmm_prior_backward": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_f32p, c_int, c_int,
                              c_i32p, c_f32p, c_f32p, c_float, c_f32p, c_i32p, c_stream],
    "jd_gmm_prior_backward_max_tri": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_f32p, c_int, c_i32p,
                                      c_float, c_f32p, c_stream],
    "jd_patch_fold": [c_f32p, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_int, c_stream],
    "jd_likelihood_supported": [c_int, c_int, c_int],
    "jd_likelihood_forward": [ctypes.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                              c_stream],
    "jd_likelihood_backward": [ctypes.c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_stream],
    "jd_probe_fp32_fma": [c_int, c_f32p, c_stream],
    "jd_adam_joint_step_dev": [c_f32p, c_f32p, c_f32p, c_f32p, c_u8p, c_f32p, c_int, c_i64, c_f32p, c_float, c_int,
                               c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p, c_float, c_float, c_float, c_stream],
    "jd_grad_reduce_local": [c_f32p, c_int, c_i64, c_f32p, c_float, c_int, c_int, c_i32p, c_int, c_int, c_int, c_f32p,
                             c_stream],
    "jd_adam_allreduce_peer_sync": [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, c_int, c_int,
                                    c_f32p, c_f32p, c_f32p, c_u8p, c_int, c_i64, c_f32p, c_float, c_float, c_float,
                                    c_stream],
    "jd_adam_step": [c_f32p, c_f32p, c_f32p, c_f32p, c_u8p, c_f32p, c_f32p, c_float, c_int, c_i64, c_int, c_float,
                     c_float, c_float, c_float, c_stream],
}


class JolidecoB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once) and attach prototypes. Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JolidecoB200Error(
            f"{LIB_PATH} not found: build it with `python -m jolideco_b200.build` "
            "(nvcc, sm_100a). jolideco_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = {"jd_last_error": ctypes.c_char_p, "jd_gmm_tc_packed_bytes": ctypes.c_size_t,
                      "jd_gmm_tc16_packed_bytes": ctypes.c_size_t,
                      "jd_gmm_tcm_packed_bytes": ctypes.c_size_t,
                      "jd_gmm_tcm_workspace_bytes": ctypes.c_int64,
                      "jd_gmm_tcm2_workspace_bytes": ctypes.c_int64,
                      "jd_gmm_backward_workspace_elems": ctypes.c_int64,
                      "jd_gmm_tc_sk_workspace_bytes": ctypes.c_int64,
                      "jd_probe_fp32_fma": ctypes.c_int64}.get(
            name, ctypes.c_int)
    if lib.jd_abi_version() != 1:
        raise JolidecoB200Error(f"ABI version mismatch: library {lib.jd_abi_version()}, binding 1")
    _lib = lib
    return lib


def call(name, *args):
    """Call an entry point; raise JolidecoB200Error with jd_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.jd_last_error()
        raise JolidecoB200Error(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")
    return rc
