"""ctypes binding of libshmgan.so (the C ABI declared in include/shmgan.h).

There is NO fallback: if the library is missing, or a call returns a non-zero status, this raises.
PyTorch tensors only carry the device pointers; nothing here computes.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libshmgan.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3


class ConvDesc(C.Structure):
    """shm_conv_desc (include/shmgan.h)."""
    _fields_ = [(n, C.c_int32) for n in ("N", "H", "W", "Cin", "Cout", "kh", "kw", "stride", "transposed", "act",
                                         "ldx", "ldy", "dtype", "tensor_core")]


class ShmError(RuntimeError):
    pass


_P, _I, _L, _F, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64
_D = C.POINTER(ConvDesc)

# name -> argtypes (return type int unless listed in _RET)
_SIG = {
    "shm_conv2d_fwd": [_D, _P, _P, _P, _P, _P],
    "shm_conv2d_dgrad": [_D, _P, _P, _P, _I, _P],
    "shm_conv2d_wgrad": [_D, _P, _P, _P, _P, _P],
    "shm_conv2d_tc_supported": [_D, _I],
    "shm_conv2d_tc_prep_weights": [_D, _P, _I, _P, _I, _P],
    "shm_conv2d_tc_route": [_D, _I],
    "shm_conv2d_tc_prep_weights_both": [_D, _P, _I, _P, _P, _P],
    "shm_conv2d_tc_prep_weights_padded": [_D, _P, _I, _I, _I, _I, _P, _P],
    "shm_conv2d_tc_prep_job_bytes": [],
    "shm_conv2d_tc_prep_job": [_D, _P, _I, _P, _P, _P],
    "shm_conv2d_tc_prep_jobs_finalize": [_P, _I],
    "shm_conv2d_tc_prep_multi": [_P, _I, _I, _P],
    "shm_conv2d_tc_fwd": [_D, _P, _P, _P, _P, _P],
    "shm_conv2d_tc_fwd_cols": [_D, _P, _P, _P, _P, _I, _P],
    "shm_conv2d_tc_fwd_stats": [_D, _P, _P, _P, _P, _P, _P],
    "shm_conv2d_tc_stats_supported": [_D],
    "shm_conv2d_tc_dgrad": [_D, _P, _P, _P, _P],
    "shm_conv2d_tc_wgrad": [_D, _P, _P, _P, _P],
    "shm_colsum": [_P, _L, _I, _I, _I, _P, _P],
    "shm_inorm_stats": [_P, _I, _I, _I, _I, _I, _P, _P],
    "shm_inorm_apply": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _F, _P, _I, _I, _P, _I, _P, _I, _P],
    "shm_inorm_bwd_stats": [_P, _I, _I, _I, _I, _I, _I, _P, _F, _P, _I, _P, _I, _P, _P],
    "shm_inorm_bwd_apply": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _F, _P, _I, _P, _I, _P, _I, _P, _I, _P, _P],
    "shm_norm_tune": [_I, _I, _I],
    "shm_act_bwd": [_P, _I, _P, _I, _P, _I, _L, _I, _I, _I, _P, _P],
    "shm_maxpool": [_P, _I, _I, _I, _I, _I, _I, _P, _I, _I, _P],
    "shm_bn_eval": [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _P, _I, _P, _I, _P],
    "shm_add": [_P, _I, _P, _I, _P, _I, _L, _I, _I, _P],
    "shm_group_sum": [_P, _I, _I, _L, _I, _P, _I, _I, _I, _P],
    "shm_mul_mask": [_P, _P, _P, _L, _F, _I, _P],
    "shm_rng_normal": [_P, _L, _U64, _U64, _F, _I, _P],
    "shm_rng_keep": [_P, _L, _U64, _U64, _F, _I, _P],
    "shm_rng_normal_dev": [_P, _L, _U64, _P, _F, _I, _P],
    "shm_rng_keep_dev": [_P, _L, _U64, _P, _F, _I, _P],
    "shm_cast": [_P, _I, _P, _I, _L, _P],
    "shm_cast2d": [_P, _I, _I, _P, _I, _I, _L, _I, _P],
    "shm_axpy": [_F, _P, _P, _L, _I, _P],
    "shm_dense_fwd": [_P, _P, _P, _I, _I, _I, _I, _P],
    "shm_dense_dgrad": [_P, _P, _P, _I, _I, _I, _I, _P],
    "shm_dense_wgrad": [_P, _P, _P, _I, _I, _I, _I, _P],
    "shm_pseudo_diffuse_min4": [_P, _P, _P, _P, _P, _L, _I, _P],
    "shm_yuv_stats": [_P, _I, _I, _P, _P],
    "shm_yuv_standardize": [_P, _I, _I, _P, _P, _P, _P],
    "shm_avg_cbcr": [_P, _P, _P, _P, _P, _P, _L, _P],
    "shm_assemble_input": [C.POINTER(_P), C.POINTER(C.c_int32), _I, _P, _I, _L, _I, _P],
    "shm_pad_channels64": [_P, _I, _I, _I, _P, _L, _P],
    "shm_pad_channels": [_P, _I, _I, _I, _P, _I, _L, _P],
    "shm_im2col_k3s2": [_P, _I, _I, _I, _I, _I, _I, _P, _P],
    "shm_col2im_k3s2": [_P, _I, _I, _I, _I, _P, _I, _I, _P],
    "shm_assemble_bwd": [_P, _I, _I, C.POINTER(C.c_int32), _I, _P, _L, _P],
    "shm_yuv2rgb": [_P, _P, _L, _P, _P, _I, _I, _L, _P],
    "shm_yuv2rgb_bwd": [_P, _P, _I, _I, _P, _L, _I, _P],
    "shm_sum_count": [_P, _L, _P, _P],
    "shm_scale_by_mean": [_P, _P, _F, _P, _L, _P],
    "shm_pw1_fwd": [_P, _I, _I, _P, _P, _I, _P, _L, _I, _P],
    "shm_pw1_bwd": [_P, _I, _I, _P, _P, _P, _I, _P, _I, _P, _P, _L, _I, _P],
    "shm_c3to1_fwd": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P],
    "shm_c3to1_dgrad": [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    "shm_c3to1_wgrad": [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P],
    "shm_lsgan": [_P, _L, _F, _P, _F, _P, _F, _I, _P],
    "shm_softmax_ce": [_P, _I, C.POINTER(C.c_float), _P, _F, _P, _F, _I, _P],
    "shm_lsgan_dev": [_P, _L, _P, _P, _F, _P, _F, _I, _P],
    "shm_softmax_ce_dev": [_P, _I, _P, _P, _F, _P, _F, _I, _P],
    "shm_l1": [_P, _P, _L, _P, _F, _P, _F, _I, _P],
    "shm_mse": [_P, _P, _L, _P, _F, _P, _F, _I, _P],
    "shm_mse_ycc": [_P, _P, _P, _L, _P, _F, _P, _F, _P],
    "shm_minmax3": [_P, _P, _I, _I, _P, _P, _P],
    "shm_gram3": [_P, _P, _I, _I, _P, _P],
    "shm_style_loss": [_P, _P, _I, _I, _I, _P, _F, _P, _F, _P],
    "shm_gram3_bwd": [_P, _P, _I, _I, _P, _P, _P],
    "shm_ssim_fwd": [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P],
    "shm_ssim_loss": [_P, _I, _P, _F, _P, _F, _P],
    "shm_ssim_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
    "shm_spec_loss": [_P, _P, _P, _P, _L, _P, _F, _P],
    "shm_clip_adam": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _F, _P],
    "shm_clip_adam_dev": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _F, _P],
    "shm_load_u8_bilinear": [_P, _I, _I, _I, _P, _I, _I, _I, _P],
    "shm_dop": [_P, _P, _P, _P, _P, _P, _L, _P],
    "shm_sqerr_per_image": [_P, _P, _I, _L, _P, _P],
    "shm_delta_e": [_P, _P, _I, _L, _P, _P],
    "shm_conv2d_tc_weight_elems": [_D],
    "shm_ssim_map_elems": [_I, _I, _I],
    "shm_zero": [_P, _L, _P],
    "shm_last_error": [],
    "shm_version": [],
    "shm_sm_count": [],
}
_RET = {"shm_last_error": C.c_char_p, "shm_conv2d_tc_weight_elems": C.c_int64, "shm_ssim_map_elems": C.c_int64}
# functions whose int return is a value, not a status
_VALUE = {"shm_conv2d_tc_route", "shm_conv2d_tc_supported", "shm_conv2d_tc_stats_supported", "shm_version", "shm_sm_count", "shm_conv2d_tc_weight_elems", "shm_ssim_map_elems",
          "shm_conv2d_tc_prep_job_bytes", "shm_conv2d_tc_prep_jobs_finalize",
          "shm_last_error"}

EXPORTS = tuple(_SIG)

_lib = None
_launches = 0


def load():
    """Loads libshmgan.so; raises if it has not been built (`python -m shmgan_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShmError("%s is missing: build it with `python -m shmgan_b200.build` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in _SIG.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RET.get(name, C.c_int)
    _lib = lib
    return lib


def launches() -> int:
    """Number of kernel-launching C-ABI calls made so far (bench.py's `gpu_launches` evidence)."""
    return _launches


def count_replayed(n: int):
    """A CUDA-graph replay re-launches the n kernels recorded at capture without passing through call()."""
    global _launches
    _launches += int(n)


def call(name: str, *args):
    """Invokes a status-returning entry point and raises ShmError with shm_last_error() on failure."""
    global _launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _VALUE:
        return rc
    if name != "shm_zero":                      # a memset node, not a kernel: kept out of the kernel-launch count bench.py reports
        _launches += 1
    if rc != 0:
        raise ShmError("%s failed (%d): %s" % (name, rc, lib.shm_last_error().decode()))
    return rc
