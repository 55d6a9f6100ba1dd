"""ctypes binding of libddpmir.so (the C ABI declared in include/ddpmir.h).

There is no CPU fallback: importing this module without the built library raises, and every call checks the
status code and raises with ddpmir_last_error().
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libddpmir.so")

F32, BF16, F16 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LRELU02, ACT_SIGMOID, ACT_SILU, ACT_GELU, ACT_TANH = range(7)
IMPL_AUTO, IMPL_SIMT, IMPL_TENSOR = 0, 1, 2


class Epilogue(ctypes.Structure):
    """Mirror of ddpmir_epilogue_t."""
    _fields_ = [("bias", c_void_p), ("bias2", c_void_p), ("row_bias", c_void_p), ("img_scale", c_void_p),
                ("mul", c_void_p), ("res", c_void_p), ("out2", c_void_p), ("act", c_int), ("freq_mode", c_int),
                ("bs", c_int), ("low", c_int), ("out_dtype", c_int), ("out2_dtype", c_int), ("mul_dtype", c_int),
                ("res_dtype", c_int)]


_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "ddpmir_version": (c_int, []),
    "ddpmir_last_error": (c_char_p, []),
    "ddpmir_ddrm_update": (c_int, [_P, _P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                   c_int, c_uint64, c_uint32, c_uint64, _P]),
    "ddpmir_gmm_update": (c_int, [_P, _P, _P, _P, c_float, _P, _P, c_int64, c_int, c_float, c_int, c_uint64, c_uint32, _P]),
    "ddpmir_lincomb": (c_int, [_P, c_float, _P, c_float, _P, c_float, _P, c_int64, c_uint64, c_uint32, _P]),
    "ddpmir_philox_normal": (c_int, [_P, c_int64, c_uint64, c_uint32, _P]),
    "ddpmir_quantize_u8_hwc": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_u8_hwc_to_nchw": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_phase_reference": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_phase_consistency": (c_int, [_P, _P, c_float, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_svd_lowrank": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P]),
    "ddpmir_color_l1": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_mse": (c_int, [_P, _P, c_int64, _P, _P, _P]),
    "ddpmir_ssim": (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_freq_loss_terms": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "ddpmir_time_embed": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P]),
    "ddpmir_linear_rows": (c_int, [_P, c_int, c_int, _P, _P, c_int, c_int, _P, _P]),
    "ddpmir_groupnorm_stats": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P]),
    "ddpmir_groupnorm_apply": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, _P, c_int, _P, _P]),
    "ddpmir_conv_input": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "ddpmir_conv3x3": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, ctypes.POINTER(Epilogue), _P, c_int, _P]),
    "ddpmir_gemm": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, ctypes.POINTER(Epilogue), _P, c_int, _P]),
    "ddpmir_attention": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "ddpmir_channel_scale_add": (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P, _P, c_int, _P]),
    "ddpmir_jpeg_roundtrip_workspace": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "ddpmir_jpeg_roundtrip_u8": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_jpeg_dct_project": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, _P]),
    "ddpmir_attention_prescaled_workspace": (ctypes.c_size_t, [c_int, c_int, c_int]),
    "ddpmir_attention_prescaled": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_attention_prescaled_f16_workspace": (ctypes.c_size_t, [c_int, c_int, c_int, c_int]),
    "ddpmir_attention_prescaled_f16": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_block_transform": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "ddpmir_maxpool2": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_upsample2_concat": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_avgpool_pyramid": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_avif_combine": (c_int, [_P, c_int, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_out_conv_tanh": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P]),
    "ddpmir_cast_f32_to_bf16": (c_int, [_P, _P, c_int64, _P]),
    "ddpmir_wgrad": (c_int, [_P, c_int, _P, c_int, _P] + [c_int] * 12 + [_P]),
    "ddpmir_colsum": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_time_features": (c_int, [_P, c_int, c_int, _P, _P]),
    "ddpmir_groupnorm_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P]),
    "ddpmir_gate_backward": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_lrelu_mask_backward": (c_int, [_P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_avif_combine_backward": (c_int, [_P, _P, _P, _P, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "ddpmir_avif_gates_backward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_avgpool_pyramid_backward": (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    "ddpmir_block_transform_wgrad": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "ddpmir_pack_weight": (c_int, [_P, c_int, c_int, c_int, _P, c_int64, c_int64, _P, c_int64, c_int64, c_int, _P]),
    "ddpmir_relu_mask_backward": (c_int, [_P, _P, c_int, _P, c_int64, _P]),
    "ddpmir_dropout": (c_int, [_P, c_int, _P, c_int, c_int64, c_float, c_uint64, _P]),
    "ddpmir_maxpool2_backward": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_upsample2_concat_backward": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_attention_train_forward": (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P]),
    "ddpmir_attention_backward": (c_int, [_P, _P, c_int, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "ddpmir_act_forward": (c_int, [_P, c_int, _P, c_int64, _P]),
    "ddpmir_act_backward": (c_int, [_P, _P, c_int, _P, c_int64, _P]),
    "ddpmir_linear_rows_backward": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, _P, _P]),
    "ddpmir_conv_input_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ddpmir_out_conv_tanh_backward": (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P]),
    "ddpmir_mse_backward": (c_int, [_P, _P, c_int64, c_float, _P, c_int, _P]),
    "ddpmir_freq_loss_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "ddpmir_fft2_loss_terms": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "ddpmir_fft2_loss_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, _P, _P, _P, _P, _P]),
    "ddpmir_edge_loss": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P]),
    "ddpmir_edge_loss_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_float, _P, c_int, _P]),
    "ddpmir_ssim_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P]),
    "ddpmir_huber": (c_int, [_P, _P, c_int64, c_float, _P, _P, _P]),
    "ddpmir_huber_backward": (c_int, [_P, _P, c_int64, c_float, c_float, _P, c_int, _P]),
    "ddpmir_color_l1_backward": (c_int, [_P, _P, c_int, c_int, c_int, c_float, _P, c_int, _P]),
    "ddpmir_sumsq": (c_int, [_P, c_int64, _P, _P]),
    "ddpmir_adamw_multi": (c_int, [_P, _P, _P, c_int, _P, _P, _P, c_float, c_float, c_float, c_float, c_float, c_int, _P, c_float, _P]),
    "ddpmir_adamw_step": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int, _P, c_float, _P]),
}

# not part of the public header: tuning hook used by tests/bench
_PRIVATE = {"ddpmir_attention_set_expmode": (c_int, [c_int]), "ddpmir_attention_set_lin": (c_int, [c_int]), "ddpmir_igemm_set_variant": (c_int, [c_int])}

_lib = None


class DdpmirError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DdpmirError(
                f"{LIB_PATH} is missing: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or `python -m ddpm_image_restoration_b200.build`). There is no CPU fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in {**_SIGNATURES, **_PRIVATE}.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().ddpmir_last_error()
        raise DdpmirError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
