"""ctypes binding of libducosy_sm100.so (the C ABI declared in include/ducosy.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libducosy_sm100.so")

F16, BF16, F16X2 = 0, 1, 2   # F16X2: split-operand (hi, lo) fp16 pairs, generator inference only (include/ducosy.h)
PAD_ZERO, PAD_REFLECT = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU02 = 0, 1, 2


class GenConfig(C.Structure):
    _fields_ = [("input_channels", C.c_int), ("num_residual_blocks", C.c_int), ("use_cbam", C.c_int),
                ("dtype", C.c_int)]


_p, _i, _f, _ll, _sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t

# name -> (restype, argtypes); mirrors include/ducosy.h one to one (tests check every symbol resolves)
SIGNATURES = {
    "ducosy_version": (_i, []),
    "ducosy_last_error": (C.c_char_p, []),
    "ducosy_check_device": (_i, []),
    "ducosy_hu_window": (_i, [_p, _p, _p, _ll, _f, _f, _f, _f, _f, _f, _p]),
    "ducosy_hu_window_soft": (_i, [_p, _p, _ll, _f, _f, _f, _f, _f, _p]),
    "ducosy_apply_windowing": (_i, [_p, _p, _ll, _f, _f, _f, _f, _p]),
    "ducosy_hu_thresholds": (_i, [_p, _p, _p, _p, _ll, _f, _f, _p]),
    "ducosy_dewindow_composite": (_i, [_p, _p, _p, _p, _p, _p, _p, _ll, _f, _f, _f, _f, _f, _f, _p]),
    "ducosy_pack_conv_weight": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_upconv_weight": (_i, [_p, _p, _i, _i, _i, _p]),
    "ducosy_pack_upconv_merged_weight": (_i, [_p, _p, _i, _i, _i, _p]),
    "ducosy_upconv2x_merged_nhwc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_stem_weight": (_i, [_p, _p, _i, _i, _p]),
    "ducosy_conv2d_nhwc": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_conv2d_nhwc_in": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_upconv2x_nhwc_in": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_upconv2x_merged_nhwc_in": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_upconv2x_nhwc": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_stem_im2col": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_stem_im2col_hu": (_i, [_p, _p, _i, _i, _i, _f, _f, _f, _f, _i, _p]),
    "ducosy_stem_prepare": (_i, [_p, _p, _f, _f, _f, _f, _p, _i, _i, _i, _i, _p]),
    "ducosy_stem_fused": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_in_finalize": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _i, _i, _p]),
    "ducosy_in_apply_pad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_cbam_pool": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_cbam_spatial_conv": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "ducosy_residual_apply_pad": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_residual_cbam_apply_pad": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_out_weight": (_i, [_p, _p, _i, _p]),
    "ducosy_out_conv7x7_tanh": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_out_conv7x7_tanh_fused": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_generator_num_params": (_i, [C.POINTER(GenConfig)]),
    "ducosy_generator_packed_bytes": (_sz, [C.POINTER(GenConfig)]),
    "ducosy_generator_workspace_bytes": (_sz, [C.POINTER(GenConfig), _i, _i, _i]),
    "ducosy_generator_pack": (_i, [C.POINTER(GenConfig), C.POINTER(_p), _i, _p, _p]),
    "ducosy_generator_forward": (_i, [C.POINTER(GenConfig), _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "ducosy_generator_forward_hu": (_i, [C.POINTER(GenConfig), _p, _p, _f, _f, _f, _f, _p, _i, _i, _i, _p, _sz, _p]),
    "ducosy_generator_num_launches": (_i, [C.POINTER(GenConfig)]),
    "ducosy_loss_scratch_bytes": (_sz, []),
    "ducosy_loss_l1_forward": (_i, [_p, _p, _ll, _p, _p, _p]),
    "ducosy_loss_l1_backward": (_i, [_p, _p, _ll, _p, _p, _p]),
    "ducosy_loss_mse_const_forward": (_i, [_p, _f, _ll, _p, _p, _p]),
    "ducosy_loss_mse_const_backward": (_i, [_p, _f, _ll, _p, _p, _p]),
    "ducosy_loss_gradient_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ducosy_loss_gradient_backward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p]),
    "ducosy_loss_contrast_attention_forward": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _f, _p, _p, _p, _p]),
    "ducosy_loss_contrast_attention_backward": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "ducosy_loss_contrast_region_forward": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _p, _p, _p]),
    "ducosy_loss_contrast_region_backward": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _p, _p, _p]),
    "ducosy_loss_contrast_edge_forward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "ducosy_loss_contrast_edge_backward": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "ducosy_loss_ssim_forward": (_i, [_p, _p, _i, _i, _i, _f, _p, _p, _p, _p, _p]),
    "ducosy_loss_ssim_backward": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "ducosy_conv2d_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "ducosy_conv2d_wgrad_nhwc": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _i, _p]),
    "ducosy_conv2d_wgrad_nhwc_oihw": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _i, _p]),
    "ducosy_conv2d_wgrad_nhwc_oihw_acc": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _i, _p]),
    "ducosy_in_backward_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "ducosy_in_backward_pad": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_in_backward_pad_folded": (_i, [_p, _i, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_dgrad_s2_weight": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "ducosy_convs2_dgrad_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_dgrad_s1_weight": (_i, [_p, _p, _i, _i, _i, _p]),
    "ducosy_conv3x3s1_dgrad_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pad_fold": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pad_fold_add": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_pack_upconv_dgrad_weight": (_i, [_p, _p, _i, _i, _i, _p]),
    "ducosy_upconv2x_dgrad_nhwc": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_upsample2x_pad": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_grad_scale": (_i, [_p, _ll, _p, _p]),
    "ducosy_out_conv_backward_scratch_bytes": (_sz, [_i, _i, _i]),
    "ducosy_out_conv_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_stem_col2im": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_unpack_stem_wgrad": (_i, [_p, _p, _i, _i, _p, _p]),
    "ducosy_add_inplace": (_i, [_p, _p, _ll, _i, _p]),
    "ducosy_upconv2x_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "ducosy_upconv2x_wgrad_nhwc": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _i, _p, _sz, _i, _p]),
    "ducosy_unpack_upconv_wgrad": (_i, [_p, _p, _i, _i, _p, _p]),
    "ducosy_postprocess_scratch_bytes": (_sz, [_i, _i, _i]),
    "ducosy_postprocess_minmax_offset_bytes": (_sz, [_i, _i, _i]),
    "ducosy_metrics_chunks": (_i, [_ll]),
    "ducosy_metrics_slice_stats": (_i, [_p, _p, _i, _i, _ll, _p, _p, _p]),
    "ducosy_metrics_ed": (_i, [_p, _p, _i, _i, _ll, _p, _p, _p, _p]),
    "ducosy_metrics_normalize": (_i, [_p, _i, _p, _ll, _p, _p]),
    "ducosy_metrics_emd_i16": (_i, [_p, _p, _i, _ll, _i, _i, _p, _p, _p]),
    "ducosy_metrics_ts": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "ducosy_metrics_ssim_tiles": (_i, [_i, _i]),
    "ducosy_metrics_ssim": (_i, [_p, _p, _i, _i, _i, _i, C.c_double, _p, _p, _p]),
    "ducosy_postprocess_volume": (_i, [_p, _p, _p, _i, _i, _i, _p, _i, _p, _i, _p, _i, C.c_double, _f, _i, _i, _i, _p]),
    "ducosy_adam_step": (_i, [_p, _p, _p, _p, _ll, _f, _f, _f, _f, _i, _p]),
    "ducosy_adam_advance": (_i, [_p, _p]),
    "ducosy_adam_step_dev": (_i, [_p, _p, _p, _p, _ll, _p, _f, _f, _f, _p]),
    "ducosy_channel_attention_scratch_bytes": (_sz, [_i, _i]),
    "ducosy_channel_attention_nchw": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_spatial_attention_scratch_bytes": (_sz, [_i, _i, _i]),
    "ducosy_spatial_attention_nchw": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_nchw_to_nhwc_pad": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ducosy_nhwc_to_nchw": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "ducosy_masks_scratch_bytes": (_sz, [_i, _i, _i]),
    "ducosy_label4": (_i, [_p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "ducosy_binary_fill_holes": (_i, [_p, _p, _i, _i, _i, _p, _sz, _p]),
    "ducosy_detect_lung": (_i, [_p, _p, _i, _i, _i, _f, _f, _i, _i, _p, _sz, _p]),
    "ducosy_detect_lung_vessels": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "ducosy_lung_hull": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _sz, _p]),
    "ducosy_detect_mediastinum": (_i, [_p, _p, _p, _i, _i, _i, _f, _f, _p, _sz, _p]),
    "ducosy_detect_bone": (_i, [_p, _p, _p, _i, _i, _i, _f, _i, _p, _sz, _p]),
    "ducosy_adam_multi_step": (_i, [_p, _p, _i, _p, _f, _f, _f, _i, _p]),
    "ducosy_cbam_channel_train": (_i, [_p] * 9 + [_i, _i, _p]),
    "ducosy_cbam_backward_scratch_bytes": (_sz, [_i, _i, _i, _i]),
    "ducosy_cbam_backward": (_i, [_p] * 20 + [_i, _i, _i, _i, _i, _p]),
    "ducosy_cbam_backward_acc": (_i, [_p] * 20 + [_i, _i, _i, _i, _i, _p]),
    "ducosy_unpack_wgrad": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "ducosy_disc_last_backward_scratch_bytes": (_sz, []),
    "ducosy_disc_last_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_disc_first_backward_scratch_bytes": (_sz, [_i, _i, _i]),
    "ducosy_disc_first_backward": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ducosy_discriminator_backward_workspace_bytes": (_sz, [_i, _i, _i]),
    "ducosy_discriminator_backward": (_i, [_p, _p, _p, _p, C.POINTER(_p), _p, _i, _i, _i, _p, _sz, _i, _p]),
    "ducosy_discriminator_packed_bytes": (_sz, []),
    "ducosy_discriminator_workspace_bytes": (_sz, [_i, _i, _i]),
    "ducosy_discriminator_pack": (_i, [C.POINTER(_p), _i, _p, _i, _p]),
    "ducosy_discriminator_forward": (_i, [_p, _p, _p, _i, _i, _i, _p, _sz, _i, _p]),
}

_lib = None


def load():
    """Load the shared library (built by ducosy_gan_b200/build.py).  Raises if absent: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m ducosy_gan_b200.build` "
                "(ducosy_gan_b200 has no CPU / PyTorch fallback path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class DucosyError(RuntimeError):
    pass


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().ducosy_last_error().decode(errors="replace")
        raise DucosyError(f"{what or 'libducosy'} failed ({rc}): {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def dtype_code(dtype) -> int:
    if dtype in (torch.float16, "fp16", "f16", F16):
        return F16
    if dtype in (torch.bfloat16, "bf16", BF16):
        return BF16
    if dtype in ("fp16x2", "f16x2", "split"):
        return F16X2
    raise ValueError(f"unsupported operand dtype {dtype!r} (fp16, bf16, or fp16x2 for the generator's inference path)")


def torch_dtype(code: int):
    return torch.float16 if code == F16 else torch.bfloat16


def call(name: str, *args):
    check(getattr(load(), name)(*args), name)
