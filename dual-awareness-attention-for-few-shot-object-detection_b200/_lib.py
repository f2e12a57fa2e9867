"""ctypes binding of libdana_b200.so (the C ABI declared in include/dana_b200.h).

The library is the only compute path of this package: if it is missing or fails to load the
import raises -- there is no CPU or eager fallback."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdana_b200.so")

DANA_OK = 0
DANA_EINVAL = -1
DANA_ECUDA = -2
DANA_EDEVICE = -3
DANA_ENOTSUP = -4


class DanaError(RuntimeError):
    pass


class ConvGemmArgs(Structure):
    """Mirror of `dana_conv_gemm_args` (include/dana_b200.h)."""
    _fields_ = [
        ("a_hi", c_void_p), ("a_lo", c_void_p),
        ("a_c", c_int64), ("a_w", c_int64), ("a_h", c_int64), ("a_n", c_int64),
        ("a_sx", c_int64), ("a_sy", c_int64), ("a_sn", c_int64),
        ("taps_r", c_int32), ("taps_s", c_int32), ("pad_y", c_int32), ("pad_x", c_int32),
        ("b_hi", c_void_p), ("b_lo", c_void_p),
        ("b_pitch", c_int64), ("b_batch_stride", c_int64),
        ("n_out", c_int32),
        ("tile_w", c_int32), ("tile_h", c_int32), ("tile_n", c_int32),
        ("out_w", c_int32), ("out_h", c_int32), ("out_n", c_int32),
        ("o_sx", c_int64), ("o_sy", c_int64), ("o_sn", c_int64),
        ("out_hi", c_void_p), ("out_lo", c_void_p), ("out_f32", c_void_p),
        ("scale", c_void_p), ("bias", c_void_p), ("bias_sn", c_int64),
        ("res_hi", c_void_p), ("res_lo", c_void_p), ("res_f32", c_void_p),
        ("r_sx", c_int64), ("r_sy", c_int64), ("r_sn", c_int64),
        ("alpha", c_float), ("relu", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64), ("sk_epoch", c_int32),
        ("softmax_ns", c_int32), ("softmax_pitch", c_int32),
        ("ab_f16", c_int32), ("io_f16", c_int32),
    ]


class CisaArgs(Structure):
    """Mirror of `dana_cisa_args` (include/dana_b200.h)."""
    _fields_ = [
        ("q_hi", c_void_p), ("q_lo", c_void_p), ("q_pitch", c_int64),
        ("s_hi", c_void_p), ("s_lo", c_void_p),
        ("batch", c_int32), ("nq", c_int32), ("sets", c_int32), ("shots", c_int32), ("ns", c_int32), ("c", c_int32),
        ("d", c_int32),
        ("pe", c_void_p),
        ("wq_hi", c_void_p), ("wq_lo", c_void_p), ("wk_hi", c_void_p), ("wk_lo", c_void_p),
        ("un_w", c_void_p), ("un_b", c_void_p), ("unary_gamma", c_float),
        ("ba_w", c_void_p), ("ba_b", c_void_p), ("gamma", c_float),
        ("out_hi", c_void_p), ("out_lo", c_void_p), ("out_pitch", c_int64), ("out_f16", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
    ]


class ConvBwdArgs(ctypes.Structure):
    """dana_conv_bwd_args (include/dana_b200.h)."""
    _fields_ = [
        ("batch", c_int32), ("height", c_int32), ("width", c_int32), ("in_channels", c_int32), ("out_channels", c_int32),
        ("ksize", c_int32), ("stride", c_int32),
        ("grad_out", c_void_p), ("relu_out", c_void_p), ("x_hi", c_void_p), ("x_lo", c_void_p),
        ("x_sn", c_int64), ("x_sy", c_int64), ("x_sx", c_int64),
        ("wd_hi", c_void_p), ("wd_lo", c_void_p), ("scale", c_void_p),
        ("dx", c_void_p), ("dw", c_void_p), ("dres", c_void_p), ("dw_accumulate", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("gemm_workspace", c_void_p), ("gemm_workspace_bytes", c_int64), ("sk_epoch", c_int32),
    ]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of the header.
SIGNATURES = {
    "dana_abi_version": (c_int, []),
    "dana_error_string": (c_char_p, [c_int]),
    "dana_last_cuda_error": (c_int, []),
    "dana_device_error": (c_int, []),
    "dana_nms_workspace_bytes": (c_int64, [c_int]),
    "dana_nms": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_proposals_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "dana_proposals": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_detections_workspace_bytes": (c_int64, [c_int, c_int]),
    "dana_detections": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_float,
                                c_float, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_roi_align_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int]),
    "dana_roi_align_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_roi_align_head": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dana_roi_align_backward_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dana_roi_align_backward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                        c_int, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_episode_resize": (c_int, [c_void_p, c_int, c_int, c_int, c_int64, c_int, c_int, c_int, c_int, c_double, c_double,
                                    c_int, c_int, c_float, c_float, c_float, c_void_p, c_int, c_int, c_void_p]),
    "dana_conv_gemm_workspace_bytes": (c_int64, []),
    "dana_conv_gemm": (c_int, [POINTER(ConvGemmArgs), c_void_p]),
    "dana_rpn_losses_workspace_bytes": (c_int64, []),
    "dana_rpn_losses": (c_int, [c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int64, c_void_p]),
    "dana_rcnn_losses": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dana_group_mean": (c_int, [c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p]),
    "dana_depthwise_xcorr": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p]),
    "dana_cisa_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dana_cisa_fwd": (c_int, [POINTER(CisaArgs), c_void_p]),
    "dana_stem_s2d": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dana_maxpool3x3s2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dana_avgpool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dana_support_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_float,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "dana_center_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dana_center_rows_pitched": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dana_attn_softmax": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dana_rpn_fg_prob": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "dana_add_pe_split": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int64, c_void_p, c_void_p, c_void_p]),
    "dana_split_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "dana_merge_pair": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_spatial_mean": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "dana_softmax2": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "dana_nhwc_pair_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dana_transpose_segments": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int64, c_void_p, c_void_p,
                                        c_void_p]),
    "dana_grad_prepare": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_void_p]),
    "dana_im2col_t": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, c_int, c_int,
                              c_void_p, c_void_p, c_int64, c_void_p]),
    "dana_pack_conv_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "dana_unpack_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "dana_conv_backward_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "dana_conv_backward": (c_int, [POINTER(ConvBwdArgs), c_void_p]),
    "dana_sgd_momentum": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_void_p]),
}

_lib = None


def load():
    """Load libdana_b200.so (building it first if nvcc is here and the sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("DANA_REBUILD") == "1":
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise DanaError("libdana_b200.so is missing and could not be built -- the CUDA extension is required "
                        "(no CPU fallback). Run `python __graft_entry__.py build`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    if code == DANA_OK:
        return
    lib = load()
    msg = lib.dana_error_string(code).decode()
    extra = ""
    if code == DANA_ECUDA:
        extra = " (cuda error %d)" % lib.dana_last_cuda_error()
    raise DanaError("%s failed: %s%s" % (what, msg, extra))
