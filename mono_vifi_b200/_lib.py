"""ctypes binding of libmonovifi_b200.so (include/monovifi_b200.h).  Fails loudly when the library is missing."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MVF_LIB selects another build of the same library (A/B timing of kernel variants); default: the in-tree build
SO_PATH = os.environ.get("MVF_LIB") or os.path.join(_HERE, "libmonovifi_b200.so")

NO_SSIM, AVG_REPROJECTION, DISABLE_AUTOMASKING = 1, 2, 4


class F1Params(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("min_disp", ctypes.c_float),
                ("disp_range", ctypes.c_float), ("smooth_w", ctypes.c_float), ("flags", ctypes.c_int)]


class Conv2dDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int), ("Cin", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int),
                ("Cout", ctypes.c_int), ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("pad", ctypes.c_int),
                ("stride", ctypes.c_int), ("x_stride", ctypes.c_longlong * 3), ("y_stride", ctypes.c_longlong * 3), ("stride_x", ctypes.c_int)]


_vp, _sz, _i, _f = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
_P = ctypes.POINTER(F1Params)
_CD = ctypes.POINTER(Conv2dDesc)

# name -> (restype, argtypes); every symbol include/monovifi_b200.h declares
SIGNATURES = {
    "mvf_version": (_i, []),
    "mvf_last_error": (ctypes.c_char_p, []),
    "mvf_f1_workspace_bytes": (_sz, [_i]),
    "mvf_workspace_init": (_i, [_vp, _sz, _vp]),
    "mvf_f1_forward": (_i, [_P] + [_vp] * 16 + [_vp, _sz, _vp]),
    "mvf_f1_backward": (_i, [_P] + [_vp] * 14 + [_vp, _sz, _vp]),
    "mvf_f1_forward_host": (_i, [_P] + [_vp] * 11),
    "mvf_disp_to_depth_fwd": (_i, [_vp, _vp, _vp, _sz, _f, _f, _vp]),
    "mvf_disp_to_depth_bwd": (_i, [_vp, _vp, _vp, _vp, _sz, _f, _f, _vp]),
    "mvf_backproject_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_backproject_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_project_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp]),
    "mvf_project_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _f, _vp]),
    "mvf_ssim_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_ssim_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_smooth_loss_fwd": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _vp]),
    "mvf_smooth_loss_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_si_log_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _sz, _f, _vp]),
    "mvf_si_log_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _sz, _f, _vp]),
    "mvf_conv2d_packed_filter_floats": (_sz, [_i, _i, _i, _i]),
    "mvf_conv2d_pack_filters": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mvf_conv2d_supported": (_i, [_CD]),
    "mvf_conv2d_forward": (_i, [_CD, _vp, _vp, _vp, _vp, _i, _vp]),
    "mvf_conv2d_pack_chunk": (_i, []),
    "mvf_conv2d_pack_filters_multi": (_i, [_vp, _i, ctypes.c_longlong, _vp]),
    "mvf_conv2d_forward_prelu": (_i, [_CD, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mvf_conv_transpose2d_s2_fwd": (_i, [_CD, _vp, _vp, _vp, _vp, _vp]),
    "mvf_conv2d_wgrad_supported": (_i, [_CD]),
    "mvf_conv2d_wgrad_workspace_floats": (_sz, [_CD]),
    "mvf_conv2d_wgrad": (_i, [_CD, _vp, _vp, _vp, _vp, _sz, _vp]),
    "mvf_upcat_pad_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mvf_upcat_pad_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mvf_maxpool3s2_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_maxpool3s2_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_bn_workspace_floats": (_sz, [ctypes.c_longlong, _i]),
    "mvf_bn_relu_fwd": (_i, [_vp] * 11 + [_sz, ctypes.c_longlong, _i, _f, _f, _i, _vp]),
    "mvf_bn_relu_bwd": (_i, [_vp] * 11 + [_sz, ctypes.c_longlong, _i, _i, _vp]),
    "mvf_act_bwd_bias": (_i, [_vp] * 5 + [_sz, ctypes.c_longlong, _i, _i, _vp]),
    "mvf_bn_sync_stats_fwd": (_i, [_vp, _vp, _vp, _sz, ctypes.c_longlong, _i, _vp]),
    "mvf_bn_sync_apply_fwd": (_i, [_vp] * 11 + [ctypes.c_longlong, _i, _f, _f, _i, _vp]),
    "mvf_bn_sync_stats_bwd": (_i, [_vp] * 9 + [_sz, ctypes.c_longlong, _i, _i, _vp]),
    "mvf_bn_sync_apply_bwd": (_i, [_vp] * 10 + [_sz, ctypes.c_longlong, _i, _i, _vp]),
    "mvf_flow_warp_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "mvf_flow_warp_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "mvf_flow_warp_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "mvf_resize_bilinear_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _f, _i, _vp]),
    "mvf_resize_bilinear_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _f, _f, _i, _f, _f, _i, _vp]),
    "mvf_prelu_cl_fwd": (_i, [_vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _vp]),
    "mvf_pose_matrix_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "mvf_pose_matrix_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "mvf_dwconv3x3_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "mvf_dwconv3x3_wgrad_workspace_floats": (_sz, [ctypes.c_longlong, _i]),
    "mvf_dwconv3x3_wgrad": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "mvf_gelu_fwd": (_i, [_vp, _vp, ctypes.c_longlong, _vp]),
    "mvf_gelu_bwd": (_i, [_vp, _vp, _vp, ctypes.c_longlong, _vp]),
    "mvf_layernorm_cl_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_longlong, _i, _f, _vp]),
    "mvf_layernorm_bwd_workspace_floats": (_sz, [ctypes.c_longlong, _i]),
    "mvf_layernorm_cl_bwd": (_i, [_vp] * 9 + [_sz, ctypes.c_longlong, _i, _vp]),
    "mvf_peer_buffer_bytes": (_sz, []),
    "mvf_peer_allreduce_f64": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "mvf_bn_eval_fwd": (_i, [_vp] * 7 + [ctypes.c_longlong, _i, _f, _i, _vp]),
    "mvf_depth_eval_workspace_bytes": (_sz, [_i, _i]),
    "mvf_depth_eval": (_i, [_vp, _i, _i, _vp, _i, _i, _f, _f, _i, _f, _vp, _sz, _vp, _vp]),
    "mvf_input_pipeline_workspace_floats": (_sz, [_i, _i]),
    "mvf_input_pipeline": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_dispconv_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_dispconv_dgrad": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_dispconv_wgrad_workspace_floats": (_sz, [ctypes.c_longlong, _i]),
    "mvf_dispconv_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp]),
    "mvf_stream_capture_id": (ctypes.c_ulonglong, [_vp]),
    "mvf_conv2d_dgrad_s2_supported": (_i, [_CD]),
    "mvf_conv2d_dgrad_s2": (_i, [_CD, _vp, _vp, _vp, _vp]),
    "mvf_conv2d_dgrad_s2_plan": (_i, [_i, _i, _i, ctypes.POINTER(ctypes.c_int), _i]),
    "mvf_adamw_workspace_bytes": (_sz, []),
    "mvf_adamw_step": (_i, [_vp, _vp, _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _vp]),
    "mvf_gather_grads": (_i, [_vp, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_longlong),
                              ctypes.POINTER(ctypes.c_longlong), ctypes.c_int, _vp]),
    "mvf_selftest_umma": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "mvf_selftest_umma_rows": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "mvf_conv2d_debug_buffer": (None, [_vp]),
    "mvf_selftest_division": (_i, [_vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise LibraryMissing(
                "%s not found: build it with `python -m mono_vifi_b200.build` (nvcc, sm_100a). "
                "There is no CPU / PyTorch fallback for these ops." % SO_PATH)
        l = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libmonovifi_b200 %s failed (%d): %s" % (what, rc, lib().mvf_last_error().decode()))


def f1_params(B, H, W, min_depth=0.1, max_depth=100.0, smooth_w=1e-3, flags=0):
    """python-double arithmetic then one cast to fp32, exactly what layers.py:21-23 hands to torch."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    return F1Params(B, H, W, min_disp, max_disp - min_disp, smooth_w, flags)
