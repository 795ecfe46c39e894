"""ctypes binding of ``libcodd_b200.so`` (the C ABI declared in ``include/codd_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C codd_b200/csrc``.
There is no fallback: if the shared object is missing, importing the ops raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcodd_b200.so")

# error codes / activation codes (mirror include/codd_b200.h)
E_BADARG, E_SHAPE, E_UNSUPPORTED, E_ALIGN = -1, -2, -3, -4
ACT_NONE, ACT_LEAKY, ACT_RELU, ACT_RELU_CH0, ACT_SIGMOID, ACT_MISH, ACT_TANH = 0, 1, 2, 3, 4, 5, 6


class ConvDesc(Structure):
    _fields_ = [(k, c_int) for k in (
        "n", "h", "w", "c0", "ld0", "c1", "ld1", "cout", "ldo", "kh", "kw", "sh", "sw",
        "ph", "pw", "dil", "ho", "wo", "act", "ldr", "res_bcast", "res_after_act")]


_FP = c_void_p  # device pointers travel as integers

# name -> (restype, argtypes); every symbol include/codd_b200.h declares
SIGNATURES = {
    "codd_version": (c_int, []),
    "codd_error_string": (c_char_p, [c_int]),
    "codd_conv2d_nhwc": (c_int, [POINTER(ConvDesc), _FP, _FP, _FP, _FP, _FP, _FP, c_void_p]),
    "codd_conv3x3_tc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, c_int, c_int, c_int, c_int, _FP,
                                c_int, c_int, c_void_p]),
    "codd_conv3x3_tc_dil": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, c_int, c_int, c_int, c_int,
                                    _FP, c_int, c_int, c_int, c_void_p]),
    "codd_conv3x3_tc_ring": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, c_int, c_int, c_int, c_int,
                                     _FP, c_int, c_void_p]),
    "codd_conv3x3_tc_ring_dil": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, c_int, c_int, c_int, c_int,
                                         _FP, c_int, c_int, c_void_p]),
    "codd_conv3x3x2_tc_ring": (c_int, [_FP, c_int, c_int, c_int, c_int, _FP, _FP, c_int, _FP, _FP, _FP, c_int, c_int, _FP,
                                       c_int, c_void_p]),
    "codd_conv4x4s2_tc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, c_int, c_int, _FP, c_int, c_void_p]),
    "codd_upmerge_nhwc": (c_int, [_FP, c_int, c_int, _FP, c_int, c_int, _FP, _FP, c_int, _FP, _FP, c_int, c_int, c_int, c_int,
                                  _FP, c_int, c_void_p]),
    "codd_tile_features_tc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, _FP, c_int, _FP, c_void_p]),
    "codd_conv3x3_image": (c_int, [_FP, _FP, c_int, c_int, c_int, _FP, _FP, c_int, _FP, c_int, c_void_p]),
    "codd_deconv2x2_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, c_int, _FP, c_int, c_int,
                                    c_void_p]),
    "codd_tile_features": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, _FP, _FP, _FP, c_int, _FP, c_void_p]),
    "codd_cost_volume": (c_int, [_FP, _FP, c_int, c_int, c_int, c_int, _FP, _FP, _FP, c_void_p]),
    "codd_cost_volume_pyramid": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), c_int, POINTER(c_int), POINTER(c_int),
                                         POINTER(c_int), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), c_void_p]),
    "codd_tile_hyp_init": (c_int, [_FP, _FP, _FP, c_int, c_int, _FP, _FP, c_int, c_int, c_int, _FP, c_int,
                                   c_void_p]),
    "codd_plane_upsample": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, c_float, _FP, c_int, c_void_p]),
    "codd_tile_warp_cost": (c_int, [_FP, c_int, _FP, c_int, _FP, c_int, _FP, c_int, _FP, _FP, c_int, c_int,
                                    c_int, _FP, c_int, _FP, c_void_p]),
    "codd_tile_warp_cost_nhwc": (c_int, [_FP, c_int, _FP, c_int, c_int, _FP, c_int, _FP, c_int, _FP, _FP, c_int, c_int,
                                         c_int, _FP, c_int, _FP, c_void_p]),
    "codd_hyp_select": (c_int, [_FP, c_int, _FP, c_int, c_int, c_int, c_int, _FP, c_int, c_void_p]),
    "codd_fusion_cues_lowres": (c_int, [_FP, c_int, _FP, c_int, _FP, c_int, _FP, c_int, _FP, _FP, c_int, c_int, c_int,
                                        c_int, _FP, c_int, _FP, c_int, _FP, c_int, c_void_p]),
    "codd_fusion_forget_in": (c_int, [_FP, _FP, _FP, _FP, _FP, _FP, c_int, c_int, c_int, _FP, c_int, _FP, c_void_p]),
    "codd_fusion_blend": (c_int, [_FP, _FP, _FP, c_int, _FP, _FP, _FP, c_int, c_int, c_int, c_int, _FP, _FP, _FP,
                                  c_void_p]),
    "codd_raft_motion_info": (c_int, [_FP, _FP, _FP, _FP, c_int, c_int, c_int, _FP, _FP, c_int, c_void_p]),
    "codd_avgpool2_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, c_int, c_void_p]),
    "codd_corr_lookup": (c_int, [_FP, c_int, POINTER(c_void_p), POINTER(c_int), c_int, _FP, c_int, c_int, c_int, c_int,
                                 c_int, c_int, _FP, c_int, c_void_p]),
    "codd_se3_gn_step": (c_int, [_FP, _FP, c_int, _FP, c_int, _FP, c_int, _FP, _FP, c_int, c_int, c_int, c_int, c_float,
                                 c_float, _FP, c_void_p]),
    "codd_cvx_upsample": (c_int, [_FP, c_int, c_int, _FP, c_int, c_int, c_int, c_int, _FP, c_int, c_void_p]),
    "codd_se3_upsample_flow": (c_int, [_FP, _FP, c_int, _FP, _FP, c_int, c_int, c_int, _FP, _FP, _FP, c_void_p]),
    "codd_splat_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "codd_splat_warp": (c_int, [_FP, _FP, _FP, _FP, c_int, c_int, c_int, c_int, c_int, c_float, c_float, _FP, c_int, _FP,
                                _FP, _FP, c_size_t, c_void_p]),
    "codd_im2col_split": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _FP, _FP, c_int,
                                  c_void_p]),
    "codd_gemm_tc": (c_int, [_FP, _FP, c_int, _FP, _FP, c_int, c_int, c_int, c_int, _FP, _FP, c_int, c_int, _FP, c_int,
                             c_void_p]),
    "codd_instance_norm_workspace_bytes": (c_size_t, [c_int, c_int]),
    "codd_instance_norm_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, c_float, c_int, _FP, c_int, _FP, c_int,
                                        _FP, c_size_t, c_void_p]),
    "codd_resize_bilinear_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, c_int, _FP, c_int, c_int, c_int,
                                          c_int, c_int, c_void_p]),
    "codd_eltwise_nhwc": (c_int, [c_int, c_int, _FP, c_int, _FP, c_int, _FP, c_int, _FP, c_int, c_size_t, c_int,
                                  c_void_p]),
    "codd_disp_to_depth": (c_int, [_FP, c_size_t, c_float, _FP, c_void_p]),
    "codd_subsample_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _FP, c_int, c_void_p]),
    "codd_stage_images_u8": (c_int, [_FP, c_int, c_int, c_int, POINTER(c_float), POINTER(c_float), c_int, c_int, c_int, _FP,
                                     c_void_p]),
    "codd_disp_metrics": (c_int, [_FP, ctypes.c_longlong, c_int, _FP, _FP, c_int, c_int, c_int, c_float, c_float, _FP, _FP,
                                  c_void_p]),
    "codd_temporal_metrics": (c_int, [_FP, _FP, _FP, ctypes.c_longlong, c_int, _FP, _FP, _FP, ctypes.c_longlong, c_int, _FP,
                                      _FP, _FP, c_int, c_int, c_int, c_float, c_float, _FP, c_void_p]),
    "codd_gt_disp_change": (c_int, [_FP, _FP, _FP, _FP, c_int, c_int, c_int, _FP, _FP, c_void_p]),
    "codd_sceneflow_metrics": (c_int, [_FP, ctypes.c_longlong, ctypes.c_longlong, _FP, ctypes.c_longlong, c_int, _FP, _FP, _FP,
                                       _FP, _FP, _FP, c_int, c_int, c_int, c_float, c_float, _FP, c_void_p]),
    "codd_nhwc_to_nchw": (c_int, [_FP, c_int, c_int, c_int, c_int, c_int, _FP, c_void_p]),
    "codd_nchw_to_nhwc": (c_int, [_FP, c_int, c_int, c_int, c_int, _FP, c_int, c_void_p]),
}

_lib = None


class CoddError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CoddError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C codd_b200/csrc`.  codd_b200 has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().codd_error_string(code).decode()
        raise CoddError(f"{what} failed: {msg} (code {code})")
