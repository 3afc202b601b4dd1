"""ctypes binding of libssr_b200.so (include/ssr_b200.h).

There is deliberately no fallback: if the shared library is missing or the device is not an
sm_100 part, importing / calling raises.  Build the library with `python __graft_entry__.py`
(or `make -C studiosr_b200/csrc`)."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssr_b200.so")

SSR_ARCH_SWINIR, SSR_ARCH_EDSR, SSR_ARCH_RCAN, SSR_ARCH_HAT, SSR_ARCH_HAN, SSR_ARCH_SWINFIR = 0, 1, 2, 3, 4, 5
PREC_FP32, PREC_TF32, PREC_BF16, PREC_TF32X3 = 0, 1, 2, 3
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16, "tf32x3": PREC_TF32X3}
PAD_EVAL, PAD_TRAIN = 0, 1
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU = 0, 1, 2, 3
SSR_MAX_LAYERS = 16


class ModelConfig(ctypes.Structure):
    _fields_ = [
        ("arch", c_int), ("precision", c_int), ("scale", c_int), ("n_colors", c_int), ("img_range", c_float),
        ("embed_dim", c_int), ("n_layers", c_int), ("depths", c_int * SSR_MAX_LAYERS),
        ("num_heads", c_int * SSR_MAX_LAYERS), ("window_size", c_int), ("mlp_ratio", c_float), ("upsampler", c_int),
        ("n_feats", c_int), ("n_resblocks", c_int), ("res_scale", c_float),
        ("n_resgroups", c_int), ("reduction", c_int),
        ("compress_ratio", c_int), ("squeeze_factor", c_int), ("conv_scale", c_float), ("overlap_ratio", c_float),
    ]


# name -> (restype, argtypes); every symbol declared in include/ssr_b200.h
SYMBOLS = {
    "ssr_version": (c_int, []),
    "ssr_last_error": (c_char_p, []),
    "ssr_device_check": (c_int, [c_int]),
    "ssr_model_create": (c_int, [POINTER(ModelConfig), c_int, POINTER(c_void_p)]),
    "ssr_model_set_param": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "ssr_model_finalize": (c_int, [c_void_p]),
    "ssr_model_destroy": (None, [c_void_p]),
    "ssr_model_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "ssr_model_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_model_upscale_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_tiled_num_tiles": (c_int, [c_int, c_int, c_int, c_int]),
    "ssr_model_tiled_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    "ssr_model_upscale_tiled_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                           c_size_t, c_void_p]),
    "ssr_model_upscale_tiled_u8_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_void_p, c_size_t, c_void_p]),
    "ssr_tiled_tile_elems": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "ssr_model_tiles_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int, c_int]),
    "ssr_model_tiles_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "ssr_model_blend_tiles_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ssr_model_train_bind": (c_int, [c_void_p, c_int, POINTER(c_char_p), POINTER(c_int64)]),
    "ssr_model_train_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int, c_int]),
    "ssr_model_train_forward": (c_int, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_void_p, c_size_t, c_void_p]),
    "ssr_model_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_void_p), c_int, c_int, c_int, c_void_p,
                                         c_size_t, c_void_p]),
    "ssr_model_train_input_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ssr_l1_loss_workspace_bytes": (c_size_t, []),
    "ssr_l1_loss": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssr_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_double, c_double,
                              c_int64, c_float, c_void_p]),
    "ssr_psnr_workspace_bytes": (c_size_t, []),
    "ssr_psnr_mse_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssr_augment_pairs_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ssr_tensors_checksum": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "ssr_launch_count": (c_int64, []),
    "ssr_note_graph_replay": (None, [c_int64]),
    "ssr_profile_begin": (c_int, []),
    "ssr_profile_end": (c_int, [c_char_p, c_size_t]),
    "ssr_op_linear": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_op_conv3x3": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_float, c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_op_conv3x3_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_float, c_void_p, c_size_t, c_void_p]),
    "ssr_op_swin_mlp": (c_int, [c_void_p] * 14 + [c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_op_window_attention": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_void_p, c_size_t, c_void_p]),
    "ssr_op_swin_attn": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "ssr_op_workspace_bytes": (c_size_t, [c_int64]),
}

_lib = None


def load():
    """Load libssr_b200.so (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: studiosr_b200 has no CPU / PyTorch fallback. "
            "Build it with `python __graft_entry__.py` or `make -C studiosr_b200/csrc`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SsrError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().ssr_last_error()
        raise SsrError(f"libssr_b200 error {rc}: {msg.decode() if msg else ''}")
