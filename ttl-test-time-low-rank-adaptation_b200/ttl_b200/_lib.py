"""ctypes binding of libttl_b200.so (the C ABI declared in include/ttl_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing the product path
raises.  (`python build.py` in the package directory, or `__graft_entry__.build()`, produces it.)
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libttl_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
vp = C.c_void_p


class TtlConfig(C.Structure):
    _fields_ = [("image_size", C.c_int32), ("patch", C.c_int32), ("width", C.c_int32), ("layers", C.c_int32),
                ("heads", C.c_int32), ("mlp_dim", C.c_int32), ("proj_dim", C.c_int32), ("max_views", C.c_int32),
                ("max_classes", C.c_int32), ("lora_rank", C.c_int32), ("lora_alpha", C.c_float),
                ("lora_layer_lo", C.c_int32), ("lora_layer_hi", C.c_int32), ("ln_eps", C.c_float),
                ("device", C.c_int32), ("max_samples", C.c_int32), ("precision", C.c_int32), ("text_mode", C.c_int32),
                ("context", C.c_int32), ("vocab", C.c_int32)]


class TtlHparams(C.Structure):
    _fields_ = [("head", C.c_int32), ("tta_steps", C.c_int32), ("selection_p", C.c_double), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("deyo_margin_e0", C.c_float)]


class TtlGemmRecord(C.Structure):
    _fields_ = [("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("epi", C.c_int32), ("ms", C.c_float)]


class TtlViewSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("top", C.c_int32), ("left", C.c_int32), ("height", C.c_int32),
                ("width", C.c_int32), ("flip", C.c_int32)]


class TtlTextConfig(C.Structure):
    _fields_ = [("vocab", C.c_int32), ("context", C.c_int32), ("width", C.c_int32), ("layers", C.c_int32),
                ("heads", C.c_int32), ("mlp_dim", C.c_int32), ("proj_dim", C.c_int32), ("max_prompts", C.c_int32),
                ("ln_eps", C.c_float), ("device", C.c_int32)]


class TtlDeyoOptions(C.Structure):
    _fields_ = [("filter_ent", C.c_int32), ("filter_plpd", C.c_int32), ("reweight_ent", C.c_int32), ("reweight_plpd", C.c_int32),
                ("plpd_threshold", C.c_float), ("aug_type", C.c_int32), ("occlusion_size", C.c_int32), ("row_start", C.c_int32),
                ("column_start", C.c_int32), ("patch_len", C.c_int32), ("perm_host", vp), ("perm_numel", C.c_int64),
                ("forced_keep_host", vp)]


class TtlOutputs(C.Structure):
    _fields_ = [("logits0", vp), ("entropy", vp), ("idx", vp), ("loss", vp), ("pred_logits", vp)]


# weight kinds / enums (mirror include/ttl_b200.h)
W_CLASS_EMB, W_PATCH_EMB, W_POS_EMB, W_PRE_LN_G, W_PRE_LN_B, W_POST_LN_G, W_POST_LN_B, W_VIS_PROJ, W_TOKEN_EMB = range(9)
(W_LN1_G, W_LN1_B, W_Q_W, W_Q_B, W_K_W, W_K_B, W_V_W, W_V_B, W_O_W, W_O_B, W_LN2_G, W_LN2_B, W_FC1_W, W_FC1_B,
 W_FC2_W, W_FC2_B) = range(16, 32)
LORA_A_Q, LORA_B_Q, LORA_A_V, LORA_B_V = range(4)
LORA_PARAM, LORA_GRAD, LORA_INIT = range(3)
HEAD_TPT, HEAD_DEYO = 0, 1
PRECISION_BF16, PRECISION_FP32 = 0, 1
VIEW_CLEAN, VIEW_CROP = 0, 1
AUG_OCC, AUG_PATCH, AUG_PIXEL = 0, 1, 2
TW_TOKEN_EMB, TW_POS_EMB, TW_FINAL_LN_G, TW_FINAL_LN_B, TW_TEXT_PROJ = range(5)
EPI_BF16, EPI_GELU, EPI_RESID_F32, EPI_PATCH_F32, EPI_F32, EPI_GELU_BWD = range(6)

_SIGS = {
    "ttl_version": (C.c_int, []),
    "ttl_create": (C.c_int, [C.POINTER(vp), C.POINTER(TtlConfig)]),
    "ttl_destroy": (None, [vp]),
    "ttl_last_error": (C.c_char_p, [vp]),
    "ttl_set_weight": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_int64]),
    "ttl_set_text_features": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_float]),
    "ttl_lora_set_init": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_int64]),
    "ttl_lora_reset": (C.c_int, [vp, vp]),
    "ttl_lora_get": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64]),
    "ttl_lora_get_sample": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64]),
    "ttl_lora_device_ptr": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64)]),
    "ttl_lora_touch": (C.c_int, [vp, vp]),
    "ttl_adamw_step": (C.c_int, [vp, C.POINTER(TtlHparams), vp]),
    "ttl_forward": (C.c_int, [vp, vp, C.c_int32, C.c_int32, vp, vp]),
    "ttl_backward": (C.c_int, [vp, vp, vp]),
    "ttl_adapt_predict": (C.c_int, [vp, vp, C.c_int32, C.POINTER(TtlHparams), vp, C.POINTER(TtlOutputs), vp]),
    "ttl_adapt_predict_host": (C.c_int, [vp, vp, C.c_int32, C.POINTER(TtlHparams), vp, C.POINTER(TtlOutputs), vp]),
    "ttl_adapt_predict_batch": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(TtlHparams), vp, C.POINTER(TtlOutputs), vp]),
    "ttl_image_features": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
    "ttl_text_set_prompts": (C.c_int, [vp, vp, C.c_int32, C.c_float, vp]),
    "ttl_text_features": (C.c_int, [vp, vp, vp]),
    "ttl_text_adapt_predict": (C.c_int, [vp, vp, C.c_int32, C.POINTER(TtlHparams), vp, C.POINTER(TtlOutputs), vp]),
    "ttl_adapt_predict_batch_deyo": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(TtlHparams), C.POINTER(TtlDeyoOptions),
                                               C.POINTER(TtlOutputs), vp]),
    "ttl_deyo_last_plpd": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32]),
    "ttl_adapt_predict_batch_host": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(TtlHparams), vp, C.POINTER(TtlOutputs),
                                               vp]),
    "ttl_adapt_predict_batch_host_async": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.POINTER(TtlHparams), vp,
                                                     C.POINTER(TtlOutputs), vp]),
    "ttl_set_pixel_norm": (C.c_int, [vp, vp, vp]),
    "ttl_make_views": (C.c_int, [vp, vp, vp, vp, C.c_int32, vp, C.c_int32, vp, vp]),
    "ttl_adapt_predict_images_async": (C.c_int, [vp, vp, vp, vp, C.c_int32, vp, C.c_int32, C.POINTER(TtlHparams), vp,
                                                 C.POINTER(TtlOutputs), vp]),
    "ttl_text_create": (C.c_int, [C.POINTER(vp), C.POINTER(TtlTextConfig)]),
    "ttl_text_destroy": (None, [vp]),
    "ttl_text_last_error": (C.c_char_p, [vp]),
    "ttl_text_set_weight": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_int64]),
    "ttl_text_encode": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
    "ttl_set_graphs": (C.c_int, [vp, C.c_int32]),
    "ttl_last_launch_count": (C.c_int64, [vp]),
    "ttl_profile_gemm": (C.c_int, [vp, C.c_int32]),
    "ttl_profile_read": (C.c_int, [vp, C.POINTER(TtlGemmRecord), C.c_int32, C.POINTER(C.c_int32)]),
    "ttl_op_logits_entropy": (C.c_int, [vp, vp, C.c_float, vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]),
    "ttl_op_entropy": (C.c_int, [vp, vp, C.c_int32, C.c_int32, vp]),
    "ttl_op_select": (C.c_int, [vp, C.c_int32, C.c_int32, vp, vp]),
    "ttl_op_tpt_loss": (C.c_int, [vp, vp, C.c_int32, C.c_int32, vp, vp, vp]),
    "ttl_op_deyo_loss": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_float, vp, vp, vp]),
    "ttl_op_gemm": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp,
                              vp, vp, C.c_int32, C.c_int32, vp]),
    "ttl_op_layernorm": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_float, vp]),
    "ttl_op_layernorm_bwd": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_float, vp]),
    "ttl_op_attention_fwd": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp]),
    "ttl_op_attention_bwd": (C.c_int, [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp]),
    "ttl_op_im2col": (C.c_int, [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp]),
    "ttl_op_adamw": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float,
                               C.c_float, vp]),
    "ttl_op_skinny_reduce": (C.c_int, [vp, C.c_int32, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, C.c_float, vp,
                                       C.c_int32, vp, vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)
_lib = None


def load() -> C.CDLL:
    """Load the CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python build.py` (there is no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, ctx=None) -> None:
    if rc != 0:
        msg = load().ttl_last_error(ctx)
        raise RuntimeError(f"libttl_b200 error {rc}: {(msg or b'').decode()}")
