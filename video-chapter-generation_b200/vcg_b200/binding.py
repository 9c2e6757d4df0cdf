"""ctypes binding of libvcg_b200.so (C ABI: include/vcg.h).

The library is built in-tree (``csrc/Makefile`` -> ``lib/libvcg_b200.so``).  Loading it needs no GPU; every compute
entry point needs a Blackwell GPU and fails loudly otherwise — there is no CPU or PyTorch fallback.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libvcg_b200.so")
CSRC_DIR = os.path.join(PKG_ROOT, "csrc")

HEAD_MLP, HEAD_ATTN = 0, 1
PREC_BF16, PREC_FP32 = 0, 1
VISION_R50TSM, VISION_NONE = 0, 1
MODALITY_TWO_STREAM, MODALITY_VISION, MODALITY_TEXT, MODALITY_EMBED = 0, 1, 2, 3
DTYPE_F32, DTYPE_I64 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_TANH = 0, 1, 2, 3


class VcgConfig(ctypes.Structure):
    _fields_ = [
        ("clip_frames", ctypes.c_int32),
        ("max_tokens", ctypes.c_int32),
        ("hidden_size", ctypes.c_int32),
        ("head_type", ctypes.c_int32),
        ("precision", ctypes.c_int32),
        ("vision", ctypes.c_int32),
        ("max_batch", ctypes.c_int32),
        ("shift_div", ctypes.c_int32),
        ("modality", ctypes.c_int32),
    ]


class VcgMlpOp(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("in_dim", ctypes.c_int32), ("out_dim", ctypes.c_int32),
                ("eps", ctypes.c_float), ("w", ctypes.c_void_p), ("b", ctypes.c_void_p)]


_fp = ctypes.c_void_p


class VcgCrossAttnParams(ctypes.Structure):
    _fields_ = [("num_heads", ctypes.c_int32)] + [(n, _fp) for n in (
        "lang_norm_w", "lang_norm_b", "vision_norm_w", "vision_norm_b", "pos_w", "pos_b",
        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class VcgSelfAttnParams(ctypes.Structure):
    _fields_ = [("num_heads", ctypes.c_int32)] + [(n, _fp) for n in ("q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class VcgCenterAttnParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("num_heads", "bias_head_stride", "bias_offset", "add_residual")] + [
        (n, _fp) for n in ("pre_norm_w", "pre_norm_b", "post_norm_w", "post_norm_b", "pos_w", "pos_b", "pos_norm_w",
                           "pos_norm_b", "pos_bias", "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b")]


class VcgWindowLayer(ctypes.Structure):
    _fields_ = [(n, _fp) for n in (
        "attn_norm_w", "attn_norm_b", "ffn_norm_w", "ffn_norm_b", "pos_w", "pos_b", "pos_bias",
        "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b",
        "f0_w", "f0_b", "f1_w", "f1_b", "f2_w", "f2_b", "f3_w", "f3_b")]


class VcgWindowStackParams(ctypes.Structure):
    _fields_ = [("num_layers", ctypes.c_int32), ("pos_bias_stride", ctypes.c_int32), ("layers", VcgWindowLayer * 8),
                ("final_norm_w", _fp), ("final_norm_b", _fp), ("cls_w", _fp * 5), ("cls_b", _fp * 5),
                ("cls_norm_w", _fp * 4), ("cls_norm_b", _fp * 4)]


MLP_LINEAR, MLP_LAYERNORM, MLP_RELU, MLP_GELU, MLP_MULHALVES, MLP_MEANGROUPS, MLP_SOFTMAX, MLP_SAVE, MLP_ADDSAVED = range(9)


class VcgProfileEntry(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 64), ("launches", ctypes.c_int64), ("ms", ctypes.c_double),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


_vp, _i32, _i64, _f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float

# name -> (restype, argtypes); must list every symbol declared in include/vcg.h (checked by tests/test_abi.py)
PROTOTYPES = {
    "vcg_create": (ctypes.c_int, [ctypes.POINTER(VcgConfig), ctypes.POINTER(_vp)]),
    "vcg_destroy": (None, [_vp]),
    "vcg_last_error": (ctypes.c_char_p, []),
    "vcg_version": (ctypes.c_char_p, []),
    "vcg_load_tensor": (ctypes.c_int, [_vp, ctypes.c_char_p, _vp, ctypes.POINTER(_i64), _i32, _i32, _vp]),
    "vcg_finalize": (ctypes.c_int, [_vp, _vp]),
    "vcg_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "vcg_forward_vision": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "vcg_forward_text": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "vcg_score_clips_u8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_set_frame_size": (ctypes.c_int, [_vp, _i32, _i32, _vp]),
    "vcg_score_clips_u8_planned": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_score_video_u8": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_score_clips_u8_host": (ctypes.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_forward_host": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_profile_begin": (ctypes.c_int, [_vp]),
    "vcg_profile_end": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(VcgProfileEntry), _i32, ctypes.POINTER(_i32)]),
    "vcg_launch_count": (_i64, [_vp]),
    "vcg_debug_pdl_window": (ctypes.c_int, [_i32, _i32]),
    "vcg_debug_checksums": (ctypes.c_int, [_vp, _vp, _i32, ctypes.POINTER(_i32), _vp]),
    "vcg_op_preprocess_u8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _vp]),
    "vcg_op_resize_u8": (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp]),
    "vcg_op_nchw_to_stem": (ctypes.c_int, [_vp, _i32, _vp, _i32, _vp]),
    "vcg_op_gemm": (ctypes.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "vcg_op_conv2d_nhwc": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32,
                                          _i32, _vp, _i32, _vp, _i32, _i32, _vp]),
    "vcg_op_bottleneck_tail": (ctypes.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                               _i32, _vp]),
    "vcg_op_stem_conv": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "vcg_op_maxpool_tsm": (ctypes.c_int, [_vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp]),
    "vcg_op_stem_conv_act": (ctypes.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _vp]),
    "vcg_op_bn_partials": (ctypes.c_int32, [ctypes.c_int64, _i32]),
    "vcg_op_bn_batch_stats": (ctypes.c_int, [_vp, ctypes.c_int64, _i32, ctypes.c_float, _vp, _vp, _vp, _i32, _vp]),
    "vcg_op_bn_apply": (ctypes.c_int, [_vp, ctypes.c_int64, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp]),
    "vcg_op_tsm_shift": (ctypes.c_int, [_vp, ctypes.c_int64, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "vcg_op_avgpool": (ctypes.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _vp]),
    "vcg_op_bert_attention": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "vcg_op_bert_attention_packed": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, ctypes.c_int64, _vp]),
    "vcg_op_cut_points": (ctypes.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "vcg_op_pr_hits": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "vcg_op_auc_ap": (ctypes.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "vcg_embed_u8": (ctypes.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_embed": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_op_mlp_chain": (ctypes.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i32, ctypes.POINTER(VcgMlpOp), _i32, _vp, _i64, _vp]),
    "vcg_op_cross_attention": (ctypes.c_int, [ctypes.POINTER(VcgCrossAttnParams), _vp, _vp, _i32, _i32, _vp, _vp]),
    "vcg_op_self_attention_first": (ctypes.c_int, [ctypes.POINTER(VcgSelfAttnParams), _vp, _vp, _i32, _i32, _vp, _vp]),
    "vcg_op_bilinear_contract": (ctypes.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "vcg_op_center_attention": (ctypes.c_int, [ctypes.POINTER(VcgCenterAttnParams), _vp, _i32, _i32, _vp, _vp]),
    "vcg_op_window_stack": (ctypes.c_int, [ctypes.POINTER(VcgWindowStackParams), _vp, _i32, _i32, _vp, _vp, _vp]),
    "vcg_op_layernorm": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _vp]),
}

_lib = None


def build_library(force=False):
    """Compile libvcg_b200.so for sm_100a with nvcc (no GPU needed)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", CSRC_DIR, "-j8"], check=True)
    return LIB_PATH


def load_library():
    """dlopen the library and attach prototypes.  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC_DIR}` (or __graft_entry__.build()). "
            "There is no fallback implementation.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status):
    """Map a non-zero C status to RuntimeError carrying vcg_last_error()."""
    if status != 0:
        msg = load_library().vcg_last_error()
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else "vcg: unknown error")
