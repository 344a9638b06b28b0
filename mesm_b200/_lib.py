"""ctypes binding of libmesm_b200.so (C ABI declared in include/mesm_b200.h).

The library is the product: there is no Python/CPU fallback.  Importing this module without the built shared object,
or calling into it without a B200-class device, raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MESM_B200_LIB") or os.path.join(_HERE, "libmesm_b200.so")     # override: A/B builds of the library


class MesmCfg(Structure):
    _fields_ = [(n, c_int32) for n in (
        "v_feat_dim", "t_feat_dim", "hidden_dim", "nheads", "dim_feedforward", "num_queries", "num_recfw_layers",
        "t2v_layers", "enc_layers", "dec_layers", "num_recss_layers", "n_input_proj", "rec_fw", "rec_ss", "share_mlp",
        "qvh_grouping", "max_words_l", "max_video_l")]


class MesmInputs(Structure):
    _fields_ = [("B", c_int32), ("Lv", c_int32), ("Lt", c_int32), ("G", c_int32),
                ("video_feat", c_void_p), ("video_mask", c_void_p), ("words_feat", c_void_p),
                ("num_clips", POINTER(c_int64)), ("neg_index", c_void_p), ("video_len", POINTER(c_int32)),
                ("shared_group_video", c_int32), ("video_feat_f16", c_int32)]


class MesmOutputs(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "pred_logits", "pred_spans", "saliency_scores", "neg_saliency_scores", "aux_logits", "aux_spans",
        "projed_video_feat", "enhanced_video_feat", "recon_feat", "projed_recon_feat", "expanded_words_feat",
        "expanded_words_mask", "memory", "memory_global", "hs")]


class MesmDecodeParams(Structure):
    _fields_ = [("clip_len", c_double), ("min_ts_val", c_double), ("max_ts_val", c_double), ("nms_thd", c_double),
                ("max_before_nms", c_int32), ("max_after_nms", c_int32), ("sort_results", c_int32)]


# every symbol include/mesm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mesm_abi_version": (c_int, []),
    "mesm_create": (c_void_p, [POINTER(MesmCfg), c_int]),
    "mesm_destroy": (None, [c_void_p]),
    "mesm_last_error": (c_char_p, [c_void_p]),
    "mesm_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_int, c_void_p]),
    "mesm_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "mesm_set_chunk_pairs": (c_int, [c_void_p, c_int32]),
    "mesm_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32, c_int32, c_int32]),
    "mesm_forward": (c_int, [c_void_p, POINTER(MesmInputs), POINTER(MesmOutputs), c_void_p, c_size_t, c_void_p]),
    "mesm_last_launch_count": (c_int64, [c_void_p]),
    "mesm_last_feature_bytes": (c_int64, [c_void_p]),
    "mesm_profile_begin": (None, []),
    "mesm_profile_end": (None, [POINTER(c_double)]),
    "mesm_profile_report": (c_char_p, []),
    "mesm_upload_clips": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, POINTER(c_int64), c_int32,
                                  POINTER(c_int64), c_void_p]),
    "mesm_upload_clips_f16": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, POINTER(c_int64), c_int32,
                                      POINTER(c_int64), c_void_p]),
    "mesm_memcpy_batch_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mesm_clip_create": (c_void_p, [c_int32] * 7),
    "mesm_clip_destroy": (None, [c_void_p]),
    "mesm_clip_last_error": (c_char_p, [c_void_p]),
    "mesm_clip_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int, c_int, c_void_p]),
    "mesm_clip_finalize": (c_int, [c_void_p, c_void_p]),
    "mesm_clip_workspace_bytes": (c_size_t, [c_void_p, c_int32]),
    "mesm_clip_forward": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mesm_video_feat_rows": (c_int32, [POINTER(c_int32), c_int32, c_int32]),
    "mesm_video_feat_workspace_bytes": (c_size_t, [POINTER(c_int32), c_int32, c_int32]),
    "mesm_build_video_feat": (c_int, [POINTER(c_void_p), POINTER(c_int32), POINTER(c_int32), c_int32, c_int32, c_int32, c_int32, c_int32,
                                      c_void_p, c_int32, c_void_p, c_size_t, c_void_p]),
    "mesm_decode_nms": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, POINTER(MesmDecodeParams), c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "mesm_temporal_nms": (c_int, [c_void_p, c_void_p, c_int32, c_double, c_int32, c_void_p, c_void_p, c_void_p]),
    "mesm_post_process": (c_int, [c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_void_p]),
    "mesm_temporal_iou": (c_int, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mesm_span_convert": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "mesm_align_scores": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_float, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "mesm_align_workspace_bytes": (c_size_t, [c_int32]),
    "mesm_mr_metrics": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "mesm_saliency_loss_workspace_bytes": (c_size_t, [c_int32]),
    "mesm_saliency_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p, c_int32,
                                   c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mesm_mha_noproj": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mesm_mha_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32, c_int32]),
    "mesm_t2v_encoder": (c_int, [c_void_p, c_char_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int32, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mesm_t2v_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int32]),
    "mesm_transformer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mesm_debug_watchdog": (c_int, [c_void_p]),
    "mesm_debug_attn_trace": (c_int, [c_void_p]),
    "mesm_debug_attention": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "mesm_debug_linear": (c_int, [c_void_p] * 9 + [c_int32] * 5 + [c_float, c_void_p, c_void_p, c_int32, c_void_p]),
    "mesm_transformer_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32]),
}

_lib = None


def lib():
    """The loaded shared library (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C mesm_b200/csrc` (or `python -c 'import "
                "__graft_entry__ as g; g.build()'`).  mesm_b200 has no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)      # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.mesm_abi_version() != 3:
            raise ImportError("libmesm_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(code, ctx=None):
    if code != 0:
        msg = lib().mesm_last_error(ctx)
        raise RuntimeError(f"mesm_b200 error {code}: {msg.decode() if msg else '?'}")
