"""Drive the REAL reference modules (lntzm/MESM ``model`` / ``utils`` packages) the way runner.py / eval.py do.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``runner`` and ``eval`` themselves are not importable offline (they pull
ftfy / h5py / nltk), so the two pieces of glue are restated here: ``build_reference`` = runner.build_model
(runner.py:255-298) and ``reference_decode`` = the loop of eval.py:64-99, 111-116, 476-485, both
calling nothing but the reference's own classes and functions.  Used by oracle/gen_golden.py (from /root/reference) and
by bench.py's reference legs (from oracle/_ref, staged by oracle/make_ref.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_root():
    """Directory holding the reference's ``model`` and ``utils`` packages, or None."""
    for root in (os.environ.get("MESM_REFERENCE"), os.path.join(HERE, "_ref"), "/root/reference"):
        if root and os.path.isfile(os.path.join(root, "model", "model.py")) and os.path.isfile(os.path.join(root, "utils", "temporal_nms.py")):
            return root
    return None


def import_reference(root=None):
    """Put the reference on sys.path and import its ``model`` / ``utils`` packages (returns them)."""
    root = root or reference_root()
    if root is None:
        raise ImportError("reference modules not found (oracle/_ref is staged by `python -m oracle.make_ref`)")
    sys.dont_write_bytecode = True
    if root not in sys.path:
        sys.path.insert(0, root)
    import model as ref_model
    import utils as ref_utils
    if not os.path.abspath(ref_model.__file__).startswith(os.path.abspath(root)):
        raise ImportError(f"`model` resolved to {ref_model.__file__}, not to the reference under {root}")
    return ref_model, ref_utils


def build_reference(cfg):
    """runner.build_model (runner.py:255-298) for an OracleConfig / any object with the same attributes; text_encoder=None
    (word features are the input, runner.py:261-263)."""
    from model.model import MESM
    from model.transformer import T2VEncoder, T2VEncoder_TwoMLP, Transformer
    from model.position_encoding import PositionEmbeddingSine, TrainablePositionalEncoding
    kw = dict(d_model=cfg.hidden_dim, dropout=0.1, nhead=cfg.nheads, dim_feedforward=cfg.dim_feedforward,
              normalize_before=False, activation="prelu")
    enh = (T2VEncoder if cfg.share_mlp else T2VEncoder_TwoMLP)(num_encoder_layers=cfg.num_recfw_layers, **kw)
    t2v = T2VEncoder(num_encoder_layers=cfg.t2v_layers, **kw)
    tr = Transformer(num_encoder_layers=cfg.enc_layers, num_decoder_layers=cfg.dec_layers,
                     return_intermediate_dec=True, **kw)
    vpos = PositionEmbeddingSine(cfg.hidden_dim, normalize=True)
    tpos = TrainablePositionalEncoding(cfg.max_words_l + 1 if cfg.rec_ss else cfg.max_words_l, cfg.hidden_dim, 0.5)
    m = MESM(text_encoder=None, enhance_encoder=enh, t2v_encoder=t2v, transformer=tr, vid_position_embed=vpos,
             txt_position_embed=tpos, txt_dim=cfg.t_feat_dim, vid_dim=cfg.v_feat_dim, num_queries=cfg.num_queries,
             input_dropout=0.5, aux_loss=cfg.aux_loss, max_video_l=cfg.max_video_l, max_words_l=cfg.max_words_l,
             normalize_txt=True, use_txt_pos=False, span_loss_type="l1", n_input_proj=cfg.n_input_proj,
             rec_fw=cfg.rec_fw, vocab_size=cfg.vocab_size, rec_ss=cfg.rec_ss, num_recss_layers=cfg.num_recss_layers)
    return m.eval()


def reference_decode(logits, spans, duration, cfg, nms_thd, max_before_nms=10, max_after_nms=10):
    """eval.py:64-99, 111-116, 476-485 with the reference's own utils (tensors on any device, as eval.py has them)."""
    import torch.nn.functional as F
    from utils import span_cxw_to_xx, PostProcessorDETR, temporal_nms
    prob = F.softmax(logits, -1)
    scores = prob[..., 0]
    res = []
    for idx, (sp, sc) in enumerate(zip(spans, scores)):
        sp = span_cxw_to_xx(sp) * duration[idx]
        rows = torch.cat([sp, sc[:, None]], dim=1).cpu().tolist()
        order = sorted(range(len(rows)), key=lambda i: rows[i][2], reverse=True)
        rows = sorted(rows, key=lambda x: x[2], reverse=True)
        rows = [[float(f"{e:.4f}") for e in row] for row in rows]
        res.append(dict(pred_relevant_windows=rows, order=order))
    pp = PostProcessorDETR(clip_length=cfg.clip_len, min_ts_val=0, max_ts_val=cfg.max_ts_val, min_w_l=2, max_w_l=150,
                           move_window_method="left",
                           process_func_names=("clip_ts", "round_multiple") if cfg.clip_len != -1 else ("clip_ts",))
    res = pp(res)
    out = []
    for r in res:
        w = r["pred_relevant_windows"]
        kept = temporal_nms(w[:max_before_nms], nms_thd=nms_thd, max_after_nms=max_after_nms) if nms_thd != -1 else w
        out.append(dict(windows=w, order=r["order"], nms_windows=kept))
    return out


def reference_forward(ref, cfg, batch, neg_index=None):
    """model(**batch, dataset_name=..., is_training=False) as eval.py:63 calls it (text_encoder=None: ``words_id`` carries the
    word features).  ``neg_index`` pins the RNG draw of model/model.py:260 (sample_outclass_neg) when given."""
    import model.model as mm
    saved = mm.sample_outclass_neg
    if neg_index is not None:
        mm.sample_outclass_neg = lambda nc, _neg=neg_index: _neg
    try:
        with torch.inference_mode():
            return ref(video_feat=batch["video_feat"], video_mask=batch["video_mask"], words_id=batch["words_feat"].clone(),
                       words_mask=None, words_weight=None, num_clips=batch["num_clips"], dataset_name=cfg.dataset_name,
                       is_training=False)
    finally:
        mm.sample_outclass_neg = saved
