"""GPU: the drop-in Python classes (reference names / signatures / state_dict keys) against the oracle."""
import numpy as np
import pytest
import torch

from oracle import mesm_oracle as mo
from oracle.config import CONFIGS
from tests.helpers import engine_cfg, load_case, rel_err

pytestmark = pytest.mark.gpu


def _build(cfg, sd):
    from mesm_b200.model import build_model
    m = build_model(engine_cfg(cfg))
    m.load_state_dict(sd, strict=True)          # reference key names load as-is
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["tiny_ragged", "tiny_qvh_groups", "tiny_twomlp", "qvh_groups"])
def test_mesm_module_is_a_drop_in(name):
    cfg, sd, inp, neg, gold, meta = load_case(name)
    m = _build(cfg, sd)
    out = m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), inp["words_feat"].cuda(), None, None, inp["num_clips"],
            dataset_name=cfg.dataset_name, is_training=False, neg_index=neg.cuda(), qid=[0], sentence=["ignored kwargs"])
    assert set(out) == {"pred_logits", "pred_spans", "saliency_scores", "neg_saliency_scores", "aux_outputs", "projed_video_feat",
                        "recon_feat", "projed_recon_feat", "expanded_words_feat", "expanded_words_mask", "enhanced_video_feat",
                        "projed_words_feat"}                                   # model/model.py:334-351
    vm = inp["video_mask"]
    assert rel_err(out["pred_logits"], gold["pred_logits"]) <= 1e-3
    assert rel_err(out["pred_spans"], gold["pred_spans"]) <= 1e-3
    assert rel_err(out["saliency_scores"], gold["saliency_scores"], vm) <= 1e-3
    assert rel_err(out["neg_saliency_scores"], gold["neg_saliency_scores"], vm) <= 1e-3
    assert rel_err(out["aux_outputs"][0]["pred_spans"], gold["aux_spans"]) <= 1e-3
    assert rel_err(out["projed_words_feat"], gold["projed_words_feat"]) <= 1e-3
    # without an injected neg_index the module samples one (other video group) and still runs
    out2 = m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), inp["words_feat"].cuda(), None, None, inp["num_clips"],
             dataset_name=cfg.dataset_name, is_training=False)
    assert torch.equal(out2["pred_logits"], out["pred_logits"])                 # independent of the negative draw
    # host clip counts in the batch (mesm_b200.prepare_batch_input adds them) -> packed rows, same results
    from mesm_b200.ingest import clip_counts
    out3 = m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), inp["words_feat"].cuda(), None, None, inp["num_clips"],
             dataset_name=cfg.dataset_name, is_training=False, neg_index=neg.cuda(), video_len=clip_counts(inp["video_mask"]))
    assert rel_err(out3["pred_logits"], out["pred_logits"]) < 1e-4 and rel_err(out3["pred_spans"], out["pred_spans"]) < 1e-4
    assert rel_err(out3["saliency_scores"], out["saliency_scores"], vm) < 1e-4
    with pytest.raises(NotImplementedError):
        m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), inp["words_feat"].cuda(), None, None, inp["num_clips"],
          dataset_name=cfg.dataset_name, is_training=True)
    with pytest.raises(RuntimeError):                                            # CPU tensors: no fallback
        m(inp["video_feat"], inp["video_mask"], inp["words_feat"], None, None, inp["num_clips"], dataset_name=cfg.dataset_name,
          is_training=False)


def test_t2v_encoder_and_transformer_submodules():
    cfg = CONFIGS["tiny"]
    _, sd, inp, neg, gold, meta = load_case("tiny_ragged")
    m = _build(cfg, sd)
    g = torch.Generator().manual_seed(5)
    B, Lv, Lt = 6, 24, 9
    txt, vid = torch.randn(B, Lt, 256, generator=g), torch.randn(B, Lv, 256, generator=g)
    pos_t, pos_v = torch.randn(B, Lt, 256, generator=g) * 0.1, torch.randn(B, Lv, 256, generator=g) * 0.1
    tl = torch.randint(2, Lt + 1, (B,), generator=g); vl = torch.randint(5, Lv + 1, (B,), generator=g)
    tpad = torch.arange(Lt)[None] >= tl[:, None]; vpad = torch.arange(Lv)[None] >= vl[:, None]
    ref = mo.t2v_stack(sd, "t2v_encoder.t2v_encoder", cfg.t2v_layers, txt, vid, tpad, pos_t, vpad, pos_v, cfg.nheads)
    out = m.t2v_encoder(txt.cuda(), vid.cuda(), src_txt_key_padding_mask=tpad.cuda(), pos_txt=pos_t.cuda(),
                        src_vid_key_padding_mask=vpad.cuda(), pos_vid=pos_v.cuda())
    assert rel_err(out, ref) <= 1e-4
    # Transformer.forward (model/transformer.py:174-205)
    qe = sd["query_embed.weight"]
    gt, gp = sd["global_rep_token"], sd["global_rep_pos"]
    osd = dict(sd)
    hs, refs, mem, memg = mo.transformer(osd, cfg, vid, vpad, qe, pos_v)
    o = m.transformer(vid.cuda(), vpad.cuda(), qe.cuda(), pos_v.cuda(), gt.view(1, 1, -1).repeat(B, 1, 1).cuda(),
                      gp.view(1, 1, -1).repeat(B, 1, 1).cuda())
    assert rel_err(o[0], hs) <= 1e-4 and rel_err(o[1], refs) <= 1e-4
    assert rel_err(o[2], mem, (~vpad)[..., None]) <= 1e-4 and rel_err(o[3], memg) <= 1e-4


def test_projection_free_multihead_attention():
    """model/attention.py:61-182 (the decoder's MHA): q/k dims differ from the value dim, seq-first tensors."""
    from mesm_b200.model import MultiheadAttention
    g = torch.Generator().manual_seed(1)
    L, S, B, E, Ev, H = 10, 37, 5, 512, 256, 8
    mha = MultiheadAttention(E, H, vdim=Ev).cuda()
    with torch.no_grad():
        mha.out_proj.bias.normal_(0, 0.1)
    q, k, v = torch.randn(L, B, E, generator=g), torch.randn(S, B, E, generator=g), torch.randn(S, B, Ev, generator=g)
    pad = torch.arange(S)[None] >= torch.randint(3, S + 1, (B, 1), generator=g)
    out, w = mha(q.cuda(), k.cuda(), v.cuda(), key_padding_mask=pad.cuda())
    sd = {"out_proj.weight": mha.out_proj.weight.detach().cpu(), "out_proj.bias": mha.out_proj.bias.detach().cpu()}
    ref = mo.plain_mha(sd, "", q.permute(1, 0, 2), k.permute(1, 0, 2), v.permute(1, 0, 2), pad, H).permute(1, 0, 2)
    assert rel_err(out, ref) <= 1e-5
    qh = (q.permute(1, 0, 2) * (E // H) ** -0.5).reshape(B, L, H, E // H).permute(0, 2, 1, 3)
    kh = k.permute(1, 0, 2).reshape(B, S, H, E // H).permute(0, 2, 1, 3)
    p = torch.softmax((qh @ kh.transpose(-1, -2)).masked_fill(pad[:, None, None, :], float("-inf")), -1).mean(1)
    assert rel_err(w, p) <= 1e-5


def test_alignment_scores_match_reference_formula():
    import mesm_b200
    for name in ("tiny_ragged", "qvh_groups"):
        cfg, sd, inp, neg, gold, meta = load_case(name)
        o = mo.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
        S = mesm_b200.align_scores(o["projed_video_feat"].cuda(), torch.from_numpy(gold["align_clip_mask"]).cuda(),
                                   o["expanded_words_feat"].cuda(), o["expanded_words_mask"].cuda(), cfg.recss_tau)
        assert rel_err(S, gold["align_scores"]) <= 1e-5
