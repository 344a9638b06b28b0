"""GPU parity of the path's neighbours against fixtures produced by the reference's own code
(oracle/gen_golden_aux.py): the eval-time saliency criterion and the feature-ingest front-end."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["charades", "qvh", "one_row"])
def test_loss_saliency_matches_reference(golden_dir, name):
    """Criterion.loss_saliency (model/criterion.py:139-221): charades-style 0/1 labels; QVHighlights-style integer labels with the
    triplet term; fp32 reductions in a different order -> 2e-5 relative."""
    import mesm_b200
    g = np.load(os.path.join(golden_dir, "criterion_saliency.npz"))
    t = lambda k: torch.from_numpy(g[f"crit_{name}_{k}"]).cuda()
    outputs = dict(saliency_scores=t("sal"), neg_saliency_scores=t("neg"))
    qvh = name == "qvh"
    targets = dict(video_mask=t("mask"))
    if qvh:
        targets.update(saliency_label=t("label"), pos_idx=t("pos_idx"), neg_idx=t("neg_idx"))
    else:
        targets.update(clip_mask=t("label") > 0)
    r = mesm_b200.loss_saliency(outputs, targets, rank_coef=12, use_triplet=qvh, saliency_margin=0.2)
    ref = float(g[f"crit_{name}_loss"])
    assert abs(float(r["loss_saliency"]) - ref) <= 2e-5 * abs(ref), (float(r["loss_saliency"]), ref)
    assert abs(float(r["loss_neg_pair"] + r["loss_rank_contrastive"] + r["loss_triplet"]) - float(r["loss_saliency"])) < 1e-4
    assert (float(r["loss_triplet"]) > 0) == qvh


@pytest.mark.parametrize("name", ["csf_short", "csf_pool", "vgg_pool", "c3d_exact"])
def test_feature_frontend_matches_reference(golden_dir, name):
    """get_video_feat + sample_video_feat + add_tef (dataset/charades.py:108-119, dataset/base.py:100-114, 225-230) on raw fp32 /
    fp16 per-source arrays: normalise, truncate to the shortest source, concatenate, mean-pool to max_video_l, tef columns."""
    import mesm_b200
    from oracle.weights import make_raw_features
    g = np.load(os.path.join(golden_dir, "frontend.npz"))
    raws, max_l = make_raw_features(name)
    ref = torch.from_numpy(g[f"fe_{name}_out"])
    out = mesm_b200.build_video_feat(raws, max_l)
    assert out.shape == ref.shape and out.dtype == torch.float32
    o = out.cpu()
    assert torch.equal(o[:, -2:], ref[:, -2:])                                  # tef columns: bit-exact fp32 arithmetic
    assert float((o - ref).abs().max()) <= 2e-6 * float(ref.abs().max()) + 1e-7
    h = mesm_b200.build_video_feat([r.cuda() for r in raws], max_l, out_dtype=torch.float16)      # 16-bit storage of the same rows
    assert h.dtype == torch.float16 and float((h.float().cpu() - ref).abs().max()) <= 1e-3 * float(ref.abs().max())
    assert torch.equal(h, out.half())


@pytest.mark.parametrize("name", ["charades", "qvh", "tacos"])
def test_moment_retrieval_metrics_match_reference(golden_dir, name):
    """eval_moment_retrieval (eval.py:233-263: MR-R1 / MR-mAP per ground-truth length range, multi-window ground truth, duplicate
    predictions, equal scores) - the reference's own functions produced tests/golden/metrics_expected.json."""
    import json
    from mesm_b200 import utils as U
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    exp = json.load(open(os.path.join(golden_dir, "metrics_expected.json")))[name]
    res = U.eval_moment_retrieval(torch.from_numpy(g[f"met_{name}_windows"]).cuda(), torch.from_numpy(g[f"met_{name}_gt"]).cuda(),
                                  torch.from_numpy(g[f"met_{name}_gt_off"]).cuda(), dataset_name=exp["dataset_name"])
    ref = exp["metrics"]
    assert sorted(res) == sorted(ref)
    for rng_name in ref:
        for fam in ("MR-mAP", "MR-R1"):
            assert sorted(res[rng_name][fam]) == sorted(ref[rng_name][fam]), (rng_name, fam)
            for k, v in ref[rng_name][fam].items():
                assert res[rng_name][fam][k] == v, (rng_name, fam, k, res[rng_name][fam][k], v)       # 2-decimal percents: exact


def test_clip_text_tower_matches_reference(golden_dir):
    """CLIPTextEncoder (model/text_encoder.py:240-354; 12 x 512, 8 heads, 77 tokens, causal): our fp32 / bf16x3 run against the
    reference module evaluated in fp32 (tight) and as shipped in fp16 (the reference's own rounding: 1.7e-3 from its fp32 run)."""
    from mesm_b200.model import CLIPTextEncoder
    from oracle.weights import CLIP_TEXT_CFG, make_clip_state_dict, make_clip_tokens
    g = np.load(os.path.join(golden_dir, "clip_text.npz"))
    enc = CLIPTextEncoder(**CLIP_TEXT_CFG)
    enc.load_state_dict(make_clip_state_dict(3), strict=True)
    enc = enc.cuda().eval()
    out = enc(make_clip_tokens(5, 4).cuda())
    h, p = out["last_hidden_state"][:, :32].cpu(), out["pooler_output"].cpu()
    assert out["last_hidden_state"].shape == (5, 77, 512) and not torch.isnan(out["last_hidden_state"]).any()
    rel = lambda a, b: float((a.double() - torch.from_numpy(b).double()).abs().max() / np.abs(b).max())
    assert rel(h, g["clip_hidden_f32"]) < 1e-4 and rel(p, g["clip_pooled_f32"]) < 1e-4
    assert rel(h, g["clip_hidden_f16"].astype(np.float32)) < 5e-3 and rel(p, g["clip_pooled_f16"].astype(np.float32)) < 5e-3


def test_mesm_forward_with_clip_text_encoder():
    """MESM with a CLIPTextEncoder (the C+SF configs): words_id are token ids; equals feeding the tower's masked hidden states as word
    features (CLIP_encode_text, model/model.py:103-125)."""
    from mesm_b200.model import CLIPTextEncoder, build_model
    from oracle.config import CONFIGS
    from oracle.weights import CLIP_TEXT_CFG, make_clip_state_dict, make_clip_tokens, make_inputs, make_state_dict
    from tests.helpers import engine_cfg
    cfg = CONFIGS["qvhighlights"]
    nc = [2, 1, 2]
    inp = make_inputs(cfg, nc, 2)
    m = build_model(engine_cfg(cfg))
    m.load_state_dict(make_state_dict(cfg, 1), strict=True)
    enc = CLIPTextEncoder(**CLIP_TEXT_CFG)
    enc.load_state_dict(make_clip_state_dict(3), strict=True)
    m = m.cuda().eval()
    enc = enc.cuda().eval()
    text = make_clip_tokens(5, 4).cuda()
    wmask = (text != 0)
    feats = enc(text)["last_hidden_state"][:, :cfg.max_words_l].masked_fill(~wmask[:, :cfg.max_words_l, None], 0)
    a = m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), feats, None, None, inp["num_clips"], dataset_name="qvhighlights", is_training=False,
          neg_index=torch.tensor([2, 3, 0, 0, 1]).cuda())
    m.text_encoder = enc
    b = m(inp["video_feat"].cuda(), inp["video_mask"].cuda(), text, wmask, None, inp["num_clips"], dataset_name="qvhighlights", is_training=False,
          neg_index=torch.tensor([2, 3, 0, 0, 1]).cuda())
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        assert torch.equal(a[k], b[k]), k
