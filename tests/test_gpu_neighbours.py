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
