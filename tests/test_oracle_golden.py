"""CPU: the oracle reproduces the committed REFERENCE outputs (tests/golden/, written by oracle/gen_golden.py from
the real lntzm/MESM modules) — this is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import decode_oracle, mesm_oracle
from tests.helpers import golden_cases, load_case

FAST = ["tiny_uniform", "tiny_ragged", "tiny_qvh_groups", "tiny_twomlp", "qvh_groups", "charades_vgg_l64"]


@pytest.mark.parametrize("name", FAST)
def test_forward_oracle_matches_reference(name):
    cfg, sd, inp, neg, gold, meta = load_case(name)
    assert np.array_equal(neg.numpy(), gold["neg_index"])
    o = mesm_oracle.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
    vm = inp["video_mask"].numpy()
    for k in ("pred_logits", "pred_spans", "recon_feat", "projed_recon_feat", "projed_words_feat"):
        assert np.abs(o[k].numpy() - gold[k]).max() < 2e-5, k
    for k in ("saliency_scores", "neg_saliency_scores"):
        assert (np.abs(o[k].numpy() - gold[k]) * vm).max() < 2e-5, k
    assert np.abs(o["aux_outputs"][0]["pred_spans"].numpy() - gold["aux_spans"]).max() < 2e-5
    S = mesm_oracle.align_scores(o["projed_video_feat"], torch.from_numpy(gold["align_clip_mask"]),
                                 o["expanded_words_feat"], o["expanded_words_mask"], cfg.recss_tau)
    assert np.abs(S.numpy() - gold["align_scores"]).max() < 2e-5


@pytest.mark.parametrize("name", sorted(golden_cases()))
def test_decode_oracle_matches_reference(name):
    cfg, sd, inp, neg, gold, meta = load_case(name)
    for i in range(gold["pred_logits"].shape[0]):
        od = decode_oracle.decode_pair(gold["pred_logits"][i], gold["pred_spans"][i], float(inp["duration"][i]), cfg.clip_len,
                                       cfg.max_ts_val, meta["nms_thd"], 10, 10)
        assert od["order"] == gold["order"][i].tolist()
        assert np.array_equal(np.asarray(od["windows"]), gold["windows"][i])
        n = int(gold["nms_count"][i])
        assert len(od["nms_windows"]) == n and np.array_equal(np.asarray(od["nms_windows"]), gold["nms_windows"][i, :n])


def test_span_utils_doctest_vectors(golden_dir):
    """The reference's only golden numbers: utils/span_utils.py:12-19, 31-38, 54-60, 105-109."""
    import os
    g = np.load(os.path.join(golden_dir, "span_utils_doctest.npz"))
    iou, union = decode_oracle.temporal_iou(g["s1"], g["s2"])
    assert np.array_equal(iou.astype(np.float32), g["iou"]) and np.array_equal(union.astype(np.float32), g["union"])
    assert np.allclose(g["iou"], [[0.6667, 0.2], [0.0, 0.5]], atol=5e-5) and np.allclose(g["union"], [[0.3, 1.0], [0.8, 1.0]])
    assert np.array_equal(decode_oracle.generalized_temporal_iou(g["s1"], g["s2"]), g["giou"])
    assert np.allclose(g["giou"], [[0.6667, 0.2], [-0.2, 0.5]], atol=5e-5)
    assert np.array_equal(decode_oracle.span_xx_to_cxw(np.float32([[0, 1], [0.2, 0.4]])), g["cxw"])
    assert np.array_equal(decode_oracle.span_cxw_to_xx(np.float32([[0.5, 1.0], [0.3, 0.2]])), g["xx"])


def test_nms_probe_cases():
    """IoU exactly at the threshold is kept, zero-length duplicates are both kept, three
    identical windows leave one survivor, a single prediction is returned untouched."""
    nms = decode_oracle.temporal_nms
    assert nms([[0, 10, .9], [3, 10, .8]], 0.7)[1] == [0, 1]              # hull IoU = 7/10, not > 0.7
    assert nms([[5, 5, .9], [5, 5, .8]], 0.7)[1] == [0, 1]                # hull == 0 -> IoU 0
    assert nms([[1, 4, .9], [1, 4, .8], [1, 4, .7]], 0.7)[1] == [0]
    assert nms([[1, 4, .3]], 0.7, max_after_nms=0)[1] == [0]
    assert nms([[0, 1, .2], [5, 6, .9], [10, 11, .5]], 0.7, max_after_nms=2)[1] == [1, 2]


def test_decode_oracle_matches_reference_random_fixture(golden_dir):
    """tests/golden/decode_random.npz (oracle/gen_golden.py: gen_decode_fixture): 3000 random candidate lists through the
    REFERENCE's utils.temporal_nms and 4 x 300 random decode chains through the eval.py loop with the reference's own
    span_cxw_to_xx / PostProcessorDETR / temporal_nms - the in-repo pin of oracle/decode_oracle.py."""
    import os
    g = np.load(os.path.join(golden_dir, "decode_random.npz"))
    offs, koffs = g["nms_offsets"], g["nms_kept_offsets"]
    for i in range(len(offs) - 1):
        w = g["nms_lists"][offs[i]:offs[i + 1]]
        thd, na = float(g["nms_params"][i, 0]), int(g["nms_params"][i, 1])
        kept_w, kept_pos = decode_oracle.temporal_nms(w.tolist(), thd, na)
        ref = g["nms_kept"][koffs[i]:koffs[i + 1]]
        assert len(kept_w) == len(ref) and np.array_equal(np.asarray(kept_w).reshape(-1, 3), ref), i
        assert np.array_equal(w[kept_pos], ref), i
    from oracle.config import CONFIGS
    for cname in ("qvhighlights", "charades_csf", "charades_vgg", "tacos"):
        cfg = CONFIGS[cname]
        lg, sp, dur = g[f"dec_{cname}_logits"], g[f"dec_{cname}_spans"], g[f"dec_{cname}_duration"]
        for i in range(lg.shape[0]):
            od = decode_oracle.decode_pair(lg[i], sp[i], float(dur[i]), cfg.clip_len, cfg.max_ts_val, 0.7, 10, 10)
            assert od["order"] == g[f"dec_{cname}_order"][i].tolist(), (cname, i)
            assert np.array_equal(np.asarray(od["windows"]), g[f"dec_{cname}_windows"][i]), (cname, i)
            n = int(g[f"dec_{cname}_nms_count"][i])
            assert np.array_equal(np.asarray(od["nms_windows"]).reshape(-1, 3), g[f"dec_{cname}_nms_windows"][i, :n]), (cname, i)
