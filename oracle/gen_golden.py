"""Pin the oracle against the real reference and write the golden fixtures under tests/golden/.

Runs ONLY in the build container (needs /root/reference).  TEST INFRASTRUCTURE (see oracle/__init__.py).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

For every case: build ``model.MESM`` exactly as runner.py:255-298 does (Appendix B of SURVEY.md; ``runner`` itself is
not importable here: it pulls ftfy/h5py/nltk), load ``oracle.weights.make_state_dict`` with strict=True (which also
proves ``state_spec`` matches the reference key set and shapes), run the reference forward with
``sample_outclass_neg`` patched to the injected neg_index, run the oracle, assert agreement, and save the REFERENCE
outputs.  The decode chain is the reference's own ``span_cxw_to_xx`` / ``PostProcessorDETR`` / ``temporal_nms`` driven
as eval.py:64-99,111-116,476-485 drives them.
"""
import json
import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("MESM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")

from oracle.config import CONFIGS  # noqa: E402
from oracle import mesm_oracle, decode_oracle  # noqa: E402
from oracle.weights import make_state_dict, make_inputs, make_neg_index, state_spec  # noqa: E402

CASES = [
    # name, config, num_clips, kwargs for make_inputs
    ("tiny_uniform", "tiny", [1] * 6, dict(ragged_video=False, ragged_text=False)),
    ("tiny_ragged", "tiny", [2, 1, 3, 1, 2, 4, 1, 2], dict()),
    ("tiny_qvh_groups", "tiny_qvh", [2, 1, 3, 2], dict()),
    ("tiny_twomlp", "tiny_twomlp", [3, 2, 4], dict()),
    ("qvh_b6", "qvhighlights", [1] * 6, dict(ragged_video=False)),
    ("qvh_groups", "qvhighlights", [2, 1, 2], dict()),
    ("charades_csf_ragged", "charades_csf", [2, 3, 1, 2], dict()),
    ("charades_vgg_l64", "charades_vgg", [2, 2, 1], dict(lv=64)),
    ("tacos_l96", "tacos", [4, 3], dict(lv=96)),
    # benchmark shapes of BASELINE configs[2] / [3] (Lv = 200, K = 4098, 201 encoder keys; TACoS
    # groups of up to 10 queries) and the shipped max_video_l = 600 (beyond the 224-key tcgen05 attention tile)
    ("charades_vgg_l200", "charades_vgg", [2, 1, 2], dict(lv=200)),
    ("tacos_l200", "tacos", [10, 3, 4], dict(lv=200)),
    ("charades_vgg_l600", "charades_vgg", [2, 1], dict(lv=600)),
    # 16-bit feature storage: the reference fed the fp16-representable values, upcast to fp32
    ("charades_csf_ragged_f16", "charades_csf", [2, 3, 1, 2], dict(f16_features=True)),
    ("tacos_l96_f16", "tacos", [4, 3], dict(lv=96, f16_features=True)),
]
NMS_THD = 0.7


from oracle.ref_harness import build_reference, reference_decode  # noqa: E402


def gen_decode_fixture():
    """Random differential fixture of the decode chain: the REFERENCE's utils.temporal_nms on random candidate lists
    (ties, duplicates, zero-length windows, several thresholds / max_after_nms) and the eval.py decode loop
    (reference_decode above) on random logits / spans for every shipped (clip_len, max_ts) setting.  Inputs are
    regenerated from the seeds by the tests; only the reference's answers are stored."""
    from utils import temporal_nms
    rng = np.random.default_rng(2024)
    lists, answers, params = [], [], []
    for case in range(3000):
        n = int(rng.choice([1, 2, 3, 5, 10, 10, 10, 17, 40, 100]))
        st = rng.uniform(0, 140, n).round(int(rng.integers(0, 3)))
        w = np.stack([st, st + rng.uniform(0, 30, n).round(int(rng.integers(0, 3))), rng.uniform(0, 1, n).round(4)], 1)
        if n >= 3:
            if case % 3 == 0:
                w[1] = w[0]                                   # exact duplicate
            if case % 4 == 0:
                w[2, 2] = w[0, 2]                             # score tie
            if case % 5 == 0:
                w[n - 1, 1] = w[n - 1, 0]                     # zero-length window
            if case % 7 == 0:
                w[:, :2] = w[0, :2]                           # all identical spans
        thd = float(rng.choice([0.3, 0.5, 0.7, 0.9]))
        na = int(rng.choice([1, 3, 10, 100]))
        kept = temporal_nms([list(map(float, r)) for r in w], nms_thd=thd, max_after_nms=na)
        lists.append(w)
        answers.append(np.asarray(kept, dtype=np.float64).reshape(-1, 3))
        params.append((thd, na))
    out = dict(nms_lists=np.concatenate(lists), nms_offsets=np.concatenate([[0], np.cumsum([len(l) for l in lists])]).astype(np.int64),
               nms_params=np.asarray(params, dtype=np.float64),
               nms_kept=np.concatenate(answers), nms_kept_offsets=np.concatenate([[0], np.cumsum([len(a) for a in answers])]).astype(np.int64))
    for ci, cname in enumerate(("qvhighlights", "charades_csf", "charades_vgg", "tacos")):
        cfg = CONFIGS[cname]
        B, nq = 300, cfg.num_queries
        lg = (rng.normal(size=(B, nq, 2)) * 2).astype(np.float32)
        lg[::7, 3] = lg[::7, 5]
        sp = np.stack([rng.uniform(0, 1, (B, nq)), rng.uniform(0, 0.6, (B, nq))], -1).astype(np.float32)
        sp[::5, 2] = sp[::5, 4]
        sp[::11, 6, 1] = 0
        dur = rng.uniform(5, 150 if cfg.max_ts_val <= 150 else 900, B).astype(np.float32)
        rd = reference_decode(torch.from_numpy(lg), torch.from_numpy(sp), torch.from_numpy(dur), cfg, NMS_THD)
        nms_n = np.array([len(r["nms_windows"]) for r in rd], dtype=np.int32)
        nms_w = np.zeros((B, 10, 3))
        for i, r in enumerate(rd):
            nms_w[i, :nms_n[i]] = np.asarray(r["nms_windows"], dtype=np.float64)
        out.update({f"dec_{cname}_logits": lg, f"dec_{cname}_spans": sp, f"dec_{cname}_duration": dur,
                    f"dec_{cname}_windows": np.array([r["windows"] for r in rd], dtype=np.float64),
                    f"dec_{cname}_order": np.array([r["order"] for r in rd], dtype=np.int32),
                    f"dec_{cname}_nms_windows": nms_w, f"dec_{cname}_nms_count": nms_n})
    np.savez_compressed(os.path.join(GOLD, "decode_random.npz"), **out)
    print("decode_random.npz:", len(lists), "NMS lists,", 4 * 300, "decode chains")


def main():
    os.makedirs(GOLD, exist_ok=True)
    sys.path.insert(0, REF)
    if "--decode-only" in sys.argv:
        return gen_decode_fixture()
    summary = {}
    for name, cfg_name, num_clips, kw in CASES:
        cfg = CONFIGS[cfg_name]
        torch.manual_seed(0)
        ref = build_reference(cfg)
        sd = make_state_dict(cfg, seed=1)
        ref_sd = ref.state_dict()
        assert set(ref_sd) == set(sd), (sorted(set(ref_sd) ^ set(sd)))
        for k, shp in state_spec(cfg):
            assert tuple(ref_sd[k].shape) == tuple(shp), (k, ref_sd[k].shape, shp)
        ref.load_state_dict(sd, strict=True)
        inp = make_inputs(cfg, num_clips, seed=2, **kw)
        neg = make_neg_index(num_clips, seed=3)
        import model.model as mm
        mm.sample_outclass_neg = lambda nc, _neg=neg: _neg
        with torch.inference_mode():
            ro = ref(video_feat=inp["video_feat"], video_mask=inp["video_mask"], words_id=inp["words_feat"].clone(),
                     words_mask=None, words_weight=None, num_clips=inp["num_clips"],
                     dataset_name=cfg.dataset_name, is_training=False)
        oo = mesm_oracle.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"],
                                      neg_index=neg)
        vm = inp["video_mask"]
        errs = {}
        for k in ("pred_logits", "pred_spans", "recon_feat", "projed_recon_feat", "projed_video_feat",
                  "projed_words_feat", "enhanced_video_feat", "expanded_words_feat"):
            errs[k] = float((ro[k] - oo[k]).abs().max())
        for k in ("saliency_scores", "neg_saliency_scores"):
            errs[k] = float(((ro[k] - oo[k]) * vm).abs().max())
        errs["aux_logits"] = float((ro["aux_outputs"][0]["pred_logits"] - oo["aux_outputs"][0]["pred_logits"]).abs().max())
        errs["aux_spans"] = float((ro["aux_outputs"][0]["pred_spans"] - oo["aux_outputs"][0]["pred_spans"]).abs().max())
        assert bool((ro["expanded_words_mask"] == oo["expanded_words_mask"]).all())
        worst = max(errs.values())
        print(f"{name:24s} B={sum(num_clips):3d} max|ref-oracle|={worst:.2e}  " +
              " ".join(f"{k}={v:.1e}" for k, v in errs.items() if v > 5e-6))
        assert worst < 2e-5, errs
        # alignment scores: reference formula (model/criterion.py:241-266) evaluated with torch on reference outputs
        g = torch.Generator().manual_seed(11)
        clip_mask = (torch.rand(vm.shape, generator=g) < 0.4) & vm
        clip_mask[:, 0] = True
        import torch.nn.functional as F
        cm = clip_mask.unsqueeze(-1)
        cf = (ro["projed_video_feat"] * cm).sum(1) / cm.sum(1)
        wm = ro["expanded_words_mask"].unsqueeze(-1)
        wf = (ro["expanded_words_feat"] * wm).sum(1) / wm.sum(1)
        S_ref = F.normalize(cf, dim=-1, p=2) @ F.normalize(wf, dim=-1, p=2).permute(1, 0) / cfg.recss_tau
        S_or = mesm_oracle.align_scores(oo["projed_video_feat"], clip_mask, oo["expanded_words_feat"],
                                        oo["expanded_words_mask"], cfg.recss_tau)
        assert float((S_ref - S_or).abs().max()) < 2e-5
        # decode chain
        rd = reference_decode(ro["pred_logits"], ro["pred_spans"], inp["duration"], cfg, NMS_THD)
        win = np.array([r["windows"] for r in rd], dtype=np.float64)
        order = np.array([r["order"] for r in rd], dtype=np.int32)
        nms_n = np.array([len(r["nms_windows"]) for r in rd], dtype=np.int32)
        nms_w = np.zeros((len(rd), 10, 3), dtype=np.float64)
        for i, r in enumerate(rd):
            nms_w[i, :nms_n[i]] = np.array(r["nms_windows"], dtype=np.float64)
        lg, sp = ro["pred_logits"].numpy(), ro["pred_spans"].numpy()
        for i in range(len(rd)):
            od = decode_oracle.decode_pair(lg[i], sp[i], float(inp["duration"][i]), cfg.clip_len, cfg.max_ts_val,
                                           NMS_THD, 10, 10)
            assert od["order"] == rd[i]["order"], (name, i)
            assert od["windows"] == rd[i]["windows"], (name, i, od["windows"], rd[i]["windows"])
            assert od["nms_windows"] == rd[i]["nms_windows"], (name, i)
        np.savez_compressed(
            os.path.join(GOLD, f"{name}.npz"),
            pred_logits=lg, pred_spans=sp, saliency_scores=ro["saliency_scores"].numpy(),
            neg_saliency_scores=ro["neg_saliency_scores"].numpy(), recon_feat=ro["recon_feat"].numpy(),
            projed_recon_feat=ro["projed_recon_feat"].numpy(),
            aux_logits=ro["aux_outputs"][0]["pred_logits"].numpy(), aux_spans=ro["aux_outputs"][0]["pred_spans"].numpy(),
            projed_video_row0=ro["projed_video_feat"][:, 0].numpy(), enhanced_video_row0=ro["enhanced_video_feat"][:, 0].numpy(),
            projed_words_feat=ro["projed_words_feat"].numpy(),
            align_clip_mask=clip_mask.numpy(), align_scores=S_ref.numpy(),
            windows=win, order=order, nms_windows=nms_w, nms_count=nms_n, neg_index=neg.numpy())
        summary[name] = dict(config=cfg_name, num_clips=num_clips, kwargs=kw, weight_seed=1, input_seed=2, neg_seed=3,
                             nms_thd=NMS_THD, max_abs_ref_minus_oracle=worst)
    # doctest vectors of utils/span_utils.py (the reference's only golden numbers), evaluated by the reference itself
    from utils import span_utils as su
    s1, s2 = torch.Tensor([[0, 0.2], [0.5, 1.0]]), torch.Tensor([[0, 0.3], [0., 1.0]])
    iou, union = su.temporal_iou(s1, s2)
    np.savez(os.path.join(GOLD, "span_utils_doctest.npz"), s1=s1.numpy(), s2=s2.numpy(), iou=iou.numpy(),
             union=union.numpy(), giou=su.generalized_temporal_iou(s1, s2).numpy(),
             cxw=su.span_xx_to_cxw(torch.Tensor([[0, 1], [0.2, 0.4]])).numpy(),
             xx=su.span_cxw_to_xx(torch.Tensor([[0.5, 1.0], [0.3, 0.2]])).numpy())
    with open(os.path.join(GOLD, "cases.json"), "w") as f:
        json.dump(summary, f, indent=1)
    gen_decode_fixture()
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
