"""Golden fixtures of the path's neighbours: the eval-time saliency criterion and the feature-ingest front-end.

Runs ONLY in the build container (needs /root/reference).  TEST INFRASTRUCTURE (see oracle/__init__.py).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_aux

* criterion: the reference's own ``Criterion.loss_saliency`` (model/criterion.py:139-221) called unbound on seeded tensors
  (charades-style 0/1 clip-mask labels; QVHighlights-style integer saliency labels + the triplet term).
* front-end: ``dataset`` is not importable offline (h5py / ftfy), so the three functions are evaluated from their SOURCE TEXT:
  the bodies of ``sample_video_feat`` / ``add_tef`` (dataset/base.py:100-114, 225-230) are exec'ed from the reference file, and
  ``get_video_feat`` (dataset/charades.py:108-119) is restated around the same ``F.normalize`` / ``torch.cat`` calls.
"""
import os
import re
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn.functional as F

sys.dont_write_bytecode = True
REF = os.environ.get("MESM_REFERENCE", "/root/reference")
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def criterion_cases():
    sys.path.insert(0, REF)
    from model.criterion import Criterion
    out = {}
    g = torch.Generator().manual_seed(99)
    for name, B, L, qvh in (("charades", 37, 61, False), ("qvh", 24, 75, True), ("one_row", 1, 9, False)):
        lens = torch.randint(max(2, L // 2), L + 1, (B,), generator=g)
        lens[0] = L
        mask = torch.arange(L)[None] < lens[:, None]
        sal = torch.randn(B, L, generator=g) * 2
        neg = torch.randn(B, L, generator=g) * 2
        if qvh:
            label = (torch.randint(0, 13, (B, L), generator=g) * (torch.rand(B, L, generator=g) < 0.3)).float() * mask
            label[3] = 0                                           # a row without positives
            pos_idx = torch.stack([torch.randint(0, int(n), (2,), generator=g) for n in lens])
            neg_idx = torch.stack([torch.randint(0, int(n), (2,), generator=g) for n in lens])
            targets = dict(video_mask=mask, saliency_label=label, pos_idx=pos_idx, neg_idx=neg_idx)
        else:
            label = ((torch.rand(B, L, generator=g) < 0.3) & mask)
            targets = dict(video_mask=mask, clip_mask=label)
        fake = types.SimpleNamespace(rank_coef=12, use_triplet=qvh, saliency_margin=0.2)
        ref = Criterion.loss_saliency(fake, dict(saliency_scores=sal, neg_saliency_scores=neg), targets, None)["loss_saliency"]
        out.update({f"crit_{name}_sal": sal.numpy(), f"crit_{name}_neg": neg.numpy(), f"crit_{name}_mask": mask.numpy(),
                    f"crit_{name}_label": (targets.get("saliency_label", targets.get("clip_mask"))).float().numpy(),
                    f"crit_{name}_loss": np.float32(ref)})
        if qvh:
            out.update({f"crit_{name}_pos_idx": pos_idx.numpy(), f"crit_{name}_neg_idx": neg_idx.numpy()})
        print(f"criterion {name}: loss_saliency = {float(ref):.6f}")
    return out


def _method_source(path, name):
    src = open(path).read()
    m = re.search(r"^    def %s\(self.*?(?=^    def |^\S|\Z)" % name, src, re.S | re.M)
    return textwrap.dedent(m.group(0))


def frontend_cases():
    ns = {"torch": torch}
    exec(_method_source(os.path.join(REF, "dataset", "base.py"), "sample_video_feat"), ns)
    exec(_method_source(os.path.join(REF, "dataset", "base.py"), "add_tef"), ns)
    from oracle.weights import FRONTEND_CASES, make_raw_features
    out = {}
    for name in FRONTEND_CASES:
        raws, max_l = make_raw_features(name)
        feats = [F.normalize(torch.from_numpy(r.numpy().astype(np.float32)), dim=1) for r in raws]      # charades.py:112-116
        min_len = min(len(e) for e in feats)
        feat = torch.cat([e[:min_len] for e in feats], dim=1)     # :117-119
        fake = types.SimpleNamespace(max_video_l=max_l)
        feat = ns["sample_video_feat"](fake, feat)                # base.py:168
        feat = ns["add_tef"](fake, feat.shape[0], feat)           # base.py:169-171
        out[f"fe_{name}_out"] = feat.numpy()
        print(f"front-end {name}: raw {[tuple(r.shape) for r in raws]} -> {tuple(feat.shape)}")
    return out


def _function_source(path, name):
    src = open(path).read()
    m = re.search(r"^def %s\(.*?(?=^def |^if __name__|\Z)" % name, src, re.S | re.M)
    return m.group(0)


def metrics_cases():
    """eval_moment_retrieval (eval.py:233-263) and everything below it, exec'ed from the reference's eval.py (the module itself
    imports runner / dataset, which need ftfy / h5py).  One harness change: compute_mr_ap's multiprocessing pool is switched off
    (num_workers=8 -> 1) because exec'ed functions cannot be pickled; the per-query code is untouched."""
    import copy
    import json
    import time
    from collections import OrderedDict, defaultdict
    sys.path.insert(0, REF)
    from utils import compute_temporal_iou_batch_cross, compute_temporal_iou_batch_paired, get_window_len, interpolated_precision_recall
    ns = dict(np=np, copy=copy, time=time, defaultdict=defaultdict, OrderedDict=OrderedDict, mp=None,
              compute_temporal_iou_batch_cross=compute_temporal_iou_batch_cross, compute_temporal_iou_batch_paired=compute_temporal_iou_batch_paired,
              get_window_len=get_window_len, interpolated_precision_recall=interpolated_precision_recall)
    path = os.path.join(REF, "eval.py")
    for fn in ("eval_moment_retrieval", "compute_mr_ap", "compute_average_precision_detection_wrapper", "compute_average_precision_detection",
               "compute_mr_r1", "get_data_by_range"):
        exec(_function_source(path, fn).replace("num_workers=8", "num_workers=1"), ns)
    out, expect = {}, {}
    rng = np.random.default_rng(5)
    for name, dataset, B, max_ts, ngt in (("charades", "charades", 300, 60.0, 1), ("qvh", "qvhighlights", 200, 150.0, 4), ("tacos", "tacos", 250, 600.0, 1)):
        nq = 10
        st = rng.uniform(0, max_ts * 0.8, (B, nq)).round(2)
        ed = np.minimum(st + rng.uniform(0.5, max_ts * 0.5, (B, nq)).round(2), max_ts)
        sc = -np.sort(-rng.uniform(0, 1, (B, nq)).round(4), axis=1)
        sc[::9, 3] = sc[::9, 2]                                      # equal scores
        win = np.stack([st, ed, sc], -1)
        n_gt = rng.integers(1, ngt + 1, B)
        gts, offs = [], [0]
        for b in range(B):
            g0 = rng.uniform(0, max_ts * 0.7, n_gt[b]).round(2)
            g1 = np.minimum(g0 + rng.choice([3.0, 8.0, 20.0, 45.0, 200.0], n_gt[b]) * rng.uniform(0.5, 1.5, n_gt[b]), max_ts).round(2)
            if b % 4 == 0:                                           # make the top-1 prediction overlap a ground-truth window
                win[b, 0, 0], win[b, 0, 1] = g0[0] + 0.5, g1[0]
            if b % 6 == 0 and nq > 2:
                win[b, 2, :2] = win[b, 0, :2]                        # duplicate prediction: only one may match (lock_gt)
            gts.append(np.stack([g0, g1], 1))
            offs.append(offs[-1] + n_gt[b])
        sub = [dict(qid=b, pred_relevant_windows=win[b].tolist()) for b in range(B)]
        gt = [dict(qid=b, relevant_windows=gts[b].tolist()) for b in range(B)]
        res = ns["eval_moment_retrieval"](sub, gt, verbose=False, dataset_name=dataset)
        out.update({f"met_{name}_windows": win, f"met_{name}_gt": np.concatenate(gts), f"met_{name}_gt_off": np.asarray(offs, dtype=np.int64)})
        expect[name] = dict(dataset_name=dataset, metrics=res)
        print(f"metrics {name}: full R1@0.5 {res['full']['MR-R1']['0.5']} mAP {res['full']['MR-mAP']['average']} ranges {sorted(res)}")
    with open(os.path.join(GOLD, "metrics_expected.json"), "w") as f:
        json.dump(expect, f, indent=1)
    return out


def clip_cases():
    """The reference's CLIPTextEncoder (model/text_encoder.py:240-354) on seeded weights / tokens: once as shipped (fp16 weights via
    convert_weights, fp16 activations) and once with its `dtype` property overridden to float32 (the same graph in fp32: what the
    fp16 run approximates).  The first 32 token rows (max_words_l) and the pooled output are stored."""
    import copy
    sys.path.insert(0, REF)
    from model.text_encoder import CLIPTextEncoder, convert_weights
    from oracle.weights import CLIP_TEXT_CFG, make_clip_state_dict, make_clip_tokens
    sd, text = make_clip_state_dict(3), make_clip_tokens(5, 4)
    ref16 = CLIPTextEncoder(**CLIP_TEXT_CFG).eval()
    ref16.load_state_dict(sd, strict=True)
    ref32 = type("CLIPTextEncoderF32", (CLIPTextEncoder,), {"dtype": property(lambda self: torch.float32)})(**CLIP_TEXT_CFG).eval()
    ref32.load_state_dict(sd, strict=True)
    convert_weights(ref16)
    o32, o16 = ref32(text), ref16(text)
    d = (o16["last_hidden_state"].float() - o32["last_hidden_state"]).abs().max() / o32["last_hidden_state"].abs().max()
    print(f"clip text tower: fp16 run vs fp32 run of the reference module: {float(d):.2e} relative")
    return dict(clip_hidden_f32=o32["last_hidden_state"][:, :32].numpy(), clip_pooled_f32=o32["pooler_output"].numpy(),
                clip_hidden_f16=o16["last_hidden_state"][:, :32].numpy(), clip_pooled_f16=o16["pooler_output"].numpy())


def main():
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "clip_text.npz"), **clip_cases())
    np.savez_compressed(os.path.join(GOLD, "metrics.npz"), **metrics_cases())
    np.savez_compressed(os.path.join(GOLD, "criterion_saliency.npz"), **criterion_cases())
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **frontend_cases())
    print("written to", GOLD)


if __name__ == "__main__":
    main()
