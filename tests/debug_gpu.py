"""Manual GPU debugging aid: per-tap errors of the CUDA forward vs the oracle (python -m tests.debug_gpu [case])."""
import sys

import torch

from oracle import mesm_oracle
from tests.helpers import engine_cfg, load_case, rel_err


def main(name):
    import mesm_b200
    cfg, sd, inp, neg, gold, meta = load_case(name)
    oo = mesm_oracle.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
    eng = mesm_b200.Engine(engine_cfg(cfg))
    eng.load_state_dict(sd)
    dev = eng.device
    out = eng.forward(inp["video_feat"].to(dev), inp["video_mask"].to(dev), inp["words_feat"].to(dev), inp["num_clips"],
                      neg_index=neg.to(dev), want=("core", "aux", "rec", "taps"))
    torch.cuda.synchronize()
    vm = inp["video_mask"]
    t = oo["_taps"]
    rows = [("projed_video_feat", out["projed_video_feat"], oo["projed_video_feat"], None),
            ("projed_words_feat", out["expanded_words_feat"][:, 1:], oo["projed_words_feat"], None),
            ("enhanced_video_feat", out["enhanced_video_feat"], oo["enhanced_video_feat"], vm[..., None]),
            ("recon_feat", out["recon_feat"], oo["recon_feat"], None),
            ("projed_recon_feat", out["projed_recon_feat"], oo["projed_recon_feat"], None),
            ("memory", out["memory"], t["memory"], vm[..., None]),
            ("memory_global", out["memory_global"], t["memory_global"], None),
            ("hs", out["hs"], t["hs"], None),
            ("pred_logits", out["pred_logits"], oo["pred_logits"], None),
            ("pred_spans", out["pred_spans"], oo["pred_spans"], None),
            ("saliency", out["saliency_scores"], oo["saliency_scores"], vm),
            ("neg_saliency", out["neg_saliency_scores"], oo["neg_saliency_scores"], vm)]
    for n, a, b, m in rows:
        print(f"{name:22s} {n:22s} rel={rel_err(a, b, m):.3e} nan={bool(torch.isnan(a).any())}")
    print("launches", eng.last_launch_count)


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["tiny_uniform", "tiny_ragged", "tiny_qvh_groups", "tiny_twomlp", "qvh_b6"]):
        main(n)
