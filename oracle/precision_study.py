"""Which tensor-core operand format keeps the forward within the 1e-3 parity bar?  (SURVEY §7 "hard parts":
"decide with the oracle taps, not by guess".)  TEST INFRASTRUCTURE.  python -m oracle.precision_study"""
import torch
from oracle.config import CONFIGS
from oracle import mesm_oracle as mo
from oracle.weights import make_state_dict, make_inputs, make_neg_index


def rel(a, b, mask=None):
    d = (a - b).abs()
    if mask is not None:
        d = d * mask
        b = b * mask
    return float(d.max() / b.abs().max()), float(d.max())


def main():
    for cfg_name, nc, kw in (("qvhighlights", [1] * 16, dict(ragged_video=False)),
                             ("charades_csf", [2, 3, 1, 2, 4, 4], dict())):
        cfg = CONFIGS[cfg_name]
        sd = make_state_dict(cfg, 1)
        inp = make_inputs(cfg, nc, 2, **kw)
        neg = make_neg_index(nc, 3)
        mo.set_matmul_precision("fp32")
        ref = mo.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
        for mode in ("bf16", "tf32_trunc", "tf32_rn", "bf16x2", "bf16x3"):
            mo.set_matmul_precision(mode)
            o = mo.mesm_forward(sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
            vm = inp["video_mask"]
            print(f"{cfg_name:14s} {mode:10s} logits rel/abs {rel(o['pred_logits'], ref['pred_logits'])} "
                  f"spans {rel(o['pred_spans'], ref['pred_spans'])} sal {rel(o['saliency_scores'], ref['saliency_scores'], vm)} "
                  f"projv {rel(o['projed_video_feat'], ref['projed_video_feat'])[0]:.1e}")
        mo.set_matmul_precision("fp32")


if __name__ == "__main__":
    main()
