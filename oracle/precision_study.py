"""Which tensor-core operand format keeps the forward within the 1e-3 parity bar?  (SURVEY §7 "hard parts":
"decide with the oracle taps, not by guess".)  TEST INFRASTRUCTURE.

    python -m oracle.precision_study            # blanket modes + the per-GEMM mixes (table in DESIGN.md section 2)

Blanket: every product in one operand format.  Mix: bf16x3 everywhere except the named sites (``ffn1`` = W1.X of the
fused FFN blocks, ``ffn2`` = W2.H, ``attn`` = QK^T / PV), which run in the cheaper format.  Cost unit = one bf16/fp16
tcgen05.mma pass (kind::tf32 runs at half rate = 2 units)."""
import sys

import torch
from oracle.config import CONFIGS
from oracle import mesm_oracle as mo
from oracle.weights import make_state_dict, make_inputs, make_neg_index


def rel(a, b, mask=None):
    d = (a - b).abs()
    if mask is not None:
        d = d * mask
        b = b * mask
    return float(d.max() / b.abs().max())


CASES = (("qvhighlights", [1] * 16, dict(ragged_video=False)),
         ("charades_csf", [2, 3, 1, 2, 4, 4], dict()),
         ("tacos", [4, 3, 5], dict(lv=96)))

BLANKET = ("bf16", "tf32_trunc", "tf32_rn", "fp16", "bf16x2", "fp16x2a", "fp16x2w", "bf16x3", "fp16x3")
MIXES = (
    ("ffn1+ffn2 fp16", dict(ffn1="fp16", ffn2="fp16")),
    ("ffn1+ffn2 tf32_rn", dict(ffn1="tf32_rn", ffn2="tf32_rn")),
    ("ffn1+ffn2 fp16x2a", dict(ffn1="fp16x2a", ffn2="fp16x2a")),
    ("ffn1+ffn2 fp16x2w", dict(ffn1="fp16x2w", ffn2="fp16x2w")),
    ("ffn1+ffn2 bf16x2", dict(ffn1="bf16x2", ffn2="bf16x2")),
    ("ffn1 fp16", dict(ffn1="fp16")),
    ("ffn2 fp16", dict(ffn2="fp16")),
    ("ffn1 fp16x2a, ffn2 fp16", dict(ffn1="fp16x2a", ffn2="fp16")),
    ("ffn1 fp16, ffn2 fp16x2a", dict(ffn1="fp16", ffn2="fp16x2a")),
    ("attn fp16", dict(attn="fp16")),
    ("attn fp16x2a", dict(attn="fp16x2a")),
    ("ffn fp16x2a + attn fp16", dict(ffn1="fp16x2a", ffn2="fp16x2a", attn="fp16")),
)


def run(cfg_name, nc, kw, seeds=(1,)):
    cfg = CONFIGS[cfg_name]
    rows = {}
    for seed in seeds:
        sd = make_state_dict(cfg, seed)
        inp = make_inputs(cfg, nc, seed + 1, **kw)
        neg = make_neg_index(nc, seed + 2)
        args = (sd, cfg, inp["video_feat"], inp["video_mask"], inp["words_feat"], inp["num_clips"], neg)
        mo.set_matmul_precision("fp32")
        ref = mo.mesm_forward(*args)
        vm = inp["video_mask"]

        def measure(label):
            o = mo.mesm_forward(*args)
            r = (rel(o["pred_logits"], ref["pred_logits"]), rel(o["pred_spans"], ref["pred_spans"]),
                 rel(o["saliency_scores"], ref["saliency_scores"], vm),
                 rel(o["neg_saliency_scores"], ref["neg_saliency_scores"], vm))
            rows[label] = tuple(max(a, b) for a, b in zip(rows.get(label, (0, 0, 0, 0)), r))

        for mode in BLANKET:
            mo.set_matmul_precision(mode)
            measure("blanket " + mode)
        for label, sites in MIXES:
            mo.set_matmul_precision("bf16x3", sites)
            measure("bf16x3 + " + label)
        mo.set_matmul_precision("fp32")
    for label, r in rows.items():
        print(f"{cfg_name:13s} {label:42s} logits {r[0]:.1e} spans {r[1]:.1e} sal {r[2]:.1e} neg_sal {r[3]:.1e}", flush=True)


def main():
    seeds = (1, 11, 21) if "--seeds3" in sys.argv else (1,)
    for c in CASES:
        run(*c, seeds=seeds)


if __name__ == "__main__":
    main()
