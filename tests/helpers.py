"""Shared helpers for the parity tests (oracle side = checker only)."""
import json
import os

import numpy as np
import torch

from oracle.config import CONFIGS
from oracle.weights import make_inputs, make_neg_index, make_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    with open(os.path.join(GOLD, "cases.json")) as f:
        return json.load(f)


def load_case(name):
    meta = golden_cases()[name]
    cfg = CONFIGS[meta["config"]]
    sd = make_state_dict(cfg, seed=meta["weight_seed"])
    inp = make_inputs(cfg, meta["num_clips"], seed=meta["input_seed"], **meta["kwargs"])
    neg = make_neg_index(meta["num_clips"], seed=meta["neg_seed"])
    gold = dict(np.load(os.path.join(GOLD, f"{name}.npz")))
    return cfg, sd, inp, neg, gold, meta


def engine_cfg(cfg):
    """OracleConfig -> the reference's JSON keys that mesm_b200.Engine reads."""
    d = cfg.asdict()
    d["share_MLP"] = d.pop("share_mlp")
    return d


def rel_err(a, b, mask=None):
    """max |a-b| / max |b|  (the parity metric of DESIGN.md; b = reference)."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    d = (a - b).abs()
    if mask is not None:
        m = torch.as_tensor(mask).cpu()
        d = d * m
        b = b * m
    return float(d.max() / b.abs().max().clamp_min(1e-30))
