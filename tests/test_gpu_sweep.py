"""GPU: the batch-sharded scoring sweep (BASELINE.json configs[4], sweep.py) — the running top-k over batches equals the
top-k of all scored pairs, and re-running is deterministic."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_sweep_running_topk_equals_global_topk():
    import mesm_b200
    import sweep
    from mesm_b200.model import build_model
    from mesm_b200.sharding import local_topk
    dev = torch.device("cuda", 0)
    cfg = dict(sweep.QVH)
    torch.manual_seed(0)
    model = build_model(cfg).to(dev)
    best, done = sweep.run_sweep(model, cfg, 300, 128, 10, 0, 1, dev)
    assert done == 300 and best.shape == (10, 5)
    wins, orders = [], []
    for bi, B in enumerate([128, 128, 44]):
        wl = sweep.make_batch(cfg, B, 4242 + bi, dev)
        out = model(wl["video_feat"], wl["video_mask"], wl["words_feat"], None, None, wl["num_clips"],
                    dataset_name="qvhighlights", is_training=False)
        w, o, _, _ = mesm_b200.decode_nms(out["pred_logits"], out["pred_spans"], wl["duration"], cfg["clip_len"],
                                          cfg["max_ts_val"], 0.7, 10, 10)
        wins.append(w)
        orders.append(o)
    ref = local_topk(torch.cat(wins), torch.cat(orders), 10)
    assert torch.equal(best, ref)
    best2, _ = sweep.run_sweep(model, cfg, 300, 128, 10, 0, 1, dev)
    assert torch.equal(best, best2)
