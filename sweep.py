#!/usr/bin/env python
"""sweep.py — BASELINE.json configs[4]: synthetic video-query pair scoring sweep at the QVHighlights shape, batch-sharded
across the ranks of one node with ONE NCCL all_gather of the per-rank top-k span records at the end.

    python sweep.py --pairs 1000000                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        sweep.py --pairs 1000000                                      8 GPUs, 125 k pairs per rank

Every rank generates its shard on the device in batches (seed = base + rank * 100003 + batch), scores each batch through the
product path (MESM forward + decode / post-processing / NMS), folds the batch's best spans into a running top-k, and the
ranks exchange k records each exactly once.  There is no collective on the scoring path.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QVH = dict(dataset_name="qvhighlights", v_feat_dim=2818, t_feat_dim=512, hidden_dim=256, nheads=8, dim_feedforward=1024,
           num_queries=10, num_recfw_layers=2, t2v_layers=2, enc_layers=2, dec_layers=2, num_recss_layers=4, n_input_proj=2,
           rec_fw=True, rec_ss=True, share_MLP=True, max_words_l=32, max_video_l=75, aux_loss=True, vocab_size=1111,
           clip_len=2.0, max_ts_val=150.0)


def make_batch(cfg, B, seed, device):
    """QVHighlights-shaped synthetic batch: one query per video, all 75 clips valid, per-source
    L2-normalised features + tef columns, words N(0,1) with lengths U{3..32}, duration 150 s."""
    g = torch.Generator(device=device).manual_seed(seed)
    Lv, Lt, Dv, Dt = cfg["max_video_l"], cfg["max_words_l"], cfg["v_feat_dim"], cfg["t_feat_dim"]
    x = torch.randn(B, Lv, Dv - 2, device=device, generator=g)
    x[..., :512] = torch.nn.functional.normalize(x[..., :512], dim=-1)
    x[..., 512:] = torch.nn.functional.normalize(x[..., 512:], dim=-1)
    ar = torch.arange(Lv, device=device, dtype=torch.float32)
    tef = torch.stack([ar / Lv, (ar + 1) / Lv], dim=-1)[None].expand(B, Lv, 2)
    video = torch.cat([x, tef], dim=-1).contiguous()
    wl = torch.randint(3, Lt + 1, (B,), device=device, generator=g)
    words = torch.randn(B, Lt, Dt, device=device, generator=g)
    words = words * (torch.arange(Lt, device=device)[None] < wl[:, None])[..., None]
    return dict(video_feat=video, video_mask=torch.ones(B, Lv, dtype=torch.bool, device=device), words_feat=words,
                num_clips=torch.ones(B, dtype=torch.int64), duration=torch.full((B,), 150.0, device=device))


def run_sweep(model, cfg, pairs_rank, batch, topk, rank, world, device, base_seed=4242, nms_thd=0.7):
    """Scores ``pairs_rank`` pairs on this rank; returns (global top-k records f64[k,5], pairs scored)."""
    import mesm_b200
    from mesm_b200.sharding import local_topk, merge_topk
    best = torch.empty(0, 5, dtype=torch.float64, device=device)
    done = 0
    bi = 0
    while done < pairs_rank:
        B = min(batch, pairs_rank - done)
        wl = make_batch(cfg, B, base_seed + rank * 100003 + bi, device)
        out = model(wl["video_feat"], wl["video_mask"], wl["words_feat"], None, None, wl["num_clips"],
                    dataset_name=cfg["dataset_name"], is_training=False)
        win, order, keep, cnt = mesm_b200.decode_nms(out["pred_logits"], out["pred_spans"], wl["duration"], cfg["clip_len"],
                                                     cfg["max_ts_val"], nms_thd, 10, 10)
        rec = local_topk(win, order, topk, pair_offset=rank * pairs_rank + done)
        best = merge_topk(torch.cat([best, rec]), topk)
        done += B
        bi += 1
    if world > 1:
        import torch.distributed as dist
        pad = torch.full((topk, 5), float("-inf"), dtype=torch.float64, device=device)
        pad[:best.shape[0]] = best
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)                                   # the one collective of the path
        allr = torch.cat(bufs)
        best = merge_topk(allr[allr[:, 4] > float("-inf")], topk)
    return best, done


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=1_000_000, help="total pairs over all ranks")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--topk", type=int, default=100)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from mesm_b200.model import build_model
    torch.manual_seed(0)
    model = build_model(dict(QVH)).to(dev)
    pairs_rank = args.pairs // world
    run_sweep(model, QVH, min(pairs_rank, 2 * args.batch), args.batch, args.topk, rank, world, dev)      # warm-up
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    best, done = run_sweep(model, QVH, pairs_rank, args.batch, args.topk, rank, world, dev)
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "video-query pairs/sec", "workload": "qvh_1m_pair_sweep", "value": world * done / (float(t) / 1e3),
                          "unit": "pairs/s", "n_gpus": world, "pairs_total": world * done, "seconds": float(t) / 1e3,
                          "includes": "on-device synthetic generation + forward + decode/NMS + running top-k + one all_gather",
                          "topk": int(best.shape[0]), "best": best[0].tolist()}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
