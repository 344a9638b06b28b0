"""Data-parallel sharding of the pair-scoring path: pairs are independent, video groups stay whole, there is
no collective on the scoring path and ONE all_gather of per-rank top-k span records at the end.

Host-side logic only (pure functions of (rank, world_size, num_clips)) + the torch.distributed plumbing.
"""
import torch


def shard_groups(num_clips, rank, world):
    """Contiguous block of video groups for ``rank``: groups [G*rank/W, G*(rank+1)/W).
    Returns (group_lo, group_hi, pair_lo, pair_hi)."""
    nc = [int(x) for x in (num_clips.tolist() if torch.is_tensor(num_clips) else num_clips)]
    G = len(nc)
    glo, ghi = G * rank // world, G * (rank + 1) // world
    plo = sum(nc[:glo])
    return glo, ghi, plo, plo + sum(nc[glo:ghi])


def shard_batch(batch, rank, world):
    """Slice a collated batch dict (video_feat, video_mask, words_feat, duration, num_clips, ...) to this rank's groups."""
    glo, ghi, plo, phi = shard_groups(batch["num_clips"], rank, world)
    out = {}
    for k, v in batch.items():
        if k == "num_clips":
            out[k] = v[glo:ghi]
        elif torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == int(sum(batch["num_clips"])):
            out[k] = v[plo:phi]
        else:
            out[k] = v
    out["pair_offset"] = plo
    return out


def local_topk(windows, order, k, pair_offset=0):
    """Per-rank top-k records by best-span score.  windows f64[B,nq,3] ranked, order i32[B,nq].
    Record = [global pair id, query idx, st, ed, score] (f64[k',5], k' = min(k,B)), sorted by (score desc, pair id asc)."""
    B = windows.shape[0]
    score = windows[:, 0, 2]
    kk = min(k, B)
    # deterministic tie-break on pair id: sort by (-score, pair)
    pair = torch.arange(B, device=windows.device, dtype=torch.float64) + pair_offset
    idx = torch.argsort(pair, stable=True)
    idx = idx[torch.argsort(-score[idx], stable=True)][:kk]
    return torch.stack([pair[idx], order[idx, 0].to(torch.float64), windows[idx, 0, 0], windows[idx, 0, 1], score[idx]], dim=1)


def merge_topk(records, k):
    """k-way merge of per-rank record blocks (concatenated f64[*,5]) -> global top-k, same ordering rule."""
    idx = torch.argsort(records[:, 0], stable=True)
    idx = idx[torch.argsort(-records[idx, 4], stable=True)][:k]
    return records[idx]


def gather_topk(windows, num_clips, k, rank, world, pair_stride, order=None):
    """One all_gather (NCCL on GPU tensors, gloo on CPU tensors) of the per-rank top-k records, merged on every rank.
    ``pair_stride``: pairs per rank (weak scaling) used to form global pair ids."""
    if order is None:
        order = torch.zeros(windows.shape[:2], dtype=torch.int32, device=windows.device)
    rec = local_topk(windows, order, k, pair_offset=rank * pair_stride)
    if world > 1:
        import torch.distributed as dist
        pad = torch.full((k, 5), float("-inf"), dtype=torch.float64, device=rec.device)
        pad[:rec.shape[0]] = rec
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        rec = torch.cat(bufs)
        rec = rec[rec[:, 4] > float("-inf")]
    return merge_topk(rec, k)
