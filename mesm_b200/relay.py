"""NVLink-assisted host -> device ingest for multi-GPU nodes whose GPUs do not get equal host bandwidth.

On the 8 x B200 boxes this was measured on (profiles/r2_h2d_probe_n8.md) four GPUs receive 22.7 GB/s from the host and the
other four 35.5 GB/s when all eight ranks upload at once; one step of the benchmark moves 1.4 GB per GPU, so the slow ranks
are copy-bound (65 ms against 55 ms of kernels) and, under max-over-ranks timing, so is the node.  A rank on a slow link can
send part of its rows over a FAST peer's link instead: pinned host -> staging buffer on the peer GPU (that GPU's PCIe link),
then peer GPU -> own GPU over NVLink.  Every byte still crosses PCIe exactly once.  Both hops are plain copies issued by
THIS process on streams of the peer device (no kernel is launched there, so the peer rank's compute is not time-sliced), no
inter-process communication is involved, and the result in the destination tensor is bit-identical to a direct upload.

``plan_ingest_relay`` probes the links (all ranks at once), pairs the slowest ranks with the fastest and returns an
``IngestRelay`` for the ranks that should offload (None elsewhere, and everywhere when the links are symmetric).
``prepare_batch_input(..., relay=...)`` / ``IngestRelay.upload`` are drop-ins for the direct ragged upload.
"""
import ctypes

import numpy as np
import torch

from . import _lib, ingest
from ._lib import check
from .engine import _stream


class IngestRelay:
    def __init__(self, device, relay_device, fraction, staging_bytes, nbuf=3):
        self.device, self.relay_device = torch.device(device), torch.device(relay_device)
        if self.device == self.relay_device:
            raise ValueError("the relay device must be another GPU")
        self.fraction = float(fraction)
        self.h2d_stream = torch.cuda.Stream(device=self.relay_device)            # host -> peer GPU, then peer GPU -> own GPU
        self.dst_stream = torch.cuda.Stream(device=self.device, priority=-1)     # own-GPU side of the peer copies
        self.staging = [torch.empty(int(staging_bytes), dtype=torch.uint8, device=self.relay_device) for _ in range(nbuf)]
        self.landing = [torch.empty(int(staging_bytes), dtype=torch.uint8, device=self.device) for _ in range(nbuf)]
        self.free = [torch.cuda.Event() for _ in range(nbuf)]
        for e in self.free:
            e.record(self.dst_stream)
        self.k = 0
        self.last_relayed_bytes = 0
        self.broken = None           # set to the error text when the relay path failed once (direct uploads from then on)
        self.batched = True          # per-video staging copies in one cudaMemcpyBatchAsync submission (False: one cudaMemcpyAsync each)

    def describe(self):
        return {"relay_device": str(self.relay_device), "fraction": round(self.fraction, 3), "broken": self.broken}

    def upload(self, video_feat, video_mask, out_feat, out_mask, num_clips, non_blocking=True):
        """Relayed upload (see ``_upload``); any host-side failure of the relay path (allocation, peer access, a missing symbol)
        disables the relay for the rest of the run and falls back to the direct upload of the whole batch."""
        if not self.broken:
            try:
                return self._upload(video_feat, video_mask, out_feat, out_mask, num_clips, non_blocking)
            except Exception as e:                              # noqa: BLE001 - the direct path always works
                self.broken = f"{type(e).__name__}: {e}"
        self.last_relayed_bytes = 0
        return ingest.upload_clips(video_feat, video_mask, out_feat=out_feat, out_mask=out_mask, num_clips=num_clips, non_blocking=non_blocking)

    def _upload(self, video_feat, video_mask, out_feat, out_mask, num_clips, non_blocking=True):
        """``ingest.upload_clips(video_feat, video_mask, out_feat=..., out_mask=..., num_clips=...)`` on the current stream of
        ``out_feat``'s device, with the videos of the LAST groups of the batch (about ``fraction`` of the valid clip rows)
        routed through the relay GPU.  Returns (out_feat, out_mask, bytes crossing PCIe)."""
        nc = [int(x) for x in (num_clips.tolist() if torch.is_tensor(num_clips) else num_clips)]
        B, L, Dv = video_feat.shape
        n_valid = ingest.clip_counts(video_mask).tolist()
        first, b = [], 0
        for c in nc:
            first.append(b)
            b += c
        if b != B:
            raise ValueError("IngestRelay.upload: num_clips does not cover the batch")
        rows = [n_valid[f] for f in first]
        total = sum(rows)
        esz = video_feat.element_size()
        stg = self.staging[self.k % len(self.staging)]
        # relayed groups: from the end of the batch, up to `fraction` of the rows and the staging capacity
        budget_rows = min(int(total * self.fraction), stg.numel() // (Dv * esz))
        Gd, acc = len(nc), 0
        while Gd > 1 and acc + rows[Gd - 1] <= budget_rows:
            Gd -= 1
            acc += rows[Gd]
        if Gd == len(nc):                                   # nothing fits: plain direct upload
            self.last_relayed_bytes = 0
            return ingest.upload_clips(video_feat, video_mask, out_feat=out_feat, out_mask=out_mask, num_clips=nc, non_blocking=non_blocking)
        Bd = first[Gd]
        cur = torch.cuda.current_stream(self.device)
        slot = self.k % len(self.staging)
        # mask and zero fill of the relayed pairs first (their pad rows must be zero; the scatter below writes the valid rows) ...
        with torch.cuda.stream(cur):
            vm = video_mask[Bd:]
            out_mask[Bd:].copy_(vm if vm.dtype == torch.bool else vm != 0, non_blocking=non_blocking)
            out_feat[Bd:].zero_()
            zeroed = torch.cuda.Event()
            zeroed.record(cur)
        # ... then the direct part on the caller's stream (uploads, zero fill, mask, the shared-video guard)
        _, _, n_direct = ingest.upload_clips(video_feat[:Bd], video_mask[:Bd], out_feat=out_feat[:Bd], out_mask=out_mask[:Bd],
                                             num_clips=nc[:Gd], non_blocking=non_blocking)
        # relayed part: valid rows of the groups' first pairs, compacted: host -> staging on the peer GPU (its PCIe link), ONE peer
        # copy staging -> landing buffer on this GPU (NVLink), one scatter kernel landing -> rows of the padded tensor
        nrows = acc
        stg2 = stg[:nrows * Dv * esz].view(video_feat.dtype).view(nrows, Dv)
        land2 = self.landing[slot][:nrows * Dv * esz].view(video_feat.dtype).view(nrows, Dv)
        # one batched submission of the per-video host -> staging copies, destination rows of the scatter (both built vectorised: the
        # thread that runs this also enqueues the forward passes)
        gi = [g for g in range(Gd, len(nc)) if rows[g] > 0]
        n_g = torch.tensor([rows[g] for g in gi], dtype=torch.int64)
        f_g = torch.tensor([first[g] for g in gi], dtype=torch.int64)
        start_g = torch.cumsum(n_g, 0) - n_g
        rowb = Dv * esz
        src_p = (video_feat.data_ptr() + f_g * (L * rowb)).numpy().astype(np.uint64)
        dst_p = (stg.data_ptr() + start_g * rowb).numpy().astype(np.uint64)
        size_p = (n_g * rowb).numpy().astype(np.uint64)
        idx = (torch.repeat_interleave(f_g * L - start_g, n_g) + torch.arange(nrows)).pin_memory()
        self.h2d_stream.wait_event(self.free[slot])                          # staging / landing of this slot have been consumed
        with torch.cuda.stream(self.h2d_stream):                             # (makes the relay device current for these copies)
            if self.batched:
                check(_lib.lib().mesm_memcpy_batch_h2d(src_p.ctypes.data_as(ctypes.c_void_p), dst_p.ctypes.data_as(ctypes.c_void_p),
                                                       size_p.ctypes.data_as(ctypes.c_void_p), len(gi), _stream()))
            else:
                for g, s0, n in zip(gi, start_g.tolist(), n_g.tolist()):
                    stg2[s0:s0 + n].copy_(video_feat[first[g], :n], non_blocking=True)
            with torch.cuda.stream(self.dst_stream):
                self.dst_stream.wait_event(zeroed)
                idx_d = idx.to(self.device, non_blocking=True)
                # torch runs a cross-device copy on the SOURCE device's current stream (h2d_stream: behind the uploads above) and makes
                # the destination device's current stream (dst_stream) wait for it
                land2.copy_(stg2, non_blocking=True)
                out_feat.view(B * L, Dv).index_copy_(0, idx_d, land2)
                done = torch.cuda.Event()
                done.record(self.dst_stream)
                self.free[slot].record(self.dst_stream)
            ev = torch.cuda.Event()
            ev.record(self.h2d_stream)
            ingest._inflight.append((ev, (video_feat, video_mask)))          # the pinned sources stay alive until the DMA has read them
        ingest._inflight.append((done, (idx, idx_d)))
        cur.wait_event(done)
        if not non_blocking:
            cur.synchronize()
        self.k += 1
        relayed = nrows * Dv * esz
        self.last_relayed_bytes = relayed
        return out_feat, out_mask, n_direct + relayed + (B - Bd) * L


def probe_h2d_gbs(device, nbytes=256 << 20, reps=4):
    """Host -> device bandwidth of this rank's link right now (call it on every rank at the same time)."""
    dev = torch.device(device)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(s):
        d.copy_(h, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            d.copy_(h, non_blocking=True)
        e1.record(s)
    e1.synchronize()
    return reps * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def plan_ingest_relay(local_rank, dist, staging_bytes, nbuf=3, min_ratio=1.15, max_fraction=0.3, force_fraction=None):
    """Single node, one rank per GPU.  Returns (IngestRelay or None, info dict).  The slowest rank is paired with the fastest, the
    second slowest with the second fastest, ...; a pair is used when the fast link has at least ``min_ratio`` times the
    bandwidth of the slow one, and the slow rank then offloads the share of its rows that equalises the two links' load."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", local_rank)
    dist.barrier()
    bw = probe_h2d_gbs(dev)
    t = torch.zeros(world, 2, device=dev)
    t[rank, 0], t[rank, 1] = bw, float(local_rank)
    dist.all_reduce(t)
    bws, locs = t[:, 0].tolist(), [int(x) for x in t[:, 1].tolist()]
    info = {"h2d_gbs": [round(x, 1) for x in bws], "pairs": pair_links(bws, min_ratio, max_fraction, force_fraction)}
    relay, err = None, None
    try:
        for pr in info["pairs"]:
            if rank == pr["rank"] and pr["fraction"] > 0:
                relay = IngestRelay(dev, torch.device("cuda", locs[pr["via_rank"]]), pr["fraction"], staging_bytes, nbuf)
    except Exception as e:                                      # noqa: BLE001 - e.g. no memory for the staging buffers on the peer
        relay, err = None, f"{type(e).__name__}: {e}"
    # the decision has to be the same on every rank (the caller's trial runs collectives): one failure disables the relay for all
    ok = torch.tensor([0.0 if err else 1.0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if float(ok) < 1.0:
        info["disabled"] = err or "the relay could not be set up on another rank"
        info["pairs"] = []
        relay = None
    return relay, info


def pair_links(bws, min_ratio=1.15, max_fraction=0.3, force_fraction=None):
    """Pairs of (slow rank, fast rank, share of the slow rank's rows to send through the fast rank's link) from the per-rank
    host -> device bandwidths: slowest with fastest, second slowest with second fastest, ...  With loads equal before the
    split, x = (f - s) / (f + s) makes both links finish together ((1 - x) / s = (1 + x) / f)."""
    world = len(bws)
    order = sorted(range(world), key=lambda r: (bws[r], r))
    pairs = []
    for i in range(world // 2):
        slow, fast = order[i], order[world - 1 - i]
        if force_fraction is None and bws[fast] < min_ratio * max(bws[slow], 1e-9):
            continue
        f = force_fraction if force_fraction is not None else min(max_fraction, (bws[fast] - bws[slow]) / (bws[fast] + bws[slow]))
        pairs.append({"rank": slow, "via_rank": fast, "fraction": round(float(f), 3)})
    return pairs
