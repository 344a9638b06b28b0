"""Host -> device ingest of a collated batch: drop-in for ``prepare_batch_input`` (dataset/base.py:358-383, called at
eval.py:62 / train.py:60).

Same name, arguments and in-place semantics as the reference.  The one difference is how ``video_feat`` crosses PCIe:
the collate function zero-pads every video to the longest of the batch (utils/data_utils.py:66-82), so on ragged
batches a quarter or more of the padded fp32 tensor is zeros.  ``mesm_upload_clips`` (include/mesm_b200.h) copies only
the valid prefix rows of each pair and zero-fills the pad rows on the device; the resulting device tensor is
bit-identical to ``video_feat.to(device)``.  Everything else is a plain ``.to(device, non_blocking=...)``.  The clip
counts the upload derives from the host mask are kept in the batch as ``video_len`` (host int32[B]); ``MESM.forward``
picks them up from ``**kwargs`` and runs on packed variable-length rows.
"""
import collections
import ctypes
from ctypes import byref, c_int64

import torch

from . import _lib
from ._lib import check
from .engine import _ptr, _stream

_SKIP = ("words_weight",        # stays on the CPU in the reference as well (dataset/base.py:360-361)
         "video_len")           # host clip counts added by this function: the engine reads them on the host


# Host tensors whose bytes an enqueued (stream-ordered) copy still has to read: (event, tensors).  mesm_upload_clips issues
# raw cudaMemcpy(Batch)Async calls, which - unlike tensor.to(non_blocking=True) - tell PyTorch's caching host allocator
# nothing: a pinned source dropped by the caller right after the call would go back to the allocator's free list and could
# be overwritten (e.g. by the DataLoader's pin-memory thread) before the DMA has read it.  Every upload therefore parks its
# sources here until an event recorded behind the copies has completed.
_inflight = collections.deque()


def _reap_inflight():
    while _inflight and _inflight[0][0].query():
        _inflight.popleft()


def _hold_until_copied(tensors, non_blocking):
    """Keep ``tensors`` (host sources of copies just enqueued on the current stream) alive until the copies are done.
    Pageable sources, and ``non_blocking=False``, wait for the stream instead (the reference's blocking ``.to(device)``)."""
    stream = torch.cuda.current_stream()
    if not non_blocking or not all(t.is_pinned() for t in tensors):
        stream.synchronize()
        return
    ev = torch.cuda.Event()
    ev.record(stream)
    _inflight.append((ev, tensors))
    _reap_inflight()


def upload_clips(video_feat, video_mask, device=None, out_feat=None, out_mask=None, num_clips=None, non_blocking=True):
    """Ragged upload of host ``video_feat`` f32 (or f16: the 16-bit storage option) [B,L,Dv] / ``video_mask`` bool[B,L] on the current stream.
    Returns (dev_feat, dev_mask, bytes_copied).  Pin the host tensors for the copies to be asynchronous.
    ``num_clips`` (optional, charades / tacos batches only): the collate step replicates a group's video for each of
    its queries (dataset/base.py:307-309); only the first pair of every group is then uploaded and the rows of the
    other pairs are left untouched - to be consumed with ``shared_group_video=True``."""
    if video_feat.is_cuda or video_mask.is_cuda:
        raise RuntimeError("upload_clips takes host tensors")
    vf = video_feat.contiguous()
    f16 = vf.dtype == torch.float16                     # 16-bit feature storage: uploaded and consumed as is (half the PCIe bytes)
    if not f16 and vf.dtype != torch.float32:
        vf = vf.float()
    vm = video_mask.contiguous()
    vm8 = vm.view(torch.uint8) if vm.dtype == torch.bool else (vm != 0).view(torch.uint8)
    B, L, Dv = vf.shape
    if out_feat is None:
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        out_feat = torch.empty(B, L, Dv, dtype=vf.dtype, device=device)
    if out_mask is None:
        out_mask = torch.empty(B, L, dtype=torch.bool, device=out_feat.device)
    if tuple(out_feat.shape) != (B, L, Dv) or tuple(out_mask.shape) != (B, L) or not out_feat.is_contiguous() or out_feat.dtype != vf.dtype:
        raise RuntimeError("upload_clips: output buffers do not match the batch shape")
    n = c_int64(0)
    nc_arr, G = None, 0
    if num_clips is not None:
        nc = [int(x) for x in (num_clips.tolist() if torch.is_tensor(num_clips) else num_clips)]
        # sampled guard of the contract (not a proof: a collate that perturbs other elements of a replicated video is the
        # caller's responsibility): four clip rows per pair - first, last valid and two in between - on a
        # column stride that covers the whole feature width must equal the same elements of the group's first pair
        first = torch.repeat_interleave(torch.cumsum(torch.tensor([0] + nc[:-1]), 0), torch.tensor(nc))
        ok = sum(nc) == B and torch.equal(vm, vm[first])
        if ok:
            n_valid = clip_counts(vm).long()
            ar = torch.arange(B)
            cstep = max(1, Dv // 64)
            for frac in (0.0, 1.0 / 3, 2.0 / 3, 1.0):
                rows = ((n_valid - 1).double() * frac).long()
                if not torch.equal(vf[ar, rows, ::cstep], vf[first, rows, ::cstep]):
                    ok = False
                    break
        if not ok:
            raise ValueError("upload_clips(num_clips=...): the pairs of a video group do not share one video")
        nc_arr, G = (c_int64 * len(nc))(*nc), len(nc)
    with torch.cuda.device(out_feat.device):
        fn = _lib.lib().mesm_upload_clips_f16 if f16 else _lib.lib().mesm_upload_clips
        check(fn(ctypes.c_void_p(vf.data_ptr()), ctypes.c_void_p(vm8.data_ptr()), B, L, Dv,
                 _ptr(out_feat), _ptr(out_mask.view(torch.uint8)), nc_arr, G, byref(n), _stream()))
        # vf / vm8 may be temporaries (.contiguous(), .float(), != 0) and the caller may drop its own tensors right away
        _hold_until_copied((video_feat, video_mask, vf, vm8), non_blocking)
    return out_feat, out_mask, int(n.value)


def clip_counts(video_mask):
    """Host clip counts int32[B] of a host mask bool[B, L]: index of the last valid clip + 1 (>= 1), i.e. the `lengths`
    of the collate step (utils/data_utils.py:64) - what ``Engine.forward(video_len=...)`` expects."""
    m = video_mask.bool()
    L = m.shape[1]
    last = L - torch.flip(m, dims=[1]).to(torch.int8).argmax(dim=1)         # first True from the right
    last = torch.where(m.any(dim=1), last, torch.ones_like(last))
    return last.to(torch.int32)


def prepare_batch_input(batched_data, device, non_blocking=False, out=None, shared_group_video=False, relay=None):
    """dataset/base.py:358-383.  ``out`` (optional): dict of preallocated device tensors to copy into (double buffering);
    keys missing from it are allocated.  ``shared_group_video=True`` (charades / tacos batches, whose collate replicates
    the video of a group for each of its queries): every video crosses PCIe once; the batch gets
    ``shared_group_video=True`` so that ``MESM.forward`` reads the clips of a pair from its group's first pair.  ``relay``
    (optional, ``mesm_b200.relay.IngestRelay``; needs ``out`` and ``shared_group_video``): part of the videos reaches this GPU
    through a peer GPU's host link and NVLink.  ``prepare_batch_input.last_h2d_bytes`` = host->device bytes this call enqueued."""
    from .utils import span_xx_to_cxw
    device = torch.device(device)
    out = out or {}
    total = 0
    ragged = ("video_feat" in batched_data and "video_mask" in batched_data and torch.is_tensor(batched_data["video_feat"])
              and not batched_data["video_feat"].is_cuda and batched_data["video_feat"].dim() == 3
              and torch.is_tensor(batched_data["video_mask"]) and not batched_data["video_mask"].is_cuda)
    if ragged:
        host_mask = batched_data["video_mask"]
        shared = bool(shared_group_video) and "num_clips" in batched_data
        if relay is not None and shared and out.get("video_feat") is not None and out.get("video_mask") is not None:
            f, m, n = relay.upload(batched_data["video_feat"].contiguous(), batched_data["video_mask"].contiguous(), out["video_feat"],
                                   out["video_mask"], batched_data["num_clips"], non_blocking=non_blocking)
        else:
            f, m, n = upload_clips(batched_data["video_feat"], batched_data["video_mask"], device, out.get("video_feat"),
                                   out.get("video_mask"), batched_data["num_clips"] if shared else None, non_blocking=non_blocking)
        if shared:
            batched_data["shared_group_video"] = True
        batched_data["video_feat"], batched_data["video_mask"] = f, m
        total += n
    for key, value in batched_data.items():
        if key in _SKIP or (ragged and key in ("video_feat", "video_mask")):
            continue
        if isinstance(value, torch.Tensor):
            if not value.is_cuda:
                total += value.numel() * value.element_size()
            if key in out:
                out[key].copy_(value, non_blocking=non_blocking)
                batched_data[key] = out[key]
            else:
                batched_data[key] = value.to(device, non_blocking=non_blocking)
        if key == "norm_moment":
            batched_data[key] = [dict(moments=e["moments"].to(device, non_blocking=non_blocking)) for e in value]
        if key == "norm_span":
            batched_data[key] = [dict(spans=e["spans"].to(device, non_blocking=non_blocking)) for e in value]
    if ragged and "video_len" not in batched_data:
        batched_data["video_len"] = clip_counts(host_mask)        # stays on the host; MESM.forward reads it from **kwargs
    if "moment" in batched_data and "norm_span" not in batched_data:
        moment, duration = batched_data["moment"], batched_data["duration"]
        batched_data["norm_moment"] = moment / duration.unsqueeze(1)
        batched_data["norm_span"] = span_xx_to_cxw(batched_data["norm_moment"])
    prepare_batch_input.last_h2d_bytes = total
    return batched_data


def build_video_feat(raw_sources, max_video_l, normalize_video=True, use_tef=True, out_dtype=torch.float32, device=None):
    """The dataset-side front-end of ONE video on the device: get_video_feat (dataset/charades.py:108-119) + sample_video_feat
    (dataset/base.py:100-114) + add_tef (dataset/base.py:225-230).  ``raw_sources``: list of raw per-source clip features
    [L_s, D_s], all fp32 or all fp16 (host or device).  Returns [L, sum(D_s) + 2] in ``out_dtype`` (torch.float16 = the 16-bit
    storage option ``MESM.forward`` consumes directly).  For qvhighlights pass the raw arrays truncated to ``max_video_l`` rows
    (dataset/qvhighlights.py:205)."""
    from ctypes import c_int32, c_void_p
    lib = _lib.lib()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    f16 = raw_sources[0].dtype == torch.float16
    srcs = [r.to(device=device, dtype=torch.float16 if f16 else torch.float32).contiguous() for r in raw_sources]
    S = len(srcs)
    lens = (c_int32 * S)(*[int(r.shape[0]) for r in srcs])
    dims = (c_int32 * S)(*[int(r.shape[1]) for r in srcs])
    ptrs = (c_void_p * S)(*[r.data_ptr() for r in srcs])
    L = int(lib.mesm_video_feat_rows(lens, S, int(max_video_l)))
    W = sum(int(r.shape[1]) for r in srcs) + (2 if use_tef else 0)
    if out_dtype not in (torch.float32, torch.float16):
        raise ValueError("out_dtype must be torch.float32 or torch.float16")
    out = torch.empty(L, W, dtype=out_dtype, device=device)
    ws = torch.empty(int(lib.mesm_video_feat_workspace_bytes(lens, S, int(max_video_l))), dtype=torch.uint8, device=device)
    with torch.cuda.device(device):
        check(lib.mesm_build_video_feat(ptrs, lens, dims, S, int(f16), int(bool(normalize_video)), int(max_video_l), int(bool(use_tef)),
                                        _ptr(out), int(out_dtype == torch.float16), _ptr(ws), ws.numel(), _stream()))
    return out
