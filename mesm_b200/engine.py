"""Thin Python owner of a ``mesm_ctx`` (include/mesm_b200.h): weights in, forward / decode out.

PyTorch is used for device memory, streams and the dtype/shape checks only; every arithmetic step runs inside
libmesm_b200.so.  There is no fallback: a missing library or a non-CUDA tensor raises.
"""
import ctypes
import os
import time
from ctypes import byref, c_int64

import torch

from . import _lib
from ._lib import MesmCfg, MesmDecodeParams, MesmInputs, MesmOutputs, check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"mesm_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _u8(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"mesm_b200: `{name}` must be a CUDA tensor (no CPU fallback)")
    if t.dtype == torch.bool:
        t = t.contiguous().view(torch.uint8)
    elif t.dtype != torch.uint8:
        t = (t != 0).contiguous().view(torch.uint8)
    return t.contiguous()


class Engine:
    """One ``mesm_ctx`` on one device.

    cfg: dict with the reference's JSON/argparse keys (utils/config.py:26-163): v_feat_dim (incl. tef), t_feat_dim,
    hidden_dim, nheads, dim_feedforward, num_queries, num_recfw_layers, t2v_layers, enc_layers, dec_layers,
    num_recss_layers, n_input_proj, rec_fw, rec_ss, share_MLP, dataset_name, max_words_l, max_video_l.
    """

    def __init__(self, cfg: dict, device=None, chunk_pairs: int = 0):
        self.lib = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("mesm_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        c = MesmCfg()
        c.v_feat_dim = int(cfg["v_feat_dim"]); c.t_feat_dim = int(cfg["t_feat_dim"])
        c.hidden_dim = int(cfg.get("hidden_dim", 256)); c.nheads = int(cfg.get("nheads", 8))
        c.dim_feedforward = int(cfg.get("dim_feedforward", 1024)); c.num_queries = int(cfg.get("num_queries", 10))
        c.num_recfw_layers = int(cfg.get("num_recfw_layers", 2)); c.t2v_layers = int(cfg.get("t2v_layers", 2))
        c.enc_layers = int(cfg.get("enc_layers", 2)); c.dec_layers = int(cfg.get("dec_layers", 2))
        c.num_recss_layers = int(cfg.get("num_recss_layers", 4)); c.n_input_proj = int(cfg.get("n_input_proj", 2))
        c.rec_fw = int(bool(cfg.get("rec_fw", True))); c.rec_ss = int(bool(cfg.get("rec_ss", True)))
        c.share_mlp = int(bool(cfg.get("share_MLP", cfg.get("share_mlp", True))))
        c.qvh_grouping = int(cfg.get("dataset_name", "charades") == "qvhighlights")
        c.max_words_l = int(cfg.get("max_words_l", 32)); c.max_video_l = int(cfg.get("max_video_l", 75))
        self.cfg = c
        with torch.cuda.device(self.device):
            self.ctx = self.lib.mesm_create(byref(c), self.device.index or 0)
        if not self.ctx:
            raise RuntimeError("mesm_create failed: " + self.lib.mesm_last_error(None).decode())
        # Pairs per internal chunk.  Large chunks keep the GEMM grids many waves deep (the tiles of a wave drift out of
        # phase, so one CTA's epilogue overlaps its neighbour's K loop); 0 = auto: ~900 k padded clip rows per chunk (~25 GB of
        # workspace at d = 256), i.e. the whole 4096-pair Charades batch.
        self.chunk_pairs = int(chunk_pairs) if chunk_pairs else max(64, 900_000 // max(1, int(c.max_video_l)))
        self.lib.mesm_set_chunk_pairs(self.ctx, self.chunk_pairs)
        self._ws = None
        self._keep = []          # tensors that must outlive the asynchronous call that uses them

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.mesm_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------------------------
    def load_state_dict(self, sd):
        """sd: mapping reference-state_dict-key -> tensor (any device).  Replaces model.load_state_dict (eval.py:513-522)."""
        with torch.cuda.device(self.device):
            for k, v in sd.items():
                if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                    continue
                t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
                shape = (c_int64 * max(t.dim(), 1))(*t.shape)
                check(self.lib.mesm_load_weight(self.ctx, k.encode(), _ptr(t), shape, t.dim(), 1, _stream()), self.ctx)
                self._keep.append(t)
            check(self.lib.mesm_finalize_weights(self.ctx, _stream()), self.ctx)   # synchronises the stream
            self._keep.clear()

    def set_chunk_pairs(self, n):
        self.chunk_pairs = int(n)
        self.lib.mesm_set_chunk_pairs(self.ctx, self.chunk_pairs)
        self._ws = None

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- forward ---------------------------------------------------------------------------------------------------
    def forward(self, video_feat, video_mask, words_feat, num_clips, neg_index=None, want=("core",), video_len=None,
                shared_group_video=False):
        """MESM.forward in eval mode (model/model.py:154-359).  ``want``: any of "core" (logits, spans, saliency),
        "aux", "rec" (the rec_ss extras of model.py:342-351), "taps" (memory, memory_global, hs).
        ``video_len``: optional HOST clip counts [B] (list / CPU tensor; ``video_mask[b, i]`` is False for
        ``i >= video_len[b]``): the forward then runs on packed variable-length rows and does no work on the padding.
        ``shared_group_video``: the clips of a pair are read from the first pair of its video group (what
        ``prepare_batch_input(..., shared_group_video=True)`` uploads for charades / tacos batches); needs ``video_len``."""
        t_enter = time.perf_counter()
        f16 = video_feat.dtype == torch.float16          # 16-bit feature storage: used exactly (see mesm_inputs.video_feat_f16)
        if f16:
            if not video_feat.is_cuda:
                raise RuntimeError("mesm_b200: `video_feat` must be a CUDA tensor (no CPU fallback)")
            if video_len is None:
                raise RuntimeError("mesm_b200: fp16 `video_feat` needs the host clip counts (`video_len`)")
            video_feat = video_feat.contiguous()
        else:
            video_feat = _f32(video_feat, "video_feat")
        words_feat = _f32(words_feat, "words_feat")
        vmask = _u8(video_mask, "video_mask")
        B, Lv, Dv = video_feat.shape
        Lt = words_feat.shape[1]
        if Dv != self.cfg.v_feat_dim or words_feat.shape[2] != self.cfg.t_feat_dim or words_feat.shape[0] != B:
            raise RuntimeError("mesm_b200: feature shapes do not match the configuration")
        nc = [int(x) for x in (num_clips.tolist() if torch.is_tensor(num_clips) else num_clips)]
        nc_arr = (c_int64 * len(nc))(*nc)
        nq, nl, dev = self.cfg.num_queries, self.cfg.dec_layers, self.device
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        o = {}
        if "core" in want:
            o.update(pred_logits=f(B, nq, 2), pred_spans=f(B, nq, 2), saliency_scores=f(B, Lv))
        if neg_index is not None:
            o["neg_saliency_scores"] = f(B, Lv)
            neg_index = neg_index.to(device=dev, dtype=torch.int64).contiguous()
        if "aux" in want and nl > 1:
            o.update(aux_logits=f(nl - 1, B, nq, 2), aux_spans=f(nl - 1, B, nq, 2))
        if "rec" in want:
            o.update(projed_video_feat=f(B, Lv, 256), enhanced_video_feat=f(B, Lv, 256), recon_feat=f(B, 256),
                     projed_recon_feat=f(B, 256), expanded_words_feat=f(B, Lt + 1, 256),
                     expanded_words_mask=torch.empty(B, Lt + 1, dtype=torch.uint8, device=dev))
        if "taps" in want:
            o.update(memory=f(B, Lv, 256), memory_global=f(B, 256), hs=f(nl, B, nq, 256))
        vl_arr = None
        if video_len is not None:
            if torch.is_tensor(video_len):
                if video_len.is_cuda:
                    raise RuntimeError("mesm_b200: `video_len` must live on the host (a device tensor would force a sync)")
                video_len = video_len.tolist()
            vl = [int(x) for x in video_len]
            if len(vl) != B or min(vl) < 1 or max(vl) > Lv:
                raise RuntimeError("mesm_b200: `video_len` must hold B values in [1, Lv]")
            vl_arr = (ctypes.c_int32 * B)(*vl)
        if os.environ.get("MESM_DEBUG_CHECKS"):      # device-side contracts the fast path takes on trust (each check synchronises)
            if neg_index is not None and (int(neg_index.min()) < 0 or int(neg_index.max()) >= B):
                raise RuntimeError("mesm_b200: `neg_index` entries must be in [0, B)")
            if video_len is not None:
                m = vmask.bool()
                ar = torch.arange(Lv, device=dev)[None]
                if not torch.equal(m, ar < torch.tensor(vl, device=dev)[:, None]):
                    raise RuntimeError("mesm_b200: `video_len` does not describe `video_mask` (valid clips must be a prefix)")
        inp = MesmInputs(B, Lv, Lt, len(nc), _ptr(video_feat), _ptr(vmask), _ptr(words_feat), nc_arr, _ptr(neg_index), vl_arr,
                         int(bool(shared_group_video)), int(f16))
        out = MesmOutputs(**{k: _ptr(v) for k, v in o.items()})
        # a video group is never split: the internal chunk must hold the largest group
        self.lib.mesm_set_chunk_pairs(self.ctx, max(self.chunk_pairs, max(nc)))
        with torch.cuda.device(dev):
            need = self.lib.mesm_workspace_bytes(self.ctx, B, Lv, Lt, len(nc))
            ws = self._workspace(need)
            t_call = time.perf_counter()
            check(self.lib.mesm_forward(self.ctx, byref(inp), byref(out), _ptr(ws), ws.numel(), _stream()), self.ctx)
            self.last_enqueue_s = (t_call - t_enter, time.perf_counter() - t_call)     # host time: set-up, launch enqueue
        if "expanded_words_mask" in o:
            o["expanded_words_mask"] = o["expanded_words_mask"].view(torch.bool)
        return o

    def capture(self, video_feat, video_mask, words_feat, num_clips, neg_index=None, want=("core",), video_len=None,
                shared_group_video=False, decode=None, warmup=2):
        """Record ``forward`` (and, with ``decode=dict(duration=..., clip_len=..., max_ts_val=..., nms_thd=...)``, the span decode /
        NMS) for THIS batch signature into a CUDA graph and return a ``CapturedForward``: ``replay()`` re-runs the ~190 launches
        as one graph launch on the current contents of the (static) input tensors - refresh them in place with ``copy_``.  The
        ragged structure (num_clips, video_len) is part of the signature.  For small batches the per-launch CPU cost dominates the
        eager call; a replay removes it (bench.py: latency_b32)."""
        return CapturedForward(self, dict(video_feat=video_feat, video_mask=video_mask, words_feat=words_feat, num_clips=num_clips,
                                          neg_index=neg_index, want=want, video_len=video_len, shared_group_video=shared_group_video),
                               decode, warmup)

    @property
    def last_launch_count(self):
        return int(self.lib.mesm_last_launch_count(self.ctx))

    @property
    def last_feature_bytes(self):
        return int(self.lib.mesm_last_feature_bytes(self.ctx))


class CapturedForward:
    """A forward (+ decode) recorded into a CUDA graph by ``Engine.capture``."""

    def __init__(self, eng, kw, decode, warmup):
        self.eng, self.kw, self.decode = eng, kw, decode
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):               # warm-up outside the capture: workspace, function attributes, lazy module loads
            for _ in range(max(1, warmup)):
                self._run()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        # relaxed: mesm_forward allocates the pinned table of a captured forward inside the capture (a host allocation, not recorded)
        with torch.cuda.graph(self.graph, capture_error_mode="relaxed"):
            self.out = self._run()
        self.launches = eng.last_launch_count + (1 if decode else 0)

    def _run(self):
        o = self.eng.forward(**self.kw)
        if self.decode:
            d = self.decode
            o["windows"], o["order"], o["keep"], o["keep_count"] = decode_nms(
                o["pred_logits"], o["pred_spans"], d["duration"], d["clip_len"], d["max_ts_val"], d.get("nms_thd", -1.0),
                d.get("max_before_nms", 10), d.get("max_after_nms", 10))
        return o

    def replay(self):
        self.graph.replay()
        return self.out


# ---- span decode / NMS / span utils (context-free entry points) ------------------------------------------------------
def decode_nms(pred_logits, pred_spans, duration, clip_len, max_ts_val, nms_thd=-1.0, max_before_nms=10,
               max_after_nms=10, sort_results=True, min_ts_val=0.0):
    """eval.py:64-66,84-91 + PostProcessorDETR (eval.py:111-115) + temporal NMS (eval.py:476-485) on the device.
    Returns windows f64[B,nq,3] (ranked [st,ed,score]), order i32[B,nq], keep i32[B,max_after] (-1 padded),
    keep_count i32[B] (the last two only when nms_thd != -1)."""
    lib = _lib.lib()
    lg = _f32(pred_logits, "pred_logits"); sp = _f32(pred_spans, "pred_spans"); du = _f32(duration, "duration")
    B, nq = lg.shape[0], lg.shape[1]
    dev = lg.device
    windows = torch.empty(B, nq, 3, dtype=torch.float64, device=dev)
    order = torch.empty(B, nq, dtype=torch.int32, device=dev)
    do_nms = float(nms_thd) != -1.0
    keep = torch.empty(B, max_after_nms, dtype=torch.int32, device=dev) if do_nms else None
    cnt = torch.empty(B, dtype=torch.int32, device=dev) if do_nms else None
    p = MesmDecodeParams(float(clip_len), float(min_ts_val), float(max_ts_val), float(nms_thd), int(max_before_nms),
                         int(max_after_nms), int(bool(sort_results)))
    with torch.cuda.device(dev):
        check(lib.mesm_decode_nms(_ptr(lg), _ptr(sp), _ptr(du), B, nq, byref(p), _ptr(windows), _ptr(order), _ptr(keep),
                                  _ptr(cnt), _stream()))
    return windows, order, keep, cnt


def temporal_nms_lists(windows, offsets, nms_thd, max_after_nms):
    """Ragged-list utils.temporal_nms: windows f64[total,3], offsets i64[n+1] (device).  -> keep i32[n,max_after], count
    (count -1: that list has more than 1024 candidates and was not processed)."""
    lib = _lib.lib()
    n = offsets.numel() - 1
    keep = torch.empty(n, max_after_nms, dtype=torch.int32, device=windows.device)
    cnt = torch.empty(n, dtype=torch.int32, device=windows.device)
    with torch.cuda.device(windows.device):
        check(lib.mesm_temporal_nms(_ptr(windows.contiguous()), _ptr(offsets.contiguous()), n, float(nms_thd),
                                    int(max_after_nms), _ptr(keep), _ptr(cnt), _stream()))
    return keep, cnt


def align_scores(projed_video_feat, clip_mask, expanded_words_feat, expanded_words_mask, tau=0.5):
    """Segment-sentence alignment scores, model/criterion.py:241-266 (cos_sim / tau)."""
    lib = _lib.lib()
    pv = _f32(projed_video_feat, "projed_video_feat"); ew = _f32(expanded_words_feat, "expanded_words_feat")
    cm = _u8(clip_mask, "clip_mask"); em = _u8(expanded_words_mask, "expanded_words_mask")
    B, Lv, _ = pv.shape
    Lw = ew.shape[1]
    S = torch.empty(B, B, dtype=torch.float32, device=pv.device)
    ws = torch.empty(lib.mesm_align_workspace_bytes(B), dtype=torch.uint8, device=pv.device)
    with torch.cuda.device(pv.device):
        check(lib.mesm_align_scores(_ptr(pv), _ptr(cm), _ptr(ew), _ptr(em), B, Lv, Lw, float(tau), _ptr(S), _ptr(ws),
                                    ws.numel(), _stream()))
    return S


def loss_saliency(outputs, targets, rank_coef=12, use_triplet=False, saliency_margin=0.2):
    """Criterion.loss_saliency (model/criterion.py:139-221) on the forward's outputs, as train.py's per-epoch evaluation calls it
    (eval.py:101-105).  ``outputs``: dict with saliency_scores / neg_saliency_scores [B,L]; ``targets``: the batch dict
    (video_mask, saliency_label or clip_mask, and pos_idx / neg_idx when ``use_triplet``).  Returns {"loss_saliency": 0-d tensor}
    plus the three terms it is the sum of."""
    lib = _lib.lib()
    sal = _f32(outputs["saliency_scores"], "saliency_scores"); neg = _f32(outputs["neg_saliency_scores"], "neg_saliency_scores")
    vm = _u8(targets["video_mask"], "video_mask")
    label = targets["saliency_label"] if "saliency_label" in targets else targets["clip_mask"]
    label = _f32(label.to(sal.device), "saliency_label")
    B, L = sal.shape
    pos = neg_i = None
    P = 0
    if use_triplet:
        pos = targets["pos_idx"].to(device=sal.device, dtype=torch.int64).contiguous()
        neg_i = targets["neg_idx"].to(device=sal.device, dtype=torch.int64).contiguous()
        P = pos.shape[1]
    out = torch.empty(4, dtype=torch.float32, device=sal.device)
    ws = torch.empty(lib.mesm_saliency_loss_workspace_bytes(B), dtype=torch.uint8, device=sal.device)
    with torch.cuda.device(sal.device):
        check(lib.mesm_saliency_loss(_ptr(sal), _ptr(neg), _ptr(vm), _ptr(label), B, L, float(rank_coef), _ptr(pos), _ptr(neg_i), P,
                                     float(saliency_margin), _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return {"loss_saliency": out[0], "loss_neg_pair": out[1], "loss_rank_contrastive": out[2], "loss_triplet": out[3]}
