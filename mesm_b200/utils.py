"""Drop-in mirrors of the reference's ``utils`` helpers on the hot path, executed by libmesm_b200.so.

    utils/span_utils.py:5-42     span_xx_to_cxw, span_cxw_to_xx      -> mesm_span_convert
    utils/span_utils.py:45-121   temporal_iou, generalized_temporal_iou -> mesm_temporal_iou
    utils/temporal_nms.py:25-74  temporal_nms(predictions, nms_thd, max_after_nms) -> mesm_temporal_nms
    utils/post_processing.py:5-47 PostProcessorDETR(...)(lines)      -> mesm_post_process

Tensors must live on the GPU (no CPU fallback); the list-based wrappers (temporal_nms, PostProcessorDETR) keep the
reference's Python-list signatures and move the numbers to ``cuda:current`` and back.
"""
import torch

from . import _lib
from .engine import _f32, _ptr, _stream, temporal_nms_lists
from ._lib import check


def _convert(spans, to_xx):
    s = _f32(spans, "spans")
    out = torch.empty_like(s)
    with torch.cuda.device(s.device):
        check(_lib.lib().mesm_span_convert(_ptr(s), _ptr(out), s.numel() // 2, int(to_xx), _stream()))
    return out


def span_xx_to_cxw(xx_spans):
    """(st, ed) -> (center, width); any leading shape (utils/span_utils.py:5-23)."""
    return _convert(xx_spans, False)


def span_cxw_to_xx(cxw_spans):
    """(center, width) -> (st, ed) (utils/span_utils.py:26-42)."""
    return _convert(cxw_spans, True)


def _iou(spans1, spans2, want_giou):
    a, b = _f32(spans1, "spans1"), _f32(spans2, "spans2")
    N, M = a.shape[0], b.shape[0]
    iou = torch.empty(N, M, dtype=torch.float32, device=a.device)
    uni = torch.empty_like(iou)
    giou = torch.empty_like(iou) if want_giou else None
    with torch.cuda.device(a.device):
        check(_lib.lib().mesm_temporal_iou(_ptr(a), N, _ptr(b), M, _ptr(iou), _ptr(uni), _ptr(giou), _stream()))
    return iou, uni, giou


def temporal_iou(spans1, spans2):
    """(N,2),(M,2) -> iou (N,M), union (N,M)  (utils/span_utils.py:45-72)."""
    iou, uni, _ = _iou(spans1, spans2, False)
    return iou, uni


def generalized_temporal_iou(spans1, spans2):
    """utils/span_utils.py:92-121 (incl. its ed >= st assertions)."""
    spans1, spans2 = spans1.float(), spans2.float()
    assert (spans1[:, 1] >= spans1[:, 0]).all()
    assert (spans2[:, 1] >= spans2[:, 0]).all()
    return _iou(spans1, spans2, True)[2]


def temporal_nms(predictions, nms_thd, max_after_nms=100):
    """utils/temporal_nms.py:25-74 on a Python list of [st, ed, score]; returns the surviving sublists in order."""
    if len(predictions) == 0:
        return []
    if len(predictions) > 1024:
        raise ValueError("temporal_nms: at most 1024 candidates per list")
    dev = torch.device("cuda", torch.cuda.current_device())
    w = torch.tensor([list(map(float, p[:3])) for p in predictions], dtype=torch.float64, device=dev)
    offs = torch.tensor([0, len(predictions)], dtype=torch.int64, device=dev)
    keep, cnt = temporal_nms_lists(w, offs, nms_thd, max(int(max_after_nms), 1) if len(predictions) == 1 else int(max_after_nms))
    if int(cnt[0]) < 0:
        raise ValueError("temporal_nms: at most 1024 candidates per list")
    idx = keep[0, :int(cnt[0])].tolist()
    return [list(predictions[i]) if len(predictions) == 1 else [predictions[i][0], predictions[i][1], predictions[i][2]] for i in idx]


class PostProcessorDETR:
    """utils/post_processing.py:5-47 with the process functions eval.py:111-115 uses ("clip_ts", "round_multiple")."""

    def __init__(self, clip_length=2, min_ts_val=0, max_ts_val=150, min_w_l=2, max_w_l=70, move_window_method="center",
                 process_func_names=("clip_window_l", "clip_ts", "round_multiple")):
        if "clip_window_l" in process_func_names:
            raise NotImplementedError("clip_window_l is not among the functions eval.py:111-115 passes")
        self.clip_length, self.min_ts_val, self.max_ts_val = clip_length, min_ts_val, max_ts_val
        self.process_func_names = tuple(process_func_names)

    def __call__(self, lines):
        dev = torch.device("cuda", torch.cuda.current_device())
        lib = _lib.lib()
        clip = float(self.clip_length) if "round_multiple" in self.process_func_names else -1.0
        lo = float(self.min_ts_val) if "clip_ts" in self.process_func_names else float("-inf")
        hi = float(self.max_ts_val) if "clip_ts" in self.process_func_names else float("inf")
        counts = [len(l["pred_relevant_windows"]) for l in lines]
        flat = [list(map(float, w[:3])) for l in lines for w in l["pred_relevant_windows"]]
        if flat:
            w = torch.tensor(flat, dtype=torch.float64, device=dev)
            out = torch.empty_like(w)
            with torch.cuda.device(dev):
                check(lib.mesm_post_process(_ptr(w), _ptr(out), w.shape[0], clip, lo, hi, _stream()))
            rows = out.tolist()
        else:
            rows = []
        k = 0
        for l, n in zip(lines, counts):
            l["pred_relevant_windows"] = rows[k:k + n]
            k += n
        return lines


def eval_moment_retrieval(windows, gt_windows, gt_offsets, dataset_name="charades", max_pred_windows=10):
    """eval_moment_retrieval (eval.py:233-263) on the device: MR-mAP and MR-R1 per ground-truth length range.

    windows: f64[B,nq,3] ranked [st,ed,score] (``decode_nms``); gt_windows f64[total,2] + gt_offsets i64[B+1] (CUDA tensors): the
    ``relevant_windows`` of query b are rows gt_offsets[b]:gt_offsets[b+1].  Returns the reference's nested dict
    {range: {"MR-mAP": {thd: %, "average": %}, "MR-R1": {thd: %, "miou": %}}} with its 2-decimal percent formatting; ranges
    without any query are skipped like eval.py:249-250."""
    import numpy as np
    lib = _lib.lib()
    dev = windows.device
    if dataset_name in ("tacos",):
        ranges, names, max_len = [[0, 10], [10, 30], [30, 150], [150, 600], [0, 600]], ["short", "middle", "long", "superlong", "full"], 600
        r1_thds = np.array([0.1, 0.3, 0.5, 0.7])
    else:
        ranges, names, max_len = [[0, 10], [10, 30], [30, 150], [0, 150]], ["short", "middle", "long", "full"], 150
        r1_thds = np.concatenate([np.array([0.3]), np.linspace(0.5, 0.95, 10)])
    ap_thds = [float(f"{e:.2f}") for e in np.linspace(0.5, 0.95, 10)]
    r1_thds = [float(f"{e:.2f}") for e in r1_thds]
    rg = torch.tensor([[-1.0, 0.0] if (a == 0 and b == max_len) else [float(a), float(b)] for a, b in ranges], dtype=torch.float64, device=dev)
    B, nq = windows.shape[0], windows.shape[1]
    nr, nt = len(ranges), len(ap_thds)
    w = windows.to(torch.float64).contiguous()
    in_range = torch.empty(nr, B, dtype=torch.uint8, device=dev)
    top1 = torch.empty(nr, B, dtype=torch.float64, device=dev)
    ap = torch.empty(nr, B, nt, dtype=torch.float64, device=dev)
    thd_t = torch.tensor(ap_thds, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.mesm_mr_metrics(_ptr(w), B, nq, int(max_pred_windows), _ptr(gt_windows.to(torch.float64).contiguous()),
                                  _ptr(gt_offsets.to(torch.int64).contiguous()), _ptr(rg), nr, _ptr(thd_t), nt, _ptr(in_range), _ptr(top1),
                                  _ptr(ap), _stream()))
    fmt = lambda v: float(f"{100 * v:.2f}")
    out = {}
    in_range, top1, ap = in_range.bool().cpu().numpy(), top1.cpu().numpy(), ap.cpu().numpy()
    for r, name in enumerate(names):
        sel = in_range[r]
        if not sel.any():
            continue
        ap_thd = ap[r][sel].mean(0)
        m_ap = dict(zip([str(e) for e in ap_thds], ap_thd))
        m_ap["average"] = np.mean(ap_thd)
        iou = top1[r][sel]
        m_r1 = {str(t): fmt(np.mean(iou >= t)) for t in r1_thds}
        m_r1["miou"] = fmt(iou.mean())
        out[name] = {"MR-mAP": {k: fmt(v) for k, v in m_ap.items()}, "MR-R1": m_r1}
    return out
