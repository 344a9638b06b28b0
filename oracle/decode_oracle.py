"""CPU restatement of MESM's span decode -> post-process -> temporal NMS chain.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain Python / numpy scalar code; every step cites the reference
lines it follows.  Parity status: PINNED by oracle/gen_golden.py against ``utils.PostProcessorDETR``,
``utils.temporal_nms`` and the ``eval.py:64-99`` loop run on the same tensors (fixtures in tests/golden/), and by
the span_utils doctest vectors (utils/span_utils.py:12-19, 31-38, 54-60, 105-109, 133-139).
"""
import numpy as np


def round4(x: float) -> float:
    """``float(f"{x:.4f}")`` (eval.py:91, utils/post_processing.py:31) for a float32-valued double:
    x*1e4 is exact in fp64 (<= 24+14 bits), rint is half-even like Python's correctly-rounded formatting,
    and k/1e4 is the correctly rounded double of the printed decimal."""
    return float(np.rint(np.float64(x) * 1e4) / 1e4)


def softmax_fg(logits: np.ndarray) -> np.ndarray:
    """F.softmax(pred_logits, -1)[..., 0] in fp32 (eval.py:64-66): exp(x - max) / sum."""
    l = logits.astype(np.float32)
    m = l.max(-1, keepdims=True)
    e = np.exp(l - m, dtype=np.float32)
    return (e[..., 0] / e.sum(-1, dtype=np.float32)).astype(np.float32)


def span_cxw_to_xx(cxw: np.ndarray) -> np.ndarray:
    """utils/span_utils.py:40-42 in the input dtype."""
    half = np.asarray(0.5, dtype=cxw.dtype)
    return np.stack([cxw[..., 0] - half * cxw[..., 1], cxw[..., 0] + half * cxw[..., 1]], axis=-1)


def span_xx_to_cxw(xx: np.ndarray) -> np.ndarray:
    """utils/span_utils.py:21-23."""
    half = np.asarray(0.5, dtype=xx.dtype)
    return np.stack([xx.sum(-1) * half, xx[..., 1] - xx[..., 0]], axis=-1)


def temporal_iou(s1: np.ndarray, s2: np.ndarray):
    """utils/span_utils.py:61-72 (true union; 0/0 -> nan like torch)."""
    a1 = s1[:, 1] - s1[:, 0]
    a2 = s2[:, 1] - s2[:, 0]
    left = np.maximum(s1[:, None, 0], s2[None, :, 0])
    right = np.minimum(s1[:, None, 1], s2[None, :, 1])
    inter = np.clip(right - left, 0, None)
    union = a1[:, None] + a2[None, :] - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / union, union


def generalized_temporal_iou(s1: np.ndarray, s2: np.ndarray):
    """utils/span_utils.py:110-121."""
    s1 = s1.astype(np.float32)
    s2 = s2.astype(np.float32)
    assert (s1[:, 1] >= s1[:, 0]).all() and (s2[:, 1] >= s2[:, 0]).all()
    iou, union = temporal_iou(s1, s2)
    left = np.minimum(s1[:, None, 0], s2[None, :, 0])
    right = np.maximum(s1[:, None, 1], s2[None, :, 1])
    enclosing = np.clip(right - left, 0, None)
    with np.errstate(divide="ignore", invalid="ignore"):
        return iou - (enclosing - union) / enclosing


def hull_iou(a, b) -> float:
    """utils/temporal_nms.py:6-22 — 'IoU' over the convex hull, fp64, 0 when the hull is empty."""
    inter = max(0, min(a[1], b[1]) - max(a[0], b[0]))
    union = max(a[1], b[1]) - min(a[0], b[0])
    return 0 if union == 0 else 1.0 * inter / union


def temporal_nms(predictions, nms_thd, max_after_nms=100):
    """utils/temporal_nms.py:25-74 restated as index bookkeeping.
    Returns (kept windows, kept positions into ``predictions``)."""
    n = len(predictions)
    if n == 1:                                              # :38-39
        return [list(predictions[0])], [0]
    order = sorted(range(n), key=lambda i: predictions[i][2], reverse=True)   # stable, :41
    alive = list(order)
    kept = []
    while len(alive) > 1 and len(kept) < max_after_nms:     # :49
        head = predictions[alive[0]]
        alive = [alive[0]] + [j for j in alive[1:] if not hull_iou(head[:2], predictions[j][:2]) > nms_thd]  # :52
        kept.append(alive.pop(0))
    if len(kept) < max_after_nms and len(alive) >= 1:       # :68-71
        kept.append(alive.pop(0))
    return [list(predictions[i]) for i in kept], kept


def post_process_windows(windows, clip_len, max_ts_val, min_ts_val=0.0):
    """PostProcessorDETR.__call__ as configured at eval.py:111-115: torch.tensor(list) -> fp32, clamp
    (post_processing.py:35-40), round to multiples of clip_len unless clip_len == -1 (42-47, half-even, fp32),
    score re-rounded to 4 decimals (31)."""
    w = np.asarray(windows, dtype=np.float64).reshape(-1, 3).astype(np.float32)
    se = np.clip(w[:, :2], np.float32(min_ts_val), np.float32(max_ts_val))
    if clip_len != -1:
        cl = np.float32(clip_len)
        se = (np.rint(se / cl) * cl).astype(np.float32)
    return [[float(se[i, 0]), float(se[i, 1]), round4(float(w[i, 2]))] for i in range(len(w))]


def decode_pair(logits, spans, duration, clip_len, max_ts_val, nms_thd=-1.0, max_before_nms=10, max_after_nms=10,
                sort_results=True):
    """One pair through eval.py:64-66, 84-91, 111-116 and (if nms_thd != -1) eval.py:476-485.

    logits f32[nq,2], spans f32[nq,2] (center,width), duration f32 scalar.
    Returns dict(windows=[[st,ed,score]]*nq (post-processed, ranked), order=[query idx per rank],
                 nms_windows, keep=[query idx of survivors in output order])."""
    logits = np.asarray(logits, dtype=np.float32)
    spans = np.asarray(spans, dtype=np.float32)
    score = softmax_fg(logits)
    xx = (span_cxw_to_xx(spans) * np.float32(duration)).astype(np.float32)      # eval.py:86
    rows = [[float(xx[i, 0]), float(xx[i, 1]), float(score[i])] for i in range(len(score))]   # .tolist(), :88
    order = list(range(len(rows)))
    if sort_results:
        order = sorted(order, key=lambda i: rows[i][2], reverse=True)            # :89-90 (stable)
    ranked = [[round4(e) for e in rows[i]] for i in order]                       # :91
    windows = post_process_windows(ranked, clip_len, max_ts_val)
    res = dict(windows=windows, order=order)
    if nms_thd != -1:
        kept_w, kept_pos = temporal_nms(windows[:max_before_nms], nms_thd, max_after_nms)
        res.update(nms_windows=kept_w, keep=[order[p] for p in kept_pos], keep_pos=kept_pos)
    return res
