"""GPU parity of span decode / post-processing / temporal NMS and the span_utils device functions (bit-exact work)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import decode_oracle
from tests.helpers import golden_cases, load_case

pytestmark = pytest.mark.gpu


def _check_batch(lg, sp, dur, clip_len, max_ts, thd, nb, na):
    import mesm_b200
    win, order, keep, cnt = mesm_b200.decode_nms(torch.from_numpy(lg).cuda(), torch.from_numpy(sp).cuda(),
                                                 torch.from_numpy(dur).cuda(), clip_len, max_ts, thd, nb, na)
    win, order = win.cpu().numpy(), order.cpu().numpy()
    if thd == -1.0:
        assert keep is None and cnt is None
    else:
        keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    n_score_ulp = 0
    for i in range(lg.shape[0]):
        od = decode_oracle.decode_pair(lg[i], sp[i], float(dur[i]), clip_len, max_ts, thd, nb, na)
        w = np.asarray(od["windows"])
        assert od["order"] == order[i].tolist(), i
        assert np.array_equal(w[:, :2], win[i, :, :2]), i                     # st / ed bit-exact
        d = np.abs(w[:, 2] - win[i, :, 2])
        assert d.max() <= 1.0001e-4, i                                         # score: <= one 4-decimal quantum (expf ulp)
        n_score_ulp += int((d > 0).sum())
        if thd != -1.0:
            assert od["keep"] == keep[i, :cnt[i]].tolist(), i                  # identical kept-span index sets
            assert (keep[i, cnt[i]:] == -1).all()
    return n_score_ulp


@pytest.mark.parametrize("name", sorted(golden_cases()))
def test_decode_matches_reference_golden(name):
    import mesm_b200
    cfg, sd, inp, neg, gold, meta = load_case(name)
    win, order, keep, cnt = mesm_b200.decode_nms(torch.from_numpy(gold["pred_logits"]).cuda(), torch.from_numpy(gold["pred_spans"]).cuda(),
                                                 inp["duration"].cuda(), cfg.clip_len, cfg.max_ts_val, meta["nms_thd"], 10, 10)
    assert np.array_equal(order.cpu().numpy(), gold["order"])
    assert np.array_equal(win.cpu().numpy()[..., :2], gold["windows"][..., :2])
    assert np.abs(win.cpu().numpy()[..., 2] - gold["windows"][..., 2]).max() <= 1.0001e-4
    for i in range(len(gold["nms_count"])):
        n = int(gold["nms_count"][i])
        assert int(cnt[i]) == n
        kept_rows = [gold["order"][i].tolist().index(q) for q in keep[i, :n].tolist()]
        assert np.array_equal(gold["windows"][i][kept_rows][:, :2], gold["nms_windows"][i, :n, :2])


@pytest.mark.parametrize("clip_len,max_ts,thd", [(2.0, 150.0, 0.7), (1.0, 150.0, 0.5), (0.17, 150.0, 0.7), (-1.0, 1000.0, 0.3),
                                                 (2.0, 150.0, -1.0)])
def test_decode_random_vs_oracle(clip_len, max_ts, thd):
    rng = np.random.default_rng(7)
    B, nq = 2048, 10
    lg = rng.normal(size=(B, nq, 2)).astype(np.float32) * 2
    sp = np.stack([rng.uniform(0, 1, (B, nq)), rng.uniform(0, 0.6, (B, nq))], -1).astype(np.float32)
    # edge cases: exact score ties, identical windows, spans outside [0,1], zero width
    lg[::7, 3] = lg[::7, 5]
    sp[::5, 2] = sp[::5, 4]
    sp[::11, 6, 1] = 0
    sp[::13, 1] = [0.02, 0.5]
    dur = rng.uniform(5, 150, B).astype(np.float32)
    _check_batch(lg, sp, dur, clip_len, max_ts, thd, 10, 10)
    _check_batch(lg[:256], sp[:256], dur[:256], clip_len, max_ts, thd, 6, 3)


def test_decode_other_query_counts():
    rng = np.random.default_rng(3)
    for nq in (1, 2, 17, 32):
        lg = rng.normal(size=(64, nq, 2)).astype(np.float32)
        sp = np.stack([rng.uniform(0, 1, (64, nq)), rng.uniform(0, 0.5, (64, nq))], -1).astype(np.float32)
        dur = rng.uniform(5, 150, 64).astype(np.float32)
        _check_batch(lg, sp, dur, 2.0, 150.0, 0.7, nq, nq)


def test_temporal_nms_dense_candidates():
    """TACoS-style dense candidate lists (100 candidates / pair) + ragged / degenerate lists."""
    import mesm_b200
    rng = np.random.default_rng(11)
    sizes = [100] * 40 + [1, 2, 0, 3, 1000, 57]
    lists = []
    for n in sizes:
        st = rng.uniform(0, 140, n).round(2)
        w = np.stack([st, st + rng.uniform(0, 30, n).round(2), rng.uniform(0, 1, n).round(3)], 1)
        if n >= 10:
            w[5] = w[2]                      # duplicates + score ties
            w[7, 2] = w[3, 2]
        lists.append(w)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = torch.from_numpy(np.concatenate([l for l in lists if len(l)] or [np.zeros((0, 3))])).cuda()
    for thd, na in ((0.7, 10), (0.3, 100), (0.5, 1)):
        keep, cnt = mesm_b200.temporal_nms_lists(flat, torch.from_numpy(offs).cuda(), thd, na)
        keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
        for i, w in enumerate(lists):
            exp = decode_oracle.temporal_nms(w.tolist(), thd, na)[1] if len(w) else []
            assert keep[i, :cnt[i]].tolist() == exp, (i, thd, na)


def test_span_utils_known_answers(golden_dir):
    """Doctest vectors of utils/span_utils.py (12-19, 31-38, 54-60, 105-109) + random bit-exact comparison."""
    from mesm_b200 import utils as U
    g = np.load(os.path.join(golden_dir, "span_utils_doctest.npz"))
    s1, s2 = torch.from_numpy(g["s1"]).cuda(), torch.from_numpy(g["s2"]).cuda()
    iou, union = U.temporal_iou(s1, s2)
    assert np.array_equal(iou.cpu().numpy(), g["iou"]) and np.array_equal(union.cpu().numpy(), g["union"])
    assert np.array_equal(U.generalized_temporal_iou(s1, s2).cpu().numpy(), g["giou"])
    assert np.array_equal(U.span_xx_to_cxw(torch.tensor([[0, 1], [0.2, 0.4]]).cuda()).cpu().numpy(), g["cxw"])
    assert np.array_equal(U.span_cxw_to_xx(torch.tensor([[[0.5, 1.0], [0.3, 0.2]]]).cuda()).cpu().numpy()[0], g["xx"])
    rng = np.random.default_rng(0)
    a = np.sort(rng.uniform(0, 1, (37, 2)).astype(np.float32), axis=1)
    b = np.sort(rng.uniform(0, 1, (53, 2)).astype(np.float32), axis=1)
    iou, union = U.temporal_iou(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())
    ri, ru = decode_oracle.temporal_iou(a, b)
    assert np.array_equal(iou.cpu().numpy(), ri.astype(np.float32)) and np.array_equal(union.cpu().numpy(), ru.astype(np.float32))
    assert np.array_equal(U.generalized_temporal_iou(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()).cpu().numpy(),
                          decode_oracle.generalized_temporal_iou(a, b))


def test_python_wrappers_keep_reference_signatures():
    """utils.temporal_nms(predictions, nms_thd, max_after_nms) and PostProcessorDETR(...)(lines) as thin wrappers."""
    from mesm_b200 import utils as U
    preds = [[0.0, 10.0, 0.9], [3.0, 10.0, 0.8], [1.0, 9.5, 0.85], [50.0, 60.0, 0.1]]
    assert U.temporal_nms(preds, 0.7, 10) == decode_oracle.temporal_nms(preds, 0.7, 10)[0]
    lines = [dict(qid=1, pred_relevant_windows=[[1.23456, 151.0, 0.55555], [-3.0, 7.1, 0.4]])]
    pp = U.PostProcessorDETR(clip_length=2, min_ts_val=0, max_ts_val=150, process_func_names=("clip_ts", "round_multiple"))
    out = pp(lines)[0]["pred_relevant_windows"]
    assert out == decode_oracle.post_process_windows([[1.23456, 151.0, 0.55555], [-3.0, 7.1, 0.4]], 2, 150)


def test_decode_unsorted_list_still_ranks_inside_nms():
    """sort_results = 0 (eval.py:89 skipped): the windows stay in query order, but utils/temporal_nms.py:41 sorts its own
    input by the rounded score, so the kept set must equal the oracle's on the unsorted list."""
    import mesm_b200
    rng = np.random.default_rng(21)
    B, nq = 512, 10
    lg = rng.normal(size=(B, nq, 2)).astype(np.float32) * 2
    lg[::5, 2] = lg[::5, 6]
    sp = np.stack([rng.uniform(0, 1, (B, nq)), rng.uniform(0, 0.6, (B, nq))], -1).astype(np.float32)
    dur = rng.uniform(5, 150, B).astype(np.float32)
    for nb, na in ((10, 10), (6, 3)):
        win, order, keep, cnt = mesm_b200.decode_nms(torch.from_numpy(lg).cuda(), torch.from_numpy(sp).cuda(), torch.from_numpy(dur).cuda(),
                                                     2.0, 150.0, 0.5, nb, na, sort_results=False)
        order, keep, cnt = order.cpu().numpy(), keep.cpu().numpy(), cnt.cpu().numpy()
        for i in range(B):
            od = decode_oracle.decode_pair(lg[i], sp[i], float(dur[i]), 2.0, 150.0, 0.5, nb, na, sort_results=False)
            assert order[i].tolist() == list(range(nq))
            assert od["keep"] == keep[i, :cnt[i]].tolist(), i


def test_single_candidate_and_overlong_lists():
    """utils/temporal_nms.py:38-39 returns a single prediction untouched whatever max_after_nms is; lists beyond the kernel's
    1024-entry tables are flagged (count -1), not processed."""
    import mesm_b200
    from mesm_b200 import utils as U
    lg = torch.randn(8, 10, 2).cuda()
    sp = torch.rand(8, 10, 2).cuda() * 0.5
    dur = torch.full((8,), 30.0).cuda()
    win, order, keep, cnt = mesm_b200.decode_nms(lg, sp, dur, 1.0, 150.0, 0.7, 1, 0)
    assert cnt.tolist() == [1] * 8 and keep.shape == (8, 0)
    win, order, keep, cnt = mesm_b200.decode_nms(lg, sp, dur, 1.0, 150.0, 0.7, 1, 3)
    assert cnt.tolist() == [1] * 8 and torch.equal(keep[:, 0], order[:, 0]) and (keep[:, 1:] == -1).all()
    n = 1500
    w = torch.rand(n + 5, 3, dtype=torch.float64).cuda()
    offs = torch.tensor([0, n, n + 5], dtype=torch.int64).cuda()
    keep, cnt = mesm_b200.temporal_nms_lists(w, offs, 0.5, 10)
    assert int(cnt[0]) == -1 and (keep[0] == -1).all() and int(cnt[1]) >= 1
    with pytest.raises(ValueError):
        U.temporal_nms(w[:n].tolist(), 0.5, 10)


def test_decode_and_nms_match_reference_random_fixture(golden_dir):
    """CUDA decode / NMS against the reference-generated random fixture (tests/golden/decode_random.npz)."""
    import mesm_b200
    from oracle.config import CONFIGS
    g = np.load(os.path.join(golden_dir, "decode_random.npz"))
    offs, koffs, params = g["nms_offsets"], g["nms_kept_offsets"], g["nms_params"]
    flat = torch.from_numpy(g["nms_lists"]).cuda()
    for thd, na in sorted({(float(a), int(b)) for a, b in params}):
        sel = [i for i in range(len(params)) if float(params[i, 0]) == thd and int(params[i, 1]) == na]
        sub_offs = np.concatenate([[0], np.cumsum([offs[i + 1] - offs[i] for i in sel])]).astype(np.int64)
        sub = torch.cat([flat[offs[i]:offs[i + 1]] for i in sel])
        keep, cnt = mesm_b200.temporal_nms_lists(sub, torch.from_numpy(sub_offs).cuda(), thd, na)
        keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
        for j, i in enumerate(sel):
            w = g["nms_lists"][offs[i]:offs[i + 1]]
            ref = g["nms_kept"][koffs[i]:koffs[i + 1]]
            assert cnt[j] == len(ref) and np.array_equal(w[keep[j, :cnt[j]]], ref), (i, thd, na)
    mism = 0
    for cname in ("qvhighlights", "charades_csf", "charades_vgg", "tacos"):
        cfg = CONFIGS[cname]
        lg, sp, dur = (torch.from_numpy(g[f"dec_{cname}_{k}"]).cuda() for k in ("logits", "spans", "duration"))
        win, order, keep, cnt = mesm_b200.decode_nms(lg, sp, dur, cfg.clip_len, cfg.max_ts_val, 0.7, 10, 10)
        win, order, keep, cnt = win.cpu().numpy(), order.cpu().numpy(), keep.cpu().numpy(), cnt.cpu().numpy()
        rw = g[f"dec_{cname}_windows"]
        assert np.array_equal(order, g[f"dec_{cname}_order"]) and np.array_equal(win[..., :2], rw[..., :2])
        d = np.abs(win[..., 2] - rw[..., 2])
        assert d.max() <= 1.0001e-4                    # torch's fp32 softmax vs the exactly rounded score: at most one 4-decimal quantum
        mism += int((d > 0).sum())
        assert np.array_equal(cnt, g[f"dec_{cname}_nms_count"])
        for i in range(len(cnt)):
            pos = [order[i].tolist().index(q) for q in keep[i, :cnt[i]].tolist()]
            assert np.array_equal(rw[i][pos][:, :2], g[f"dec_{cname}_nms_windows"][i, :cnt[i], :2]), (cname, i)
    assert mism <= 12, mism                             # 12000 scores; a handful may sit within an ulp of a rounding boundary
