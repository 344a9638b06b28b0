"""GPU parity of the fused forward (through the C ABI) against the committed reference outputs and the oracle."""
import numpy as np
import pytest
import torch

from tests.helpers import engine_cfg, golden_cases, load_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-3          # north star: <= 1e-3 relative (max|a-b| / max|ref|) on logits, spans, saliency


def _run(name, chunk_pairs=256, packed=False, shared=False):
    import mesm_b200
    from mesm_b200.ingest import clip_counts
    cfg, sd, inp, neg, gold, meta = load_case(name)
    eng = mesm_b200.Engine(engine_cfg(cfg), chunk_pairs=chunk_pairs)
    eng.load_state_dict(sd)
    dev = eng.device
    vf = inp["video_feat"].to(dev)
    if meta["kwargs"].get("f16_features") and packed:
        vf = vf.half()                  # the 16-bit stored features themselves (packed rows only); padded: their fp32 upcast
    out = eng.forward(vf, inp["video_mask"].to(dev), inp["words_feat"].to(dev), inp["num_clips"],
                      neg_index=neg.to(dev), want=("core", "aux", "rec", "taps"),
                      video_len=clip_counts(inp["video_mask"]) if packed else None, shared_group_video=shared)
    torch.cuda.synchronize()
    return cfg, inp, gold, out


@pytest.mark.parametrize("packed", [False, True], ids=["padded", "packed"])
@pytest.mark.parametrize("name", sorted(golden_cases()))
def test_forward_matches_reference_golden(name, packed):
    """packed = the host passes the clip counts (video_len) and the forward runs on variable-length rows."""
    cfg, inp, gold, out = _run(name, packed=packed)
    vm = inp["video_mask"]
    errs = {
        "pred_logits": rel_err(out["pred_logits"], gold["pred_logits"]),
        "pred_spans": rel_err(out["pred_spans"], gold["pred_spans"]),
        "saliency_scores": rel_err(out["saliency_scores"], gold["saliency_scores"], vm),
        "neg_saliency_scores": rel_err(out["neg_saliency_scores"], gold["neg_saliency_scores"], vm),
        "recon_feat": rel_err(out["recon_feat"], gold["recon_feat"]),
        "projed_recon_feat": rel_err(out["projed_recon_feat"], gold["projed_recon_feat"]),
        "aux_logits": rel_err(out["aux_logits"][0], gold["aux_logits"]),
        "aux_spans": rel_err(out["aux_spans"][0], gold["aux_spans"]),
        "projed_words_feat": rel_err(out["expanded_words_feat"][:, 1:], gold["projed_words_feat"]),
        "projed_video_row0": rel_err(out["projed_video_feat"][:, 0], gold["projed_video_row0"]),
        "enhanced_video_row0": rel_err(out["enhanced_video_feat"][:, 0], gold["enhanced_video_row0"]),
    }
    print(name, {k: f"{v:.2e}" for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v <= TOL}
    assert not bad, bad
    # saliency additionally after the .half() of eval.py:68
    sal_h = out["saliency_scores"].half().float().cpu()
    ref_h = torch.from_numpy(gold["saliency_scores"]).half().float()
    assert rel_err(sal_h, ref_h, vm) <= 2e-3


@pytest.mark.parametrize("name", ["tiny_ragged", "charades_csf_ragged", "charades_csf_ragged_f16", "tacos_l96_f16", "tacos_l200"])
def test_per_video_projection_matches_reference_golden(name):
    """shared_group_video: the K = v_feat_dim projection runs once per video (and, for fp16 features, on the TMA-fed kernel with
    an exact fp16 operand plane); results = the reference's golden outputs."""
    cfg, inp, gold, out = _run(name, packed=True, shared=True)
    vm = inp["video_mask"]
    for k in ("pred_logits", "pred_spans", "recon_feat", "projed_recon_feat"):
        assert rel_err(out[k], gold[k]) <= TOL, k
    for k in ("saliency_scores", "neg_saliency_scores"):
        assert rel_err(out[k], gold[k], vm) <= TOL, k
    assert rel_err(out["projed_video_feat"][:, 0], gold["projed_video_row0"]) <= TOL
    assert rel_err(out["enhanced_video_feat"][:, 0], gold["enhanced_video_row0"]) <= TOL


@pytest.mark.parametrize("name", ["tiny_ragged", "charades_csf_ragged", "tiny_qvh_groups", "tacos_l96"])
@pytest.mark.parametrize("chunk_pairs", [256, 3])
def test_packed_rows_match_padded_rows(name, chunk_pairs):
    """Variable-length (packed) rows vs zero-padded rows: same numbers at every valid clip, zeros at the pad clips."""
    _, inp, gold, a = _run(name, chunk_pairs=256, packed=False)
    _, _, _, b = _run(name, chunk_pairs=chunk_pairs, packed=True)
    vm = inp["video_mask"]
    for k in ("pred_logits", "pred_spans", "aux_logits", "aux_spans", "recon_feat", "projed_recon_feat", "memory_global", "hs",
              "expanded_words_feat"):
        assert rel_err(b[k], a[k]) < 1e-4, k
    for k in ("saliency_scores", "neg_saliency_scores"):
        assert rel_err(b[k], a[k], vm) < 1e-4, k
        assert float(b[k].cpu()[~vm].abs().max() if (~vm).any() else 0.0) == 0.0, k
    for k in ("projed_video_feat", "enhanced_video_feat", "memory"):
        assert rel_err(b[k], a[k], vm[..., None]) < 1e-4, k
        assert float(b[k].cpu()[~vm].abs().max() if (~vm).any() else 0.0) == 0.0, k


@pytest.mark.parametrize("name", ["tiny_ragged", "charades_csf_ragged"])
def test_chunking_is_invisible(name):
    """Chunking by video group must not change results (the mask quirk and the negative branch reach across chunks)."""
    _, inp, gold, a = _run(name, chunk_pairs=256)
    _, _, _, b = _run(name, chunk_pairs=3)
    # (different chunk sizes route some GEMMs to the other linear kernel, so "equal" means equal to rounding)
    for k in ("pred_logits", "pred_spans", "saliency_scores", "neg_saliency_scores", "recon_feat"):
        m = inp["video_mask"] if "saliency" in k else None
        assert rel_err(a[k], b[k], m) < 1e-4, k


def test_repeated_forwards_are_bit_stable_and_watchdog_silent():
    """Bench-shaped ragged batch, repeated: identical bits every time and no barrier wait of the tcgen05 kernels gave up."""
    import ctypes
    import bench
    from mesm_b200 import _lib
    from mesm_b200.model import build_model
    torch.manual_seed(0)
    model = build_model(bench.CHARADES_CSF).cuda()
    wl = bench.make_workload(bench.CHARADES_CSF, 1024, 7, torch.device("cuda"))
    ref = None
    for _ in range(6):
        out = model(wl["video_feat"], wl["video_mask"], wl["words_feat"], None, None, wl["num_clips"], dataset_name="charades",
                    is_training=False, neg_index=wl["neg_index"], video_len=wl["video_len"])
        torch.cuda.synchronize()
        if ref is None:
            ref = {k: out[k].clone() for k in ("pred_logits", "pred_spans", "saliency_scores")}
        for k in ref:
            assert torch.equal(out[k], ref[k]), k
    wd = (ctypes.c_ulonglong * 128)()
    _lib.lib().mesm_debug_watchdog(wd)
    assert wd[0] == 0 and wd[64] == 0, list(wd)[:8] + list(wd)[64:72]
    assert not torch.isnan(ref["pred_logits"]).any()


@pytest.mark.parametrize("cfg_name,lv,groups", [("charades_vgg", 200, 40), ("tacos", 200, 24), ("qvhighlights", 75, 48)])
def test_benchmark_shapes_properties(cfg_name, lv, groups):
    """BASELINE.json configs[2] / [3] / [4] at their benchmark shapes (Lv = 200 VGG / C3D features, TwoMLP, QVH grouping), where the
    CPU oracle is too slow to run in a test: size-independent properties instead - packed rows == padded rows at the valid clips,
    chunking is invisible, a pair's scores do not depend on which other pairs share its batch when the lengths are uniform, and the
    decode of the result is well formed."""
    import mesm_b200
    from mesm_b200.ingest import clip_counts
    from oracle.config import CONFIGS
    from oracle.weights import make_inputs, make_neg_index, make_state_dict
    cfg = CONFIGS[cfg_name]
    g = torch.Generator().manual_seed(11)
    nc = torch.randint(1, 5 if cfg_name != "tacos" else 11, (groups,), generator=g).tolist()
    sd = make_state_dict(cfg, 5)
    inp = make_inputs(cfg, nc, 6, lv=lv, ragged_video=cfg_name != "qvhighlights")
    neg = make_neg_index(nc, 7)
    eng = mesm_b200.Engine(engine_cfg(cfg), chunk_pairs=64)
    eng.load_state_dict(sd)
    dev = eng.device
    args = (inp["video_feat"].to(dev), inp["video_mask"].to(dev), inp["words_feat"].to(dev), inp["num_clips"])
    vm = inp["video_mask"]
    a = eng.forward(*args, neg_index=neg.to(dev), want=("core", "aux", "rec"))
    b = eng.forward(*args, neg_index=neg.to(dev), want=("core", "aux", "rec"), video_len=clip_counts(vm))
    eng.set_chunk_pairs(17)
    c = eng.forward(*args, neg_index=neg.to(dev), want=("core",), video_len=clip_counts(vm))
    torch.cuda.synchronize()
    for k in ("pred_logits", "pred_spans", "aux_logits", "aux_spans", "recon_feat"):
        assert not torch.isnan(a[k]).any(), k
        assert rel_err(b[k], a[k]) < 1e-4, k
    for k in ("saliency_scores", "neg_saliency_scores"):
        assert rel_err(b[k], a[k], vm) < 1e-4, k
    assert rel_err(b["enhanced_video_feat"], a["enhanced_video_feat"], vm[..., None]) < 1e-4
    for k in ("pred_logits", "pred_spans"):
        assert rel_err(c[k], b[k]) < 1e-4, k
    assert rel_err(c["saliency_scores"], b["saliency_scores"], vm) < 1e-4
    # decode: ranked by score, windows inside [0, max_ts], kept set is a subset of the queries in rank order
    win, order, keep, cnt = mesm_b200.decode_nms(b["pred_logits"], b["pred_spans"], inp["duration"].to(dev), cfg.clip_len,
                                                 cfg.max_ts_val, 0.7, 10, 10)
    w = win.cpu()
    assert (w[:, :-1, 2] >= w[:, 1:, 2]).all() and (w[..., 0] >= 0).all() and (w[..., 1] <= cfg.max_ts_val).all()
    assert (w[..., 1] >= w[..., 0]).all() and (cnt.cpu() >= 1).all()
    for i in range(0, w.shape[0], 7):
        kp = keep[i, :int(cnt[i])].tolist()
        assert kp[0] == int(order[i, 0]) and len(set(kp)) == len(kp)
    if cfg_name == "qvhighlights":                  # uniform lengths: no cross-pair coupling -> a sub-batch reproduces its pairs
        n0 = sum(nc[:groups // 2])
        sub = eng.forward(args[0][:n0], args[1][:n0], args[2][:n0], inp["num_clips"][:groups // 2], want=("core",))
        assert rel_err(sub["pred_logits"], a["pred_logits"][:n0]) < 1e-4 and rel_err(sub["pred_spans"], a["pred_spans"][:n0]) < 1e-4


def test_forward_accepts_16bit_word_features():
    """Word features stored in 16 bits (bench.py --feature-dtype f16) are widened on the device: same bits as their fp32 upcast."""
    import mesm_b200
    cfg, sd, inp, neg, gold, meta = load_case(sorted(golden_cases())[0])
    eng = mesm_b200.Engine(engine_cfg(cfg), chunk_pairs=256)
    eng.load_state_dict(sd)
    dev = eng.device
    w16 = inp["words_feat"].half().to(dev)
    args = (inp["video_feat"].to(dev), inp["video_mask"].to(dev))
    a = eng.forward(*args, w16, inp["num_clips"], neg_index=neg.to(dev), want=("core",))
    b = eng.forward(*args, w16.float(), inp["num_clips"], neg_index=neg.to(dev), want=("core",))
    for k in ("pred_logits", "pred_spans", "saliency_scores"):
        assert torch.equal(a[k], b[k]), k
