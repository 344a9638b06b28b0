"""GPU: ragged host->device ingest (mesm_upload_clips / prepare_batch_input, drop-in of dataset/base.py:358-383) is
bit-identical to the reference's plain `.to(device)` of the zero-padded tensors."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, L, Dv, lens, seed=0):
    g = torch.Generator().manual_seed(seed)
    vf = torch.randn(B, L, Dv, generator=g)
    lens = torch.as_tensor(lens)
    mask = torch.arange(L)[None] < lens[:, None]
    vf = vf * mask[..., None]                          # the collate function zero-pads (utils/data_utils.py:66-82)
    return vf, mask


@pytest.mark.parametrize("Dv,L", [(2818, 194), (4098, 75), (7, 5), (256, 1)])
def test_upload_clips_bit_exact(Dv, L):
    import mesm_b200
    B = 9
    g = torch.Generator().manual_seed(Dv)
    lens = torch.randint(1, L + 1, (B,), generator=g).tolist()
    lens[0], lens[3], lens[4] = L, L, L                # full-length pairs -> merged copies across pair boundaries
    if L > 1:
        lens[-1] = 1
    vf, mask = _batch(B, L, Dv, lens)
    for pinned in (True, False):
        h, m = (vf.pin_memory(), mask.pin_memory()) if pinned else (vf, mask)
        out_f = torch.full((B, L, Dv), float("nan"), device="cuda")       # stale garbage in the staging buffer
        out_m = torch.zeros(B, L, dtype=torch.bool, device="cuda")
        f, mm, n = mesm_b200.upload_clips(h, m, out_feat=out_f, out_mask=out_m)
        torch.cuda.synchronize()
        assert f.data_ptr() == out_f.data_ptr()
        assert torch.equal(f.cpu(), vf) and torch.equal(mm.cpu(), mask)
        assert n == sum(lens) * Dv * 4 + B * L             # only valid rows (+ the mask) crossed the bus


def test_upload_clips_non_prefix_mask_and_empty_rows():
    import mesm_b200
    B, L, Dv = 4, 6, 10
    vf = torch.randn(B, L, Dv)
    mask = torch.tensor([[1, 1, 0, 1, 0, 0], [0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1], [1, 0, 0, 0, 0, 0]], dtype=torch.bool)
    f, mm, n = mesm_b200.upload_clips(vf, mask)
    torch.cuda.synchronize()
    assert torch.equal(f.cpu(), vf * mask[..., None])      # mask == 0 rows are zero on the device, valid rows exact
    assert torch.equal(mm.cpu(), mask)
    with pytest.raises(RuntimeError):
        mesm_b200.upload_clips(vf.cuda(), mask.cuda())


def test_prepare_batch_input_is_a_drop_in():
    import mesm_b200
    B, L, Dv = 5, 12, 34
    vf, mask = _batch(B, L, Dv, [12, 3, 7, 12, 1], seed=3)
    batch = dict(video_feat=vf.pin_memory(), video_mask=mask.pin_memory(), words_id=torch.randn(B, 4, 8),
                 words_weight=torch.ones(B, 4), num_clips=torch.tensor([2, 3]), duration=torch.tensor([10., 20, 30, 40, 50]),
                 moment=torch.tensor([[1., 2], [3, 9], [0, 30], [5, 6], [10, 40]]), qid=[1, 2, 3, 4, 5])
    ref = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in batch.items()}
    out = mesm_b200.prepare_batch_input(batch, "cuda", non_blocking=True)
    torch.cuda.synchronize()
    assert out is batch
    assert not out["words_weight"].is_cuda and out["qid"] == [1, 2, 3, 4, 5]            # dataset/base.py:360-361
    for k in ("video_feat", "video_mask", "words_id", "num_clips", "duration", "moment"):
        assert out[k].is_cuda and torch.equal(out[k].cpu(), ref[k]), k
    nm = ref["moment"] / ref["duration"].unsqueeze(1)                                   # dataset/base.py:379-383
    assert torch.equal(out["norm_moment"].cpu(), nm)
    cxw = torch.stack([(nm[:, 0] + nm[:, 1]) * 0.5, nm[:, 1] - nm[:, 0]], dim=-1)
    assert torch.allclose(out["norm_span"].cpu(), cxw, atol=1e-7)
    assert mesm_b200.prepare_batch_input.last_h2d_bytes < sum(v.numel() * v.element_size() for k, v in ref.items()
                                                              if torch.is_tensor(v) and k != "words_weight")


def test_upload_clips_shared_group_video():
    """charades / tacos collate replicates the group's video (dataset/base.py:307-309): with num_clips only the first pair
    of every group crosses the bus; the other pairs' rows are left untouched."""
    import mesm_b200
    nc = [2, 1, 3, 1]
    B, L, Dv = sum(nc), 9, 18
    glen = [9, 4, 6, 9]
    vids = [torch.randn(n, Dv) for n in glen]
    vf, mask = torch.zeros(B, L, Dv), torch.zeros(B, L, dtype=torch.bool)
    b, firsts = 0, []
    for g, n in enumerate(nc):
        firsts.append(b)
        for _ in range(n):
            vf[b, :glen[g]] = vids[g]; mask[b, :glen[g]] = True; b += 1
    out_f = torch.full((B, L, Dv), 7.0, device="cuda")
    f, mm, nbytes = mesm_b200.upload_clips(vf.pin_memory(), mask.pin_memory(), out_feat=out_f, num_clips=torch.tensor(nc))
    torch.cuda.synchronize()
    fc = f.cpu()
    assert torch.equal(mm.cpu(), mask)
    for b in range(B):
        if b in firsts:
            assert torch.equal(fc[b], vf[b])
        else:                                        # untouched valid rows, zero-filled pad rows
            assert torch.equal(fc[b][mask[b]], torch.full_like(vf[b][mask[b]], 7.0)) and float(fc[b][~mask[b]].abs().sum()) == 0.0
    assert nbytes == sum(glen) * Dv * 4 + B * L
    bad = vf.clone(); bad[1, 0, 0] += 1.0            # pair 1 no longer shares pair 0's video
    with pytest.raises(ValueError):
        mesm_b200.upload_clips(bad, mask, num_clips=nc)


@pytest.mark.parametrize("name", ["tiny_ragged", "charades_csf_ragged", "tacos_l96", "tiny_twomlp"])
def test_eval_loop_drop_in_with_shared_upload(name):
    """eval.py:62-63 with the two drop-ins: prepare_batch_input(batch, device, non_blocking) then model(**batch, ...); the
    video of a query group is uploaded once, the forward runs on packed rows - results = the reference's golden outputs."""
    import mesm_b200
    from mesm_b200.model import build_model
    from tests.helpers import engine_cfg, load_case, rel_err
    cfg, sd, inp, neg, gold, meta = load_case(name)
    model = build_model(engine_cfg(cfg))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    B, L, Dv = inp["video_feat"].shape
    batch = dict(video_feat=inp["video_feat"].pin_memory(), video_mask=inp["video_mask"].pin_memory(), words_id=inp["words_feat"],
                 words_mask=None, words_weight=torch.ones(B, inp["words_feat"].shape[1]), num_clips=inp["num_clips"],
                 duration=inp["duration"], qid=list(range(B)))
    stale = {"video_feat": torch.full((B, L, Dv), float("nan"), device="cuda")}       # a reused staging buffer
    mesm_b200.prepare_batch_input(batch, "cuda", non_blocking=True, out=stale, shared_group_video=True)
    assert batch["shared_group_video"] is True and not batch["video_len"].is_cuda
    out = model(**batch, dataset_name=cfg.dataset_name, is_training=False, neg_index=neg.cuda())
    torch.cuda.synchronize()
    vm = inp["video_mask"]
    assert rel_err(out["pred_logits"], gold["pred_logits"]) <= 1e-3
    assert rel_err(out["pred_spans"], gold["pred_spans"]) <= 1e-3
    assert rel_err(out["saliency_scores"], gold["saliency_scores"], vm) <= 1e-3
    assert rel_err(out["neg_saliency_scores"], gold["neg_saliency_scores"], vm) <= 1e-3
    assert rel_err(out["recon_feat"], gold["recon_feat"]) <= 1e-3
    assert rel_err(out["projed_video_feat"][:, 0], gold["projed_video_row0"]) <= 1e-3
    assert not torch.isnan(out["projed_video_feat"]).any() and not torch.isnan(out["enhanced_video_feat"]).any()


def test_pinned_source_may_be_reused_right_after_the_call():
    """the raw async copies must not outlive their host source.  The caller drops / overwrites its pinned tensors
    immediately after prepare_batch_input returns (what the DataLoader's pin-memory thread does with recycled blocks)."""
    import mesm_b200
    from mesm_b200 import ingest
    B, L, Dv = 64, 194, 2818
    lens = [L] * B
    busy = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    outs, refs = [], []
    for it in range(6):
        vf, mask = _batch(B, L, Dv, lens, seed=100 + it)
        refs.append(vf.clone())
        for _ in range(8):
            busy.add_(1)                                   # keep the stream busy so the copy is still queued on return
        batch = dict(video_feat=vf.pin_memory(), video_mask=mask.pin_memory(), num_clips=torch.ones(B, dtype=torch.long))
        out = mesm_b200.prepare_batch_input(batch, "cuda", non_blocking=True)
        outs.append(out["video_feat"])
        del batch                  # the only outside reference: the next pin_memory() would recycle the block at once
    torch.cuda.synchronize()
    for it, (o, r) in enumerate(zip(outs, refs)):
        assert torch.equal(o.cpu(), r), it                 # dropped sources stayed alive until the DMA had read them
    assert len(ingest._inflight) <= 6
    # pageable sources and non_blocking=False complete before the call returns
    vf, mask = _batch(4, 8, 16, [8, 3, 8, 1], seed=9)
    out = mesm_b200.prepare_batch_input(dict(video_feat=vf.clone(), video_mask=mask), "cuda", non_blocking=False)
    assert torch.cuda.current_stream().query()
    assert torch.equal(out["video_feat"].cpu(), vf)


@pytest.mark.parametrize("shared", [False, True])
def test_fp16_feature_storage_upload_and_forward(shared):
    """16-bit feature storage: fp16 host features cross PCIe as they are (half the bytes), bit-exact on the device,
    and the forward on them equals the reference fed the same values upcast to fp32 (golden charades_csf_ragged_f16)."""
    import mesm_b200
    from mesm_b200.model import build_model
    from tests.helpers import engine_cfg, load_case, rel_err
    cfg, sd, inp, neg, gold, meta = load_case("charades_csf_ragged_f16")
    model = build_model(engine_cfg(cfg))
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    B, L, Dv = inp["video_feat"].shape
    h16 = inp["video_feat"].half()
    assert torch.equal(h16.float(), inp["video_feat"])                    # the fixture's features are fp16-representable
    f, mm, n = mesm_b200.upload_clips(h16.pin_memory(), inp["video_mask"].pin_memory(), num_clips=inp["num_clips"] if shared else None)
    torch.cuda.synchronize()
    assert f.dtype == torch.float16
    rows = int(inp["video_mask"].sum()) if not shared else int(inp["video_mask"][torch.cumsum(inp["num_clips"], 0) - inp["num_clips"]].sum())
    assert n == rows * Dv * 2 + B * L
    if not shared:
        assert torch.equal(f.cpu(), h16)
    batch = dict(video_feat=h16.pin_memory(), video_mask=inp["video_mask"].pin_memory(), words_id=inp["words_feat"], words_mask=None,
                 words_weight=torch.ones(B, inp["words_feat"].shape[1]), num_clips=inp["num_clips"], duration=inp["duration"])
    mesm_b200.prepare_batch_input(batch, "cuda", non_blocking=True, shared_group_video=shared)
    assert batch["video_feat"].dtype == torch.float16 and mesm_b200.prepare_batch_input.last_h2d_bytes < B * L * Dv * 2 + 10 ** 6
    out = model(**batch, dataset_name=cfg.dataset_name, is_training=False, neg_index=neg.cuda())
    torch.cuda.synchronize()
    vm = inp["video_mask"]
    for k in ("pred_logits", "pred_spans", "recon_feat"):
        assert rel_err(out[k], gold[k]) <= 1e-3, k
    for k in ("saliency_scores", "neg_saliency_scores"):
        assert rel_err(out[k], gold[k], vm) <= 1e-3, k
    assert rel_err(out["projed_video_feat"][:, 0], gold["projed_video_row0"]) <= 1e-3
    assert model._eng.last_feature_bytes == rows * Dv * 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs a second GPU to relay through")
def test_relayed_upload_is_bit_identical_to_the_direct_one():
    """mesm_b200.relay: part of the videos reaches GPU 0 through GPU 1's host link and a peer copy; same bits, pad rows zero."""
    import mesm_b200
    from mesm_b200.relay import IngestRelay
    g = torch.Generator().manual_seed(11)
    nc = [1, 3, 2, 1, 4, 2, 3]
    B, L, Dv = sum(nc), 23, 130
    lens_g = torch.randint(5, L + 1, (len(nc),), generator=g)
    lens = torch.repeat_interleave(lens_g, torch.tensor(nc))
    mask = torch.arange(L)[None] < lens[:, None]
    vg = torch.randn(len(nc), L, Dv, generator=g).half()
    vf = (torch.repeat_interleave(vg, torch.tensor(nc), dim=0) * mask[..., None]).contiguous().pin_memory()
    mask = mask.pin_memory()
    dev = torch.device("cuda", 0)
    ref_f = torch.full((B, L, Dv), 7.0, dtype=torch.float16, device=dev)
    ref_m = torch.zeros(B, L, dtype=torch.bool, device=dev)
    mesm_b200.upload_clips(vf, mask, out_feat=ref_f, out_mask=ref_m, num_clips=nc)
    relay = IngestRelay(dev, torch.device("cuda", 1), 0.5, 1 << 20, nbuf=2)
    first = torch.cumsum(torch.tensor([0] + nc[:-1]), 0)
    for rep in range(5):                                  # more calls than staging buffers
        relay.batched = rep % 2 == 0                      # batched submission / one copy per video
        out_f = torch.full((B, L, Dv), 7.0, dtype=torch.float16, device=dev)
        out_m = torch.zeros(B, L, dtype=torch.bool, device=dev)
        with torch.cuda.device(dev):
            batch = mesm_b200.prepare_batch_input(dict(video_feat=vf, video_mask=mask, num_clips=torch.tensor(nc)), dev, non_blocking=True,
                                                  out=dict(video_feat=out_f, video_mask=out_m), shared_group_video=True, relay=relay)
        torch.cuda.synchronize(dev)
        assert relay.broken is None and relay.last_relayed_bytes > 0
        assert torch.equal(out_m, ref_m)
        assert torch.equal(out_f[first], ref_f[first])    # the rows the forward reads (first pair of every group)
        assert batch["video_len"].tolist() == lens.tolist()
