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
