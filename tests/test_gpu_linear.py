"""GPU: the fused linear kernels (fp32 SIMT and tcgen05 split-bf16) against a float64 torch reference of the same op."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, use_tc, act=0, bias=True, res=False, ln=False, pos=False, fold=False, lda=None, seed=0, scale=1.0, f16_inputs=False):
    from mesm_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(seed)
    lda = lda or K
    Afull = torch.randn(M, lda, device="cuda", generator=g)
    if f16_inputs:
        Afull = Afull.half().float()           # values a 16-bit feature store holds exactly
    A = Afull[:, :K]
    P = torch.randn(M, lda, device="cuda", generator=g) if pos else None
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    R = torch.randn(M, N, device="cuda", generator=g) if res else None
    lg = (1 + 0.1 * torch.randn(N, device="cuda", generator=g)) if ln else None
    lb = (0.1 * torch.randn(N, device="cuda", generator=g)) if ln else None
    slope = torch.tensor([0.25], device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda")
    pre = torch.full((M, N), float("nan"), device="cuda") if ln else None
    x = A.double() + (P[:, :K].double() if pos else 0)
    rowstat = None
    if fold:                       # LayerNorm(x) . W^T folded: the kernel receives raw x, (mean, rstd) and colsum(W)
        mu, var = x.mean(1), x.var(1, unbiased=False)
        rowstat = torch.stack([mu, 1 / torch.sqrt(var + 1e-5)], 1).float().contiguous()
        x = (x - mu[:, None]) / torch.sqrt(var + 1e-5)[:, None]
    y = x @ W.double().t()
    if bias:
        y = y + b.double()
    y = y * scale
    if act == 1:
        y = y.clamp_min(0)
    elif act == 2:
        y = torch.where(y >= 0, y, 0.25 * y)
    if res:
        y = y + R.double()
    y_pre = y
    if ln:
        y = torch.nn.functional.layer_norm(y, (N,), lg.double(), lb.double(), 1e-5)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    rc = lib.mesm_debug_linear(p(Afull), p(P), p(W), p(b), p(R), p(lg), p(lb), p(rowstat), p(slope), M, N, K, lda, act,
                               float(scale), p(out), p(pre), int(use_tc), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, lib.mesm_last_error(None)
    err = float((out.double() - y).abs().max() / y.abs().max())
    if ln:
        err = max(err, float((pre.double() - y_pre).abs().max() / y_pre.abs().max()))
    return err


CASES = [
    dict(M=128, N=256, K=64),
    dict(M=128, N=256, K=256, act=1),
    dict(M=300, N=256, K=256, res=True, ln=True),
    dict(M=1000, N=1024, K=256, act=2),
    dict(M=777, N=256, K=1024, res=True, ln=True),
    dict(M=513, N=512, K=256, pos=True, scale=0.17677669),
    dict(M=400, N=256, K=2818, fold=True, act=1, ln=True),
    dict(M=256, N=256, K=300, fold=True, act=1, ln=True),
    dict(M=260, N=768, K=256, lda=260),
    dict(M=129, N=200, K=130),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_simt_linear(case):
    assert _run(use_tc=0, **case) < 2e-6 * (30 if case.get("fold") else 1)


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_tcgen05_linear(case):
    assert _run(use_tc=1, **case) < 5e-5


TMA_CASES = [
    dict(M=128, N=256, K=64),
    dict(M=300, N=256, K=256, res=True, ln=True),
    dict(M=1000, N=1024, K=256, act=2),
    dict(M=777, N=256, K=1024, res=True, ln=True),
    dict(M=513, N=512, K=256, scale=0.17677669),
    dict(M=400, N=256, K=2818, fold=True, act=1, ln=True),
    dict(M=256, N=256, K=300, fold=True, act=1, ln=True),
    dict(M=40000, N=256, K=256, res=True, ln=True),          # persistent loop: ~2 tiles per SM, both accumulator buffers
    dict(M=19, N=256, K=256, act=1),
]


@pytest.mark.parametrize("mode", [2, 3], ids=["fp32_out", "planes_out"])
@pytest.mark.parametrize("case", TMA_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_tma_linear(case, mode):
    """linear_tma_kernel (A and W as pre-split bf16 hi/lo planes through the TMA engine) vs float64; mode 3 stores the result as
    planes (hi + lo carries 16 mantissa bits: 2^-17 relative)."""
    assert _run(use_tc=mode, **case) < (5e-5 if mode == 2 else 6e-5)


@pytest.mark.parametrize("case", [dict(M=400, N=256, K=2818, fold=True, act=1, ln=True), dict(M=1000, N=256, K=4098, fold=True, act=1, ln=True),
                                  dict(M=130, N=256, K=256)], ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_tma_linear_fp16_features(case):
    """One exact fp16 A plane (16-bit stored clip features) against fp16 hi/lo weights: two MMAs per product."""
    assert _run(use_tc=4, f16_inputs=True, **case) < 5e-5
