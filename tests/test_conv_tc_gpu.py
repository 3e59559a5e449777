"""bf16 tcgen05 convolutions (throughput mode) vs a CPU fp32 reference on the same bf16-rounded operands.
Tolerance: outputs are rounded to bf16 (2^-9 relative per element) => norm-wise 4e-3; wgrad is fp32 out => 1e-3."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [  # B, H, C, K, R, stride, pad
    (2, 16, 64, 64, 3, 1, 1),
    (2, 16, 64, 128, 3, 2, 1),
    (2, 16, 64, 128, 1, 2, 0),
    (1, 7, 128, 256, 3, 1, 1),      # M = 49 < 128 (ragged tile)
    (3, 8, 256, 512, 3, 2, 1),
    (2, 32, 8, 64, 7, 2, 3),        # stem geometry, channels padded 3 -> 8, K = 392 (tail k-block)
    (1, 14, 512, 512, 3, 1, 1),
    (5, 12, 64, 64, 3, 1, 1),       # 128-pixel tiles straddle image boundaries (144 px / image)
    (2, 10, 128, 128, 3, 2, 1),
]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def desc(B, H, C, K, R, s, p):
    from primia_b200._lib import ConvDesc

    Ho = (H + 2 * p - R) // s + 1
    return ConvDesc(B, H, H, C, K, R, R, s, p, Ho, Ho), Ho


def bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("no_tma", ["0", "1"])  # 0: TMA im2col producer where eligible; 1: cp.async gather everywhere
@pytest.mark.parametrize("case", CASES)
def test_fwd_dgrad_wgrad_bf16(case, no_tma, monkeypatch):
    from primia_b200._lib import call, ptr, stream

    monkeypatch.setenv("PRIMIA_NO_TMA", no_tma)

    B, H, C, K, R, s, p = case
    g = torch.Generator().manual_seed(sum(case))
    d, Ho = desc(*case)
    x = bf(torch.randn(B, H, H, C, generator=g))           # NHWC
    w = bf(torch.randn(K, R, R, C, generator=g) * 0.1)     # KRSC
    dy = bf(torch.randn(B, Ho, Ho, K, generator=g))
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.conv2d(xr, wr, None, s, p)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    xd, wd_, dyd = x.to(DEV), w.to(DEV), dy.to(DEV)  # keep device tensors alive across the raw-pointer calls
    # forward
    y = torch.empty(B, Ho, Ho, K, dtype=torch.bfloat16, device=DEV)
    call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(xd), ptr(wd_), ptr(y), None, stream())
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 3, 1, 2), yr.detach()) < 4e-3, "fwd"
    # wgrad
    dw = torch.empty(K, R, R, C, dtype=torch.float32, device=DEV)
    call("pm_conv_wgrad_bf16", ctypes.byref(d), ptr(xd), ptr(dyd), ptr(dw), None, stream())
    torch.cuda.synchronize()
    assert rel(dw.permute(0, 3, 1, 2), wr.grad) < 1e-3, "wgrad"
    # dgrad (needs C % 64 == 0)
    if C % 64 == 0:
        wt = w.permute(3, 1, 2, 0).contiguous().to(DEV)  # [C][R][S][K]
        dx = torch.empty(B, H, H, C, dtype=torch.bfloat16, device=DEV)
        call("pm_conv_dgrad_bf16", ctypes.byref(d), ptr(dyd), ptr(wt), ptr(dx), 0, stream())
        torch.cuda.synchronize()
        assert rel(dx.float().permute(0, 3, 1, 2), xr.grad) < 4e-3, "dgrad"
        call("pm_conv_dgrad_bf16", ctypes.byref(d), ptr(dyd), ptr(wt), ptr(dx), 1, stream())
        torch.cuda.synchronize()
        assert rel(dx.float().permute(0, 3, 1, 2), 2 * xr.grad) < 8e-3, "dgrad accumulate"
