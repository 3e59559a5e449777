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
    (2, 56, 64, 64, 3, 1, 1),       # layer1 geometry (resident-weight halo kernel, 256-pixel tiles across image boundaries)
    (3, 28, 128, 128, 3, 1, 1),     # layer2
    (3, 14, 256, 256, 3, 1, 1),     # layer3 (two n-tiles share a strip)
    (5, 7, 512, 512, 3, 1, 1),      # layer4 (8-pixel padded rows, 8 channel blocks)
    (70, 14, 128, 128, 3, 1, 1),    # enough tiles for the 256-pixel variant with streamed weights, >1 tile per CTA
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


# "halo": halo-strip kernel for 3x3/s1, persistent stride-2 dgrad, im2col-TMA producer elsewhere (the default); "tma": im2col-TMA producer wherever
# eligible; "cpasync": cp.async gather everywhere
@pytest.mark.parametrize("variant", ["halo", "tma", "cpasync"])
@pytest.mark.parametrize("case", CASES)
def test_fwd_dgrad_wgrad_bf16(case, variant, monkeypatch):
    from primia_b200._lib import call, ptr, stream

    monkeypatch.setenv("PRIMIA_NO_TMA", "1" if variant == "cpasync" else "0")
    monkeypatch.setenv("PRIMIA_NO_HALO", "0" if variant == "halo" else "1")
    monkeypatch.setenv("PRIMIA_NO_S2P", "0" if variant == "halo" else "1")  # persistent stride-2 dgrad rides with "halo"

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
    stats = torch.zeros(2 * K, dtype=torch.float64, device=DEV)
    call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(xd), ptr(wd_), ptr(y), ptr(stats), stream())
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 3, 1, 2), yr.detach()) < 4e-3, "fwd"
    # BatchNorm statistics produced by the conv (fused epilogue on the TMA path, separate pass otherwise):
    # per-channel sum and sum of squares of the stored bf16 outputs
    yf = y.double().reshape(-1, K)
    assert rel(stats[:K], yf.sum(0)) < 1e-6 and rel(stats[K:], (yf * yf).sum(0)) < 1e-6, "fused BN statistics"
    # wgrad
    dw = torch.zeros(K, R, R, C, dtype=torch.float32, device=DEV)  # wgrad accumulates
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


def test_stem_im2col_plus_dense_gemm_matches_conv7x7():
    """bf16 stem: NCHW fp32 input -> bf16 im2col [B,Ho,Wo,192] -> 1x1 tensor-core GEMM == conv 7x7 / stride 2 / pad 3."""
    from primia_b200._lib import ConvDesc, call, ptr, stream

    g = torch.Generator().manual_seed(3)
    B, H = 3, 40
    x = torch.randn(B, 3, H, H, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    Ho = (H + 6 - 7) // 2 + 1
    xd = x.to(DEV)
    im = torch.empty(B, Ho, Ho, 192, dtype=torch.bfloat16, device=DEV)
    call("pm_im2col_stem_bf16", ptr(xd), B, 3, H, H, 7, 2, 3, 192, ptr(im), stream())
    # reference im2col in the same k order (r, s, c)
    cols = F.unfold(x.bfloat16().float(), 7, padding=3, stride=2)  # [B, 3*49, L] ordered (c, r, s)
    cols = cols.view(B, 3, 49, Ho, Ho).permute(0, 3, 4, 2, 1).reshape(B, Ho, Ho, 147)
    assert torch.equal(im[..., :147].float().cpu(), cols)
    assert torch.count_nonzero(im[..., 147:]) == 0
    wk = torch.zeros(64, 1, 1, 192)
    wk[:, 0, 0, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)
    wk = wk.bfloat16().to(DEV)
    d = ConvDesc(B, Ho, Ho, 192, 64, 1, 1, 1, 0, Ho, Ho)
    y = torch.empty(B, Ho, Ho, 64, dtype=torch.bfloat16, device=DEV)
    call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(im), ptr(wk), ptr(y), None, stream())
    torch.cuda.synchronize()
    ref = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), None, 2, 3)
    assert rel(y.float().permute(0, 3, 1, 2), ref) < 4e-3


def test_halo_conv_non_square_and_odd_sizes():
    """Halo-strip kernel on H != W, odd widths and a batch that leaves a ragged last tile; accumulate on top of a gradient."""
    from primia_b200._lib import ConvDesc, call, ptr, stream

    g = torch.Generator().manual_seed(11)
    for (B, H, W, C, K) in [(3, 9, 13, 64, 64), (2, 5, 31, 128, 64), (7, 3, 3, 64, 128), (1, 1, 1, 64, 64)]:
        d = ConvDesc(B, H, W, C, K, 3, 3, 1, 1, H, W)
        x = bf(torch.randn(B, H, W, C, generator=g))
        w = bf(torch.randn(K, 3, 3, C, generator=g) * 0.1)
        dy = bf(torch.randn(B, H, W, K, generator=g))
        xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        wr = w.float().permute(0, 3, 1, 2)
        yr = F.conv2d(xr, wr, None, 1, 1)
        yr.backward(dy.float().permute(0, 3, 1, 2))
        xd, wd_, dyd = x.to(DEV), w.to(DEV), dy.to(DEV)
        y = torch.full((B, H, W, K), 7.0, dtype=torch.bfloat16, device=DEV)
        stats = torch.zeros(2 * K, dtype=torch.float64, device=DEV)
        call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(xd), ptr(wd_), ptr(y), ptr(stats), stream())
        torch.cuda.synchronize()
        assert rel(y.float().permute(0, 3, 1, 2), yr.detach()) < 4e-3, (B, H, W, C, K)
        yf = y.double().reshape(-1, K)
        assert rel(stats[:K], yf.sum(0)) < 1e-6 and rel(stats[K:], (yf * yf).sum(0)) < 1e-6
        wt = w.permute(3, 1, 2, 0).contiguous().to(DEV)
        dx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=DEV)
        call("pm_conv_dgrad_bf16", ctypes.byref(d), ptr(dyd), ptr(wt), ptr(dx), 0, stream())
        torch.cuda.synchronize()
        assert rel(dx.float().permute(0, 3, 1, 2), xr.grad) < 4e-3, (B, H, W, C, K)


@pytest.mark.parametrize("B,H,W", [(3, 40, 40), (2, 224, 224), (1, 18, 300), (5, 8, 12)])
def test_direct_stem_conv_fwd_and_wgrad(B, H, W):
    """conv 7x7 / s2 / p3, 3 -> 64 straight from the fp32 NCHW batch (no im2col matrix): forward + fused BN statistics and
    the weight gradient vs F.conv2d on the same bf16-rounded operands.  W = 300 gives two 128-pixel tiles per output row."""
    from primia_b200._lib import call, ptr, stream

    g = torch.Generator().manual_seed(B * 1000 + H + W)
    x = torch.randn(B, 3, H, W, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    dy = bf(torch.randn(B, Ho, Wo, 64, generator=g))
    xr = x.bfloat16().float().requires_grad_(False)
    wr = w.bfloat16().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, 2, 3)
    yr.backward(dy.float().permute(0, 3, 1, 2))
    xd = x.to(DEV)
    w_krsc = w.permute(0, 2, 3, 1).contiguous().to(DEV)           # fp32 master, KRSC
    w192 = torch.empty(64, 192, dtype=torch.bfloat16, device=DEV)
    call("pm_stem_prep_w_bf16", ptr(w_krsc), ptr(w192), stream())
    y = torch.full((B, Ho, Wo, 64), 3.0, dtype=torch.bfloat16, device=DEV)
    stats = torch.zeros(128, dtype=torch.float64, device=DEV)
    call("pm_stem_conv_fwd_bf16", ptr(xd), ptr(w192), B, H, W, ptr(y), ptr(stats), stream())
    torch.cuda.synchronize()
    assert rel(y.float().permute(0, 3, 1, 2), yr.detach()) < 4e-3, "stem fwd"
    yf = y.double().reshape(-1, 64)
    assert rel(stats[:64], yf.sum(0)) < 1e-6 and rel(stats[64:], (yf * yf).sum(0)) < 1e-6, "fused BN statistics"
    dyd = dy.to(DEV)
    dw = torch.zeros(64, 7, 7, 3, dtype=torch.float32, device=DEV)
    call("pm_stem_conv_wgrad_bf16", ptr(xd), ptr(dyd), B, H, W, ptr(dw), stream())
    torch.cuda.synchronize()
    assert rel(dw.permute(0, 3, 1, 2), wr.grad) < 1e-3, "stem wgrad"


@pytest.mark.parametrize("B,H,W,C,K", [(3, 28, 28, 128, 128), (5, 14, 14, 256, 256), (9, 7, 7, 512, 256), (2, 9, 13, 128, 128),
                                       (64, 14, 14, 256, 256)])
def test_halo_wgrad(B, H, W, C, K):
    """halo-strip weight gradient (wgrad_halo.cu): the default kernel for 3x3 / stride-1 layers with 128-multiple channels"""
    from primia_b200._lib import ConvDesc, call, ptr, stream

    g = torch.Generator().manual_seed(B + H + W + C + K)
    d = ConvDesc(B, H, W, C, K, 3, 3, 1, 1, H, W)
    x = bf(torch.randn(B, H, W, C, generator=g))
    dy = bf(torch.randn(B, H, W, K, generator=g))
    xr = x.float().permute(0, 3, 1, 2)
    wr = torch.zeros(K, C, 3, 3, requires_grad=True)
    F.conv2d(xr, wr, None, 1, 1).backward(dy.float().permute(0, 3, 1, 2))
    xd, dyd = x.to(DEV), dy.to(DEV)
    dw = torch.zeros(K, 3, 3, C, dtype=torch.float32, device=DEV)
    call("pm_conv_wgrad_bf16", ctypes.byref(d), ptr(xd), ptr(dyd), ptr(dw), None, stream())
    torch.cuda.synchronize()
    assert rel(dw.permute(0, 3, 1, 2), wr.grad) < 1e-3, (B, H, W, C, K, rel(dw.permute(0, 3, 1, 2), wr.grad))
