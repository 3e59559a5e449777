"""Per-geometry timing of the bf16 tensor-core convolutions (ResNet-18 layers at batch B): halo-strip vs im2col-TMA.
python scripts/conv_bench.py [B]   -> one line per (layer, op, variant): us, TFLOP/s."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from primia_b200._lib import ConvDesc, call, ptr, stream  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
DEV = "cuda:0"
LAYERS = [("layer1 56x56 64->64", 56, 64, 64), ("layer2 28x28 128->128", 28, 128, 128), ("layer3 14x14 256->256", 14, 256, 256),
          ("layer4 7x7 512->512", 7, 512, 512)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, n=10, inner=8):
    """`inner` back-to-back launches per event pair (the host-side tensor-map encoding of launch i+1 overlaps the execution
    of launch i, as it does inside a CUDA graph), L2 evicted before each group."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()  # evict L2
        torch.cuda._sleep(2_000_000)  # ~1 ms spin: the host queues all `inner` launches before the GPU reaches the bracket
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(inner):
            fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / (n * inner) * 1e3


for name, H, C, K in LAYERS:
    d = ConvDesc(B, H, H, C, K, 3, 3, 1, 1, H, H)
    x = torch.randn(B, H, H, C, device=DEV).bfloat16()
    w = (torch.randn(K, 3, 3, C, device=DEV) * 0.05).bfloat16()
    wt = w.permute(3, 1, 2, 0).contiguous()
    dy = torch.randn(B, H, H, K, device=DEV).bfloat16()
    y = torch.empty(B, H, H, K, device=DEV, dtype=torch.bfloat16)
    dx = torch.empty(B, H, H, C, device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(2 * K, dtype=torch.float64, device=DEV)
    flops = 2.0 * B * H * H * C * K * 9
    res = {}
    for variant in ("halo", "tma"):
        os.environ["PRIMIA_NO_HALO"] = "0" if variant == "halo" else "1"
        t_f = timeit(lambda: call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(x), ptr(w), ptr(y), ptr(stats), stream()))
        res[variant] = y.float().clone()
        t_d = timeit(lambda: call("pm_conv_dgrad_bf16", ctypes.byref(d), ptr(dy), ptr(wt), ptr(dx), 0, stream()))
        print(f"{name:24s} {variant:5s} fwd {t_f:7.1f} us {flops / t_f / 1e6:7.1f} TF/s | dgrad {t_d:7.1f} us {flops / t_d / 1e6:7.1f} TF/s")
    # weight gradient: im2col-TMA kernel vs the opt-in halo-strip kernel (C, K multiples of 128 only)
    dwb = torch.zeros(K, 3, 3, C, device=DEV, dtype=torch.float32)
    for variant in ("tma", "halo"):
        os.environ["PRIMIA_HALO_WGRAD"] = "1" if variant == "halo" else "0"
        t_w = timeit(lambda: call("pm_conv_wgrad_bf16", ctypes.byref(d), ptr(x), ptr(dy), ptr(dwb), None, stream()))
        print(f"{name:24s} {variant:5s} wgrad {t_w:7.1f} us {flops / t_w / 1e6:7.1f} TF/s")
    os.environ["PRIMIA_HALO_WGRAD"] = "0"
    # per-CTA cycle counters of one profiled halo launch (fwd, then dgrad)
    import numpy as np
    os.environ["PRIMIA_NO_HALO"] = "0"
    for op in ("fwd", "dgrad"):
        call("pm_halo_prof", 1, None)
        if op == "fwd":
            call("pm_conv_fwd_bf16", ctypes.byref(d), ptr(x), ptr(w), ptr(y), ptr(stats), stream())
        else:
            call("pm_conv_dgrad_bf16", ctypes.byref(d), ptr(dy), ptr(wt), ptr(dx), 0, stream())
        torch.cuda.synchronize()
        buf = np.zeros(160 * 8 + 1 + 512, dtype=np.int64)
        call("pm_halo_prof", 0, buf.ctypes.data_as(ctypes.c_void_p))
        pr = buf[:160 * 8].reshape(160, 8)[:148]
        pr = pr[pr[:, 0] > 0]
        ep = buf[160 * 8 + 1:160 * 8 + 1 + 128 * 4].reshape(128, 4)
        ep = ep[ep[:, 3] > 0]
        if len(ep):
            print(f"{'':24s} {op} epilogue (warp 4) per chunk: tcgen05.ld+wait {ep[:, 0].sum() / ep[:, 3].sum():.0f} clk, stage+store(+stats) {ep[:, 1].sum() / ep[:, 3].sum():.0f} clk, chunks/CTA {ep[:, 3].mean():.1f}")
        ntr = int(buf[160 * 8])
        tr = buf[160 * 8 + 1:160 * 8 + 1 + 2 * ntr].reshape(ntr, 2)
        t0 = tr[tr[:, 0] == 0][0, 1] if (tr[:, 0] == 0).any() else tr[:, 1].min()
        ev = sorted((int(t - t0), int(c)) for c, t in tr if c != 0)
        print(f"{'':24s} {op} kernel clk min/max over CTAs {pr[:, 0].min()}/{pr[:, 0].max()}")
        names = ["kernel", "mma:wait_strip", "mma:wait_w", "mma:wait_tmem", "epi:wait_acc", "epi:busy", "sprod:wait", "wprod:wait"]
        print(f"{'':24s} {op} clk/CTA (mean over {len(pr)} CTAs): " + "  ".join(f"{n}={int(pr[:, i].mean())}" for i, n in enumerate(names)))
    diff = (res["halo"] - res["tma"]).norm() / res["tma"].norm()
    print(f"{'':24s} halo vs tma fwd rel diff {diff.item():.2e}")
