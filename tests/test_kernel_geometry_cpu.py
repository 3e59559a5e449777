"""CPU checks of the index algebra the tensor-core conv kernels are built on (no GPU, no CUDA library).

The kernels themselves are parity-tested on the GPU (tests/test_conv_tc_gpu.py).  What is checked here is the geometry each
of them relies on, restated in a few lines of torch so that the invariants are executable documentation:

* conv_halo.cu / wgrad_halo.cu -- the virtual pixel grid: rows stored as [0, x_0 .. x_{W-1}] (Wp = W+1), images as
  [zero row, row_0 .. row_{H-1}] (Hp = H+1); tap (r, s) of output q is the element q + r*Wp + s of the linearised padded
  tensor; the strip of a tile needs (o + MT + 2*Wp + 1)//Wp + 1 padded rows.
* conv_stem.cu -- k order (r*3 + c)*8 + s, box start at input pixel 2*ow0 - 4, tap s of output pixel m at patch pixel 2m+1+s.
* conv_s2.cu -- the parity-class decomposition of the stride-2 data gradient (which taps feed class (a, b), at which offsets).
"""
import torch
import torch.nn.functional as F


def padded_linear(x):
    """x: [B, H, W, C] -> [B*(H+1)*(W+1) + slack, C]: what the per-row TMA boxes of conv_halo.cu lay down in shared memory."""
    B, H, W, C = x.shape
    p = torch.zeros(B, H + 1, W + 1, C, dtype=x.dtype)
    p[:, 1:, 1:, :] = x  # padded row hh <-> image row hh-1, padded column ww <-> image column ww-1
    flat = p.reshape(-1, C)
    return torch.cat([flat, torch.zeros(2 * (W + 1) + 2, C, dtype=x.dtype)])  # images past the end are zero-filled too


def test_halo_virtual_grid_reproduces_conv3x3():
    g = torch.Generator().manual_seed(0)
    for (B, H, W, C, K) in [(2, 5, 7, 3, 4), (3, 4, 4, 2, 2), (1, 1, 1, 2, 3)]:
        x = torch.randn(B, H, W, C, generator=g, dtype=torch.float64)
        w = torch.randn(K, 3, 3, C, generator=g, dtype=torch.float64)  # KRSC
        ref = F.conv2d(x.permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), None, 1, 1).permute(0, 2, 3, 1)
        Wp, Hp = W + 1, H + 1
        flat = padded_linear(x)
        V = B * Hp * Wp
        out = torch.zeros(V, K, dtype=torch.float64)
        q = torch.arange(V)
        for r in range(3):
            for s in range(3):
                out += flat[q + r * Wp + s] @ w[:, r, s, :].T  # a constant row shift per tap
        out = out.view(B, Hp, Wp, K)
        assert torch.allclose(out[:, :H, :W, :], ref, atol=1e-12)  # virtual pixels with h == H or w == W are discarded


def test_halo_strip_row_count_covers_every_tile_start():
    for Wp in (8, 15, 29, 57, 14):
        for MT in (128, 256):
            for q0 in range(0, 6 * Wp * MT, MT):
                R0, o = divmod(q0, Wp)
                nrows = (o + MT + 2 * Wp + 1) // Wp + 1
                last_needed = q0 + MT - 1 + 2 * Wp + 2
                assert R0 * Wp <= q0 and (R0 + nrows) * Wp > last_needed
                assert nrows <= (Wp - 1 + MT + 2 * Wp + 1) // Wp + 1  # the stage size the host allocates


def test_halo_wgrad_virtual_grid():
    """wgrad_halo.cu: dw[n][r][s][c] = sum over VIRTUAL pixels q of dy_v[q][n] * x_padded[q + r*Wp + s][c], with dy_v zero at
    the virtual pixels (w == W or h == H)."""
    g = torch.Generator().manual_seed(1)
    B, H, W, C, K = 2, 4, 5, 3, 2
    x = torch.randn(B, H, W, C, generator=g, dtype=torch.float64)
    dy = torch.randn(B, H, W, K, generator=g, dtype=torch.float64)
    wz = torch.zeros(K, C, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.permute(0, 3, 1, 2), wz, None, 1, 1).backward(dy.permute(0, 3, 1, 2))
    Wp, Hp = W + 1, H + 1
    flat = padded_linear(x)
    dyv = torch.zeros(B, Hp, Wp, K, dtype=torch.float64)
    dyv[:, :H, :W, :] = dy  # virtual row h <-> image row h, virtual column w <-> image column w
    dyv = dyv.reshape(-1, K)
    q = torch.arange(B * Hp * Wp)
    for r in range(3):
        for s in range(3):
            dw = dyv.T @ flat[q + r * Wp + s]  # [K, C]
            assert torch.allclose(dw, wz.grad[:, :, r, s], atol=1e-12)


def test_stem_k_order_and_patch_indexing():
    g = torch.Generator().manual_seed(2)
    B, H, W = 2, 12, 20
    x = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(5, 3, 7, 7, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, None, 2, 3)
    Ho, Wo = ref.shape[2:]
    # weights in k order (r*3 + c)*8 + s, s == 7 and k >= 168 zero  (stem_prep_w_kernel)
    wk = torch.zeros(5, 192, dtype=torch.float64)
    for r in range(7):
        for c in range(3):
            for s in range(7):
                wk[:, (r * 3 + c) * 8 + s] = w[:, c, r, s]
    xp = F.pad(x, (4, 8, 3, 3))  # zero fill outside the image; patch pixel j <-> input pixel 2*ow0 - 4 + j with ow0 = 0
    out = torch.zeros(B, 5, Ho, Wo, dtype=torch.float64)
    for oh in range(Ho):
        for m in range(Wo):
            a = torch.zeros(B, 192, dtype=torch.float64)
            for r in range(7):
                for c in range(3):
                    # chunk (r, c) of pixel m: patch pixels 2m+1 .. 2m+8 of input row 2*oh - 3 + r
                    a[:, (r * 3 + c) * 8:(r * 3 + c) * 8 + 8] = xp[:, c, 2 * oh + r, 2 * m + 1:2 * m + 9]
            out[:, :, oh, m] = a @ wk.T
    assert torch.allclose(out, ref, atol=1e-12)


def test_stride2_dgrad_parity_classes():
    g = torch.Generator().manual_seed(3)
    for (R, pad) in ((3, 1), (1, 0)):
        B, H, W, C, K = 2, 8, 6, 3, 4
        Ho, Wo = H // 2, W // 2
        w = torch.randn(K, C, R, R, generator=g, dtype=torch.float64)
        dy = torch.randn(B, K, Ho, Wo, generator=g, dtype=torch.float64)
        x = torch.zeros(B, C, H, W, dtype=torch.float64, requires_grad=True)
        F.conv2d(x, w, None, 2, pad).backward(dy)
        dx = torch.zeros(B, C, H, W, dtype=torch.float64)
        dyp = F.pad(dy, (0, 1, 0, 1))  # window offsets are in {0, 1}; reads past the edge are zero (TMA fill)
        for ca in range(2):
            for cb in range(2):
                r_first, s_first = (ca + pad) & 1, (cb + pad) & 1
                for r in range(r_first, R, 2):
                    for s in range(s_first, R, 2):
                        off_h, off_w = (ca + pad - r) // 2, (cb + pad - s) // 2
                        assert off_h in (0, 1) and off_w in (0, 1)
                        src = dyp[:, :, off_h:off_h + Ho, off_w:off_w + Wo]  # dy[i + off_h, j + off_w]
                        dx[:, :, ca::2, cb::2] += torch.einsum("bkij,kc->bcij", src, w[:, :, r, s])
        assert torch.allclose(dx, x.grad, atol=1e-12)


def test_encrypted_resnet_conv_geometry_walks_the_forwards_size_arithmetic():
    """ring/resnet.py conv_geometry(size): the (name, Cin, H_in, Cout, k, stride, pad) list the offline phase uses to find each
    layer's triple -- equal to the 224 table, and consistent with torch's own conv / pool output sizes at other input sizes"""
    import torch

    from primia_b200.ring.resnet import RESNET18_CONVS, conv_geometry, triple_shapes

    assert conv_geometry(224) == RESNET18_CONVS and len(RESNET18_CONVS) == 20
    assert [n for n, _ in triple_shapes()] == [c[0] for c in RESNET18_CONVS] + ["fc"]
    for size in (32, 64, 96, 100, 224, 256):
        geo = {name: (C, H, Co, k, s, p) for name, C, H, Co, k, s, p in conv_geometry(size)}
        x = torch.zeros(1, 3, size, size)
        x = torch.nn.functional.conv2d(x, torch.zeros(64, 3, 7, 7), stride=2, padding=3)
        assert geo["conv1"][1] == size
        x = torch.nn.functional.max_pool2d(x, 3, 2, 1)
        for li, planes in enumerate((64, 128, 256, 512), start=1):
            for bi in range(2):
                C, H, Co, k, s, p = geo[f"layer{li}.{bi}.conv1"]
                assert (C, H, Co) == (x.shape[1], x.shape[2], planes)
                y = torch.nn.functional.conv2d(x, torch.zeros(Co, C, k, k), stride=s, padding=p)
                C2, H2, Co2, k2, s2, p2 = geo[f"layer{li}.{bi}.conv2"]
                assert (C2, H2) == (y.shape[1], y.shape[2])
                if f"layer{li}.{bi}.downsample.0" in geo:
                    Cd, Hd, Cod, kd, sd, pd = geo[f"layer{li}.{bi}.downsample.0"]
                    d = torch.nn.functional.conv2d(x, torch.zeros(Cod, Cd, kd, kd), stride=sd, padding=pd)
                    assert d.shape == y.shape
                x = torch.nn.functional.conv2d(y, torch.zeros(Co2, C2, k2, k2), stride=s2, padding=p2)
