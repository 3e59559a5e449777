"""Pin the path-E oracle against fixtures produced by executing the reference's own sources
(tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import ring_oracle as R

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_pre_post_conv_matches_reference_output():
    g = np.load(os.path.join(GOLDEN, "ring_preconv.npz"))
    for i, (B, C, H, W, Co, k, s, p) in enumerate(g["cases"]):
        x, w = T(g[f"x{i}"]), T(g[f"w{i}"])
        im, wr, b_, co_, ho_, wo_ = R.pre_conv(x, w, int(s), int(p))
        assert torch.equal(im, T(g[f"im{i}"]))
        assert torch.equal(wr, T(g[f"wr{i}"]))
        post = R.post_conv(None, torch.matmul(im, wr), b_, co_, ho_, wo_)
        assert torch.equal(post, T(g[f"post{i}"]))


def test_spdz_mask_compute_matches_reference_output():
    g = np.load(os.path.join(GOLDEN, "ring_spdz.npz"))
    for i in range(int(g["n"])):
        op = str(g[f"op{i}"])
        xs = [T(g[f"x{i}_{j}"]) for j in range(2)]
        ys = [T(g[f"y{i}_{j}"]) for j in range(2)]
        tri = [tuple(T(g[f"{n}{i}_{j}"]) for n in "abc") for j in range(2)]
        d, e = zip(*[R.spdz_mask(xs[j], ys[j], tri[j][0], tri[j][1]) for j in range(2)])
        for j in range(2):
            assert torch.equal(d[j], T(g[f"d{i}_{j}"]))
            assert torch.equal(e[j], T(g[f"e{i}_{j}"]))
        z = R.spdz_mul(op, xs, ys, tri)
        for j in range(2):
            assert torch.equal(z[j], T(g[f"z{i}_{j}"]))


def test_newton_control_flow_matches_reference_trace():
    """oracle.newton_inv_sqrt_like must perform exactly the reference's op sequence."""
    g = np.load(os.path.join(GOLDEN, "ring_newton.npz"))
    trace = [t.split("|") for t in g["trace"]]
    # iteration 0: rsub(21 - v), div 20 ; iterations 1..79: mul(x,x), mul(v,xx), rsub 21, mul(y,x), div 20
    assert len(trace) == 2 + 79 * 5
    assert [t[1] for t in trace[:2]] == ["rsub", "div"]
    assert trace[0][3] == "21" and trace[1][3] == "20"
    for it in range(79):
        ops = trace[2 + it * 5: 7 + it * 5]
        assert [o[1] for o in ops] == ["mul", "mul", "rsub", "mul", "div"]
        x_prev = trace[1 + it * 5][0]
        assert ops[0][2] == x_prev and ops[0][3] == x_prev  # x * x
        assert ops[1][2] == "v" and ops[1][3] == ops[0][0]  # self * (x*x)
        assert ops[2][2] == ops[1][0] and ops[2][3] == "21"  # C + 1 - ...
        assert ops[3][2] == ops[2][0] and ops[3][3] == x_prev  # y * x
        assert ops[4][2] == ops[3][0] and ops[4][3] == "20"  # / C


def test_encode_fp32_promotion_and_trunc():
    # probed 0.123456789 -> 1234567948140544 at (10,16)
    x = torch.tensor([0.123456789, -0.123456789, 1.5, -2.75], dtype=torch.float32)
    q = R.encode(x, 10, 16)
    assert q[0].item() == 1234567948140544 and q[1].item() == -1234567948140544
    q4 = R.encode(x, 10, 4)
    assert q4.tolist() == [1234, -1234, 15000, -27500]
    assert torch.allclose(R.decode(q4, 10, 4), torch.tensor([0.1234, -0.1234, 1.5, -2.75]))


def test_trunc_div_is_c_truncation():
    s = torch.tensor([7, -7, 19999, -19999, 0, -(2 ** 63)], dtype=torch.int64)
    assert R.trunc_div(s, 10).tolist() == [0, 0, 1999, -1999, 0, -922337203685477580]


def test_beaver_reconstructs_product_small_precision():
    g = torch.Generator().manual_seed(1)
    base, pf = 10, 4
    x = torch.randn(1, 3, 8, 8, generator=g)
    w = torch.randn(4, 3, 3, 3, generator=g) * 0.2
    qx, qw = R.encode(x, base, pf), R.encode(w, base, pf)
    r = lambda s: torch.randint(-(2 ** 63), 2 ** 63 - 1, s, dtype=torch.int64, generator=g)
    xs = R.share_from_random(qx, r(qx.shape))
    ws = R.share_from_random(qw, r(qw.shape))
    M, K, N = 64, 27, 4
    a, b = r((1, M, K)), r((K, N))
    c = R.build_triple_c(a, b, "matmul")
    a0, b0, c0 = r(a.shape), r(b.shape), r(c.shape)
    tri = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
    out = R.conv2d_shared(xs, ws, tri, 1, 1, base, pf)
    got = R.decode(R.reconstruct(out), base, pf)
    ref = torch.nn.functional.conv2d(x, w, padding=1)
    # per-share truncation is off by at most 1 ulp of 10**-pf (+ wrap with prob ~ 2**-30)
    assert (got - ref).abs().max() < 5e-3


def test_avgpool_and_linear_shapes():
    g = torch.Generator().manual_seed(2)
    r = lambda s: torch.randint(-(2 ** 63), 2 ** 63 - 1, s, dtype=torch.int64, generator=g)
    x = torch.randint(-1000, 1000, (1, 4, 14, 14), dtype=torch.int64, generator=g)
    xs = R.share_from_random(x, r(x.shape))
    out = R.avg_pool_shared(xs, 7)
    assert out[0].shape == (1, 4, 2, 2)
    exact = x.reshape(1, 4, 2, 7, 2, 7).permute(0, 1, 2, 4, 3, 5).reshape(1, 4, 2, 2, 49).sum(-1)
    # share-wise truncation == exact/49 up to +-1 (and a 2**64/49 wrap term that reconstructs modulo)
    rec = R.reconstruct(out)
    wrap = (2 ** 64) // 49
    diff = (rec - torch.div(exact, 49, rounding_mode="trunc"))
    assert all(min(abs(int(d)), abs(abs(int(d)) - wrap)) <= 2 for d in diff.flatten())


def test_pool_restatement_matches_reference_sources():
    """_pre_pool/_post_pool outputs and the op order of _pool2d's max branch, executed from the reference (ring_pool.npz)."""
    import numpy as np
    from oracle import fss_oracle as F

    g = np.load(os.path.join(GOLDEN, "ring_pool.npz"))
    rng = np.random.default_rng(3)
    for i, (B, C, H, W, k, st, pd) in enumerate(g["cases"]):
        x = torch.from_numpy(g[f"x{i}"])
        im, *_ = R.pre_pool(x, int(k), int(st), int(pd))
        assert torch.equal(im, torch.from_numpy(g[f"im{i}"]))
        ref_trace = [t for t in g[f"trace{i}"] if t.startswith("ge|")]
        shapes = [eval(t.split("|")[1]) for t in ref_trace]
        x0 = torch.from_numpy(rng.integers(-2 ** 62, 2 ** 62, x.shape))
        xs = [x0, x - x0]
        keys, alphas, tris = [], [], []
        for shp in shapes:
            n = int(np.prod(shp))
            alpha = rng.integers(0, 2 ** 32, n, dtype=np.uint64)
            keys.append(F.dif_keygen(alpha, rng.integers(0, 2 ** 63, (2, 2, n), dtype=np.uint64)))
            alphas.append(F.split_alpha(alpha, rng.integers(0, 2 ** 32, n, dtype=np.uint64)))
            a, b = (torch.from_numpy(rng.integers(-2 ** 63, 2 ** 63 - 1, shp)) for _ in range(2))
            a0, b0, c0 = (torch.from_numpy(rng.integers(-2 ** 63, 2 ** 63 - 1, shp)) for _ in range(3))
            tris.append([(a0, b0, c0), (a - a0, b - b0, a * b - c0)])
        trace = []
        out = R.max_pool2d_shared(xs, int(k), int(st), int(pd), keys, alphas, tris, trace)
        assert trace == shapes                                   # same comparisons, same order, same shapes
        assert torch.equal(out[0] + out[1], torch.from_numpy(g[f"max{i}"]))


def test_relu_on_shares_reconstructs():
    import numpy as np
    from oracle import fss_oracle as F

    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.integers(-10 ** 6, 10 ** 6, (2, 3, 4, 4)))
    x.view(-1)[:3] = torch.tensor([0, 1, -1])
    x0 = torch.from_numpy(rng.integers(-2 ** 62, 2 ** 62, x.shape))
    n = x.numel()
    alpha = rng.integers(0, 2 ** 32, n, dtype=np.uint64)
    key = F.dif_keygen(alpha, rng.integers(0, 2 ** 63, (2, 2, n), dtype=np.uint64))
    al = F.split_alpha(alpha, rng.integers(0, 2 ** 32, n, dtype=np.uint64))
    a, b = (torch.from_numpy(rng.integers(-2 ** 63, 2 ** 63 - 1, x.shape)) for _ in range(2))
    a0, b0, c0 = (torch.from_numpy(rng.integers(-2 ** 63, 2 ** 63 - 1, x.shape)) for _ in range(3))
    out = R.relu_shared([x0, x - x0], key, al, [(a0, b0, c0), (a - a0, b - b0, a * b - c0)])
    assert torch.equal(out[0] + out[1], torch.clamp(x, min=0))


def test_full_encrypted_forward_oracle_tracks_plaintext_model():
    """resnet18_forward_shared (the restatement of inference.py:279-321) on a 32x32 image at base 10, pf 4: the decoded
    logits follow the plaintext model with pool/relu swapped (inference.py:289)."""
    from oracle import train_oracle as O

    base, pf, size = 10, 4, 32
    torch.manual_seed(42)
    model = O.ResNet18(input_size=size)
    gg = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=gg) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=gg) * 0.5 + 0.75)
                m.weight.copy_(torch.rand(m.num_features, generator=gg) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=gg) * 0.1)
    model.eval()
    g = torch.Generator().manual_seed(8)
    img = torch.randn(1, 3, size, size, generator=g)
    tape = R.GeneratingTape(1, 21 * base ** pf)
    sh = lambda q: [(s0 := tape._r(q.shape)), q - s0]
    P = {k: sh(R.encode(v.float().contiguous(), base, pf)) for k, v in model.state_dict().items()
         if not k.endswith("num_batches_tracked")}
    out = R.resnet18_forward_shared(P, sh(R.encode(img, base, pf)), tape, base, pf, size)
    logits = R.decode(out[0] + out[1], base, pf)
    with torch.no_grad():
        model.pool, model.relu = model.relu, model.pool
        want = model(img)
    assert (logits - want).abs().max() < 0.15, (logits, want)
    # protocol accounting: 20 convs + fc matmul triples, 20 BN layers x 239 elementwise triples, 17 ReLUs + 4 pool steps
    assert tape.n_consts == 20 * 80
    assert tape.n_triples == 21 + 20 * 239 + 17 + 4


def test_hoisted_protocol_algebra_equals_the_reference_protocol():
    """The offline / online split the CUDA path runs (ring/functional.py prepare_weight_side, prepare_bn_side) is the reference's
    spdz_mul (spdz.py:125-197) with the terms regrouped -- exact in Z_2^64.  Written here on the oracle's plain int64 tensors:

      conv:  z_j = delta @ (b_j + [j=0] eps) + (a_j @ eps + c_j)
      bn 1:  z_j = (a1_j + [j=0] delta1)[c] * eps1 + (delta1[c] * b1_j + c1_j)          inv_std [C] x (flat - mean) [P,C]
      bn 2:  z_j = delta2 * (b2_j + [j=0] eps2)[c] + (a2_j * eps2[c] + c2_j)            normalized [P,C] x weight [C]
    """
    g = torch.Generator().manual_seed(21)
    rnd = lambda *s: torch.randint(-(2 ** 63), 2 ** 63 - 1, s, dtype=torch.int64, generator=g)
    share = lambda q: R.share_from_random(q, rnd(*q.shape))

    def triple(ls, rs, op):
        a, b = rnd(*ls), rnd(*rs)
        c = R.build_triple_c(a, b, op)
        return list(zip(share(a), share(b), share(c)))

    # ---- Beaver matmul with the weight side known in advance
    M, K, N = 12, 20, 8
    x, w = share(rnd(M, K)), share(rnd(K, N))
    tri = triple((M, K), (K, N), "matmul")
    ref = R.spdz_mul("matmul", x, w, tri)
    eps = sum(w[j] - tri[j][1] for j in range(2))                                       # offline: open(w - b)
    b_eff = [tri[0][1] + eps, tri[1][1]]
    c_prime = [torch.matmul(tri[j][0], eps) + tri[j][2] for j in range(2)]              # offline: a @ eps + c
    delta = sum(x[j] - tri[j][0] for j in range(2))                                     # online: open(x - a)
    got = [torch.matmul(delta, b_eff[j]) + c_prime[j] for j in range(2)]
    assert all(torch.equal(got[j], ref[j]) for j in range(2))
    # ---- BatchNorm's two broadcast products
    P, C = 10, 6
    inv, centered, gamma = share(rnd(C)), share(rnd(P, C)), share(rnd(C))
    t1, t2 = triple((C,), (P, C), "mul"), triple((P, C), (C,), "mul")
    ref1 = R.spdz_mul("mul", inv, centered, t1)
    delta1 = sum(inv[j] - t1[j][0] for j in range(2))                                   # offline
    s1 = [t1[0][0] + delta1, t1[1][0]]
    d1 = [delta1 * t1[j][1] + t1[j][2] for j in range(2)]
    eps1 = sum(centered[j] - t1[j][1] for j in range(2))                                # online
    got1 = [s1[j] * eps1 + d1[j] for j in range(2)]
    assert all(torch.equal(got1[j], ref1[j]) for j in range(2))
    ref2 = R.spdz_mul("mul", ref1, gamma, t2)
    eps2 = sum(gamma[j] - t2[j][1] for j in range(2))                                   # offline
    s2 = [t2[0][1] + eps2, t2[1][1]]
    d2 = [t2[j][0] * eps2 + t2[j][2] for j in range(2)]
    delta2 = sum(ref1[j] - t2[j][0] for j in range(2))                                  # online
    got2 = [delta2 * s2[j] + d2[j] for j in range(2)]
    assert all(torch.equal(got2[j], ref2[j]) for j in range(2))
