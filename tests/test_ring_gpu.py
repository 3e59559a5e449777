"""Path E parity: CUDA ring kernels (through the C ABI) vs the CPU oracle -- bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import ring_oracle as R

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"


def rnd(gen, shape):
    return torch.randint(-(2 ** 63), 2 ** 63 - 1, tuple(shape), dtype=torch.int64, generator=gen)


def cu(t):
    return t.contiguous().to(DEV)


@pytest.fixture(scope="module")
def ring():
    import primia_b200.ring as ring

    return ring


def test_encode_decode_bit_exact(ring):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100003, generator=g) * 3
    x[:6] = torch.tensor([0.0, -0.0, 0.123456789, -0.123456789, 1e-9, -7.5])
    for base, pf in [(10, 16), (10, 4), (10, 3), (2, 20)]:
        q = ring.encode(cu(x), base, pf)
        ref = R.encode(x, base, pf)
        assert torch.equal(q.cpu(), ref)
        assert torch.equal(ring.decode(q, base, pf).cpu(), R.decode(ref, base, pf))
    with pytest.raises(AssertionError):
        ring.encode(cu(torch.tensor([1e5])), 10, 16)  # 1e21 > 2^63: precision.py:122-127
    assert ring.encode(cu(torch.zeros(0)), 10, 4).numel() == 0  # empty input


def test_share_gen_reconstructs_and_is_deterministic(ring):
    g = torch.Generator().manual_seed(1)
    for n in (1, 2, 7, 4097):
        q = rnd(g, (n,))
        s0, s1 = ring.share_gen(cu(q), 123, 5)
        assert torch.equal((s0 + s1).cpu(), q)
        t0, _ = ring.share_gen(cu(q), 123, 5)
        assert torch.equal(s0, t0)
        u0, _ = ring.share_gen(cu(q), 123, 6)
        assert not torch.equal(s0, u0) or n == 0
    r = ring.random_i64((1 << 16,), 9, 1, DEV).cpu()
    bits = ((r.view(-1, 1) >> torch.arange(64)) & 1).float().mean(0)
    assert (bits - 0.5).abs().max() < 0.02
    assert (r == 2 ** 63 - 1).sum() == 0


def test_im2col_and_post_conv_vs_reference_fixture(ring):
    g = np.load(os.path.join(GOLDEN, "ring_preconv.npz"))
    for i, (B, C, H, W, Co, k, s, p) in enumerate(g["cases"]):
        x, w = torch.from_numpy(g[f"x{i}"]), torch.from_numpy(g[f"w{i}"])
        im, wr, b_, co_, ho_, wo_ = ring.functional._pre_conv(cu(x), cu(w), None, int(s), int(p))
        assert torch.equal(im.cpu(), torch.from_numpy(g[f"im{i}"]))
        assert torch.equal(wr.cpu(), torch.from_numpy(g[f"wr{i}"]))
        res = ring.matmul(im, wr.contiguous())
        post = ring.functional._post_conv(None, res, b_, co_, ho_, wo_)
        assert torch.equal(post.cpu(), torch.from_numpy(g[f"post{i}"]))


def test_spdz_fixture_from_reference_sources(ring):
    g = np.load(os.path.join(GOLDEN, "ring_spdz.npz"))
    for i in range(int(g["n"])):
        op = str(g[f"op{i}"])
        parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
        xs = [cu(torch.from_numpy(g[f"x{i}_{j}"])) for j in range(2)]
        ys = [cu(torch.from_numpy(g[f"y{i}_{j}"])) for j in range(2)]
        for j, p in enumerate(parties):
            tri = tuple(cu(torch.from_numpy(g[f"{n}{i}_{j}"])) for n in "abc")
            p.crypto_store.add_primitives(op, (xs[j].shape, ys[j].shape), [tri])
        for j, p in enumerate(parties):
            d, e = ring.spdz_mask(p, xs[j], ys[j], op)
            assert torch.equal(d.cpu(), torch.from_numpy(g[f"d{i}_{j}"]))
            assert torch.equal(e.cpu(), torch.from_numpy(g[f"e{i}_{j}"]))
        z = ring.spdz_mul(op, xs, ys, parties)
        for j in range(2):
            assert torch.equal(z[j].cpu(), torch.from_numpy(g[f"z{i}_{j}"]))
        assert parties[0].crypto_store.count(op, (xs[0].shape, ys[0].shape)) == 0  # consumed (remove=True)
        with pytest.raises(ring.EmptyCryptoPrimitiveStoreError):
            ring.spdz_mul(op, xs, ys, parties)


@pytest.mark.parametrize("B,M,K,N", [(1, 49, 147, 64), (2, 100, 64, 128), (1, 1, 512, 3), (1, 196, 2304, 256), (3, 65, 17, 65),
                                     (1, 3136, 576, 64)])
def test_combine_matmul_bit_exact(ring, B, M, K, N):
    g = torch.Generator().manual_seed(B * 1000 + M)
    delta, a, eps, b, c = rnd(g, (B, M, K)), rnd(g, (B, M, K)), rnd(g, (K, N)), rnd(g, (K, N)), rnd(g, (B, M, N))
    for j in range(2):
        z = ring.combine_matmul(j, cu(delta), cu(eps), cu(a), cu(b), cu(c))
        ref = R.spdz_compute(j, delta, eps, a, b, c, "matmul")
        assert torch.equal(z.cpu(), ref), (j, B, M, K, N)
    assert torch.equal(ring.matmul(cu(a), cu(b)).cpu(), torch.matmul(a, b))


def test_combine_mul_broadcast_modes(ring):
    g = torch.Generator().manual_seed(5)
    P, C = 37, 64
    for ls, rs in [((P, C), (P, C)), ((C,), (P, C)), ((P, C), (C,)), ((C,), (C,))]:
        delta, a, eps, b = rnd(g, ls), rnd(g, ls), rnd(g, rs), rnd(g, rs)
        c = rnd(g, torch.broadcast_shapes(ls, rs))
        for j in range(2):
            z = ring.combine_mul(j, cu(delta), cu(eps), cu(a), cu(b), cu(c))
            assert torch.equal(z.cpu(), R.spdz_compute(j, delta, eps, a, b, c, "mul"))


def test_trunc_div_edge_cases(ring):
    s = torch.tensor([7, -7, 19999, -19999, 0, -(2 ** 63), 2 ** 63 - 1, -1, 1], dtype=torch.int64)
    for d in (10, 10 ** 4, 10 ** 16, 49, 20):
        assert torch.equal(ring.trunc_div(cu(s), d).cpu(), R.trunc_div(s, d))


CONV_SHAPES = [  # (C, H, Cout, k, stride, pad): every distinct ResNet-18 conv geometry at reduced spatial size
    (3, 32, 64, 7, 2, 3), (64, 8, 64, 3, 1, 1), (64, 8, 128, 3, 2, 1), (64, 8, 128, 1, 2, 0), (128, 6, 128, 3, 1, 1),
    (128, 6, 256, 3, 2, 1), (128, 6, 256, 1, 2, 0), (256, 4, 256, 3, 1, 1), (256, 4, 512, 3, 2, 1), (256, 4, 512, 1, 2, 0),
    (512, 3, 512, 3, 1, 1),
]


def _shared_conv_case(ring, g, B, C, H, Co, k, s, p, base, pf):
    x, w = rnd(g, (B, C, H, H)), rnd(g, (Co, C, k, k))
    xs = R.share_from_random(x, rnd(g, x.shape))
    ws = R.share_from_random(w, rnd(g, w.shape))
    Ho = (H + 2 * p - k) // s + 1
    M, K, N = Ho * Ho, C * k * k, Co
    a, b = rnd(g, (B, M, K)), rnd(g, (K, N))
    c = R.build_triple_c(a, b, "matmul")
    a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
    tri = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
    ref = R.conv2d_shared(xs, ws, tri, s, p, base, pf)
    parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
    for j, pty in enumerate(parties):
        pty.crypto_store.add_primitives("matmul", ((B, M, K), (K, N)), [tuple(cu(t) for t in tri[j])])
    X = ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([cu(t) for t in xs], parties), base, pf)
    Wt = ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([cu(t) for t in ws], parties), base, pf)
    out = ring.functional.conv2d(X, Wt, None, s, p)
    for j in range(2):
        assert torch.equal(out.child.child[j].cpu(), ref[j]), (C, H, Co, k, s, p, j)
    # the same layer with its image-independent half hoisted (offline: mask + open the weights, limb planes of a, b + eps, eps;
    # online: mask x, open INTO the planes, one 2-segment GEMM): same shares, bit for bit
    gpu_tri = [tuple(cu(t) for t in tri[j]) for j in range(2)]
    prep = ring.functional.prepare_weight_side(Wt, gpu_tri, B, Ho, Ho)
    from primia_b200.ring import ops
    assert (prep is not None) == bool(ops.tc_supported(B * M, K, N))
    if prep is not None:
        for j, pty in enumerate(parties):
            pty.crypto_store.add_primitives("matmul", ((B, M, K), (K, N)), [gpu_tri[j]])
        out2 = ring.functional.conv2d(X, Wt, None, s, p, prepared=prep)
        assert all(pty.crypto_store.count("matmul", ((B, M, K), (K, N))) == 0 for pty in parties)   # popped like spdz_compute
        for j in range(2):
            assert torch.equal(out2.child.child[j].cpu(), ref[j]), ("hoisted", C, H, Co, k, s, p, j)


@pytest.mark.parametrize("base,pf", [(10, 16), (10, 4)])
def test_conv2d_on_shares_all_resnet18_geometries(ring, base, pf):
    g = torch.Generator().manual_seed(42)
    for (C, H, Co, k, s, p) in CONV_SHAPES:
        _shared_conv_case(ring, g, 1, C, H, Co, k, s, p, base, pf)
    _shared_conv_case(ring, g, 2, 16, 9, 24, 3, 2, 1, base, pf)  # batch > 1, ragged M


def test_conv2d_full_size_layer1(ring):
    """BASELINE config C4 geometry: layer1 conv, 64ch 56x56, M=3136 K=576 N=64, pf=16."""
    g = torch.Generator().manual_seed(7)
    _shared_conv_case(ring, g, 1, 64, 56, 64, 3, 1, 1, 10, 16)


FULL_SIZE_SHAPES = [  # BASELINE config C4: every distinct conv geometry of ResNet-18 on ONE 224x224 image, at full size
    (3, 224, 64, 7, 2, 3), (64, 56, 64, 3, 1, 1), (64, 56, 128, 3, 2, 1), (64, 56, 128, 1, 2, 0), (128, 28, 128, 3, 1, 1),
    (128, 28, 256, 3, 2, 1), (128, 28, 256, 1, 2, 0), (256, 14, 256, 3, 1, 1), (256, 14, 512, 3, 2, 1), (256, 14, 512, 1, 2, 0),
    (512, 7, 512, 3, 1, 1),
]


@pytest.mark.parametrize("shape", FULL_SIZE_SHAPES, ids=lambda s: "C%d_H%d_K%d_k%d_s%d" % s[:5])
def test_conv2d_on_shares_full_size_geometries_pf16(ring, shape):
    """the Beaver conv at the sizes encrypted inference of a 224x224 image runs (M up to 12544, K up to 4608), pf = 16"""
    g = torch.Generator().manual_seed(1000 + shape[0] + shape[1])
    _shared_conv_case(ring, g, 1, *shape, 10, 16)


def test_provider_generated_triples_reconstruct_true_product(ring):
    """full-size property: with provider-made triples, reconstruct(z0+z1) == im2col(x) @ w^T mod 2^64."""
    g = torch.Generator().manual_seed(9)
    parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", DEV), seed=77)
    x, w = rnd(g, (1, 64, 28, 28)), rnd(g, (128, 64, 3, 3))
    X = ring.FixedPrecisionTensor(cu(x), 10, 16).share(*parties, crypto_provider=prov)
    Wt = ring.FixedPrecisionTensor(cu(w), 10, 16).share(*parties, crypto_provider=prov)
    x_sh = [s.cpu() for s in X.child.child]
    assert torch.equal(x_sh[0] + x_sh[1], x)
    # untruncated protocol output
    from primia_b200.ring import ops
    im, wr, *_ = R.pre_conv(x, w, 2, 1)
    z = ring.spdz_mul("matmul", [ops.im2col(s, 3, 3, 2, 1) for s in X.child.child],
                      [s.reshape(128, -1).t().contiguous() for s in Wt.child.child], parties, prov)
    assert torch.equal((z[0] + z[1]).cpu(), torch.matmul(im, wr))


class ReplayRNG:
    def __init__(self, s0_list):
        self.s0 = list(s0_list)

    def share(self, q):
        if self.s0[0].numel() == q.numel():
            s0 = self.s0.pop(0).to(q.device)
        else:  # the fused Newton kernel shares all its constants in one call
            s0 = torch.cat([self.s0.pop(0).reshape(-1) for _ in range(q.numel())]).reshape(q.shape).to(q.device)
        return s0, q - s0


def test_batch_norm_eval_newton_bit_exact(ring):
    g = torch.Generator().manual_seed(11)
    base, pf = 10, 4
    B, C, H, W = 1, 8, 3, 2
    P = B * H * W
    enc = lambda t: R.encode(t, base, pf)
    x = enc(torch.randn(B, C, H, W, generator=g))
    mean, var = enc(torch.randn(C, generator=g) * 0.1), enc(torch.rand(C, generator=g) + 0.5)
    gamma, beta = enc(torch.rand(C, generator=g) + 0.5), enc(torch.randn(C, generator=g) * 0.1)
    sh = lambda q: R.share_from_random(q, rnd(g, q.shape))
    x_sh, m_sh, v_sh, g_sh, b_sh = sh(x), sh(mean), sh(var), sh(gamma), sh(beta)

    def tri(ls, rs):
        a, b = rnd(g, ls), rnd(g, rs)
        c = a * b
        a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
        return [(a0, b0, c0), (a - a0, b - b0, c - c0)]

    iters = 80
    q21 = torch.tensor([21 * base ** pf], dtype=torch.int64)
    c_s0 = [rnd(g, (1,)) for _ in range(iters)]
    consts = [[s0, q21 - s0] for s0 in c_s0]
    triples = [None] + [[tri((C,), (C,)) for _ in range(3)] for _ in range(iters - 1)]
    tri_norm, tri_aff = tri((C,), (P, C)), tri((P, C), (C,))
    ref = R.batch_norm_eval_shared(x_sh, m_sh, v_sh, g_sh, b_sh, consts, triples, tri_norm, tri_aff, base, pf, iters)

    parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
    for j, pty in enumerate(parties):
        for it in range(1, iters):
            for t in triples[it]:
                pty.crypto_store.add_primitives("mul", ((C,), (C,)), [tuple(cu(u) for u in t[j])])
        pty.crypto_store.add_primitives("mul", ((C,), (P, C)), [tuple(cu(u) for u in tri_norm[j])])
        pty.crypto_store.add_primitives("mul", ((P, C), (C,)), [tuple(cu(u) for u in tri_aff[j])])
    rng = ReplayRNG(c_s0)
    mk = lambda s: ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([cu(t) for t in s], parties, None, rng), base, pf)
    out = ring.functional.batch_norm(mk(x_sh), mk(m_sh), mk(v_sh), mk(g_sh), mk(b_sh))
    for j in range(2):
        assert torch.equal(out.child.child[j].cpu(), ref[j])
    # the op-by-op protocol (parties on different GPUs take this path) produces the same shares as the fused kernel
    from primia_b200.ring import tensors as T
    for j, pty in enumerate(parties):
        for it in range(1, iters):
            for t in triples[it]:
                pty.crypto_store.add_primitives("mul", ((C,), (C,)), [tuple(cu(u) for u in t[j])])
        pty.crypto_store.add_primitives("mul", ((C,), (P, C)), [tuple(cu(u) for u in tri_norm[j])])
        pty.crypto_store.add_primitives("mul", ((P, C), (C,)), [tuple(cu(u) for u in tri_aff[j])])
    rng.s0 = list(c_s0)
    T.FUSE_NEWTON = False
    try:
        out2 = ring.functional.batch_norm(mk(x_sh), mk(m_sh), mk(v_sh), mk(g_sh), mk(b_sh))
    finally:
        T.FUSE_NEWTON = True
    for j in range(2):
        assert torch.equal(out2.child.child[j].cpu(), ref[j])
    # the hoisted evaluation (offline: Newton, open(inv_std - a1), open(weight - b2), d1, d2 in NCHW; online: three elementwise
    # passes per party) produces the same shares as the oracle's op-by-op protocol
    for j, pty in enumerate(parties):
        for it in range(1, iters):
            for t in triples[it]:
                pty.crypto_store.add_primitives("mul", ((C,), (C,)), [tuple(cu(u) for u in t[j])])
    rng.s0 = list(c_s0)
    inv = mk(v_sh).reciprocal(method="newton")
    gt = lambda t: [tuple(cu(u) for u in t[j]) for j in range(2)]
    t1, t2 = gt(tri_norm), gt(tri_aff)
    side = ring.functional.prepare_bn_side(inv, mk(g_sh), t1, t2, B, C, H, W)
    for j, pty in enumerate(parties):
        pty.crypto_store.add_primitives("mul", ((C,), (P, C)), [t1[j]])
        pty.crypto_store.add_primitives("mul", ((P, C), (C,)), [t2[j]])
    out3 = ring.functional.batch_norm_prepared(mk(x_sh), mk(m_sh), mk(b_sh), side)
    for j in range(2):
        assert torch.equal(out3.child.child[j].cpu(), ref[j]), "hoisted batch_norm"
    assert all(p.crypto_store.nbytes() == 0 for p in parties)
    # numerically sane at pf=4: decodes to the float batch norm (eps ignored) within fixed-point error
    got = R.decode(ref[0] + ref[1], base, pf)
    xf, mf, vf, gf, bf = (R.decode(t, base, pf) for t in (x, mean, var, gamma, beta))
    want = (xf - mf.view(1, C, 1, 1)) / vf.view(1, C, 1, 1).sqrt() * gf.view(1, C, 1, 1) + bf.view(1, C, 1, 1)
    assert (got - want).abs().max() < 0.05


def test_avgpool_and_linear_bit_exact(ring):
    g = torch.Generator().manual_seed(13)
    base, pf = 10, 16
    x = rnd(g, (1, 512, 7, 7))
    xs = R.share_from_random(x, rnd(g, x.shape))
    ref = R.avg_pool_shared(xs, 7)
    parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
    X = ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([cu(t) for t in xs], parties), base, pf)
    out = ring.functional.avg_pool2d(X, 7)
    for j in range(2):
        assert torch.equal(out.child.child[j].cpu(), ref[j])
    feat = [r.reshape(1, 512) for r in ref]
    w, bias = rnd(g, (3, 512)), rnd(g, (3,))
    ws, bs = R.share_from_random(w, rnd(g, w.shape)), R.share_from_random(bias, rnd(g, bias.shape))
    a, b = rnd(g, (1, 512)), rnd(g, (512, 3))
    c = a @ b
    a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
    tri = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
    ref_l = R.linear_shared(feat, ws, bs, tri, base, pf)
    for j, pty in enumerate(parties):
        pty.crypto_store.add_primitives("matmul", ((1, 512), (512, 3)), [tuple(cu(t) for t in tri[j])])
    mk = lambda s: ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([cu(t) for t in s], parties), base, pf)
    out = ring.functional.linear(mk(feat), mk(ws), mk(bs))
    for j in range(2):
        assert torch.equal(out.child.child[j].cpu(), ref_l[j])


@pytest.mark.parametrize("rows,K,N", [(128, 128, 32), (200, 147, 64), (49, 4608, 512), (12544, 147, 64), (777, 300, 96)])
def test_int8_limb_tensor_core_gemm_equals_imad_gemm(ring, rows, K, N):
    """tcgen05.mma.kind::i8 limb decomposition (ring_i8.cu) vs the integer-pipe GEMM and vs torch CPU: bit-exact mod 2^64."""
    from primia_b200.ring import ops

    g = torch.Generator().manual_seed(rows + K + N)
    A1, A2 = rnd(g, (rows, K)), rnd(g, (rows, K))
    B1, B2 = rnd(g, (K, N)), rnd(g, (K, N))
    C0 = rnd(g, (rows, N))
    assert ops.tc_supported(rows, K, N)
    got = ops.gemm2_tc(cu(A1), cu(B1), cu(A2), cu(B2), cu(C0)).cpu()
    want = torch.matmul(A1, B1) + torch.matmul(A2, B2) + C0
    assert torch.equal(got, want)
    # extreme limbs: all-ones bytes maximise every limb-pair sum
    A1 = torch.full((rows, K), -1, dtype=torch.int64)
    B1 = torch.full((K, N), -1, dtype=torch.int64)
    got = ops.gemm2_tc(cu(A1), cu(B1), None, None, None).cpu()
    assert torch.equal(got, torch.matmul(A1, B1))
    old = ops.USE_TENSOR_CORES
    try:
        ops.USE_TENSOR_CORES = False
        ref = ops.matmul(cu(A2), cu(B2)).cpu()
    finally:
        ops.USE_TENSOR_CORES = old
    assert torch.equal(ops.matmul(cu(A2), cu(B2)).cpu(), ref)


def test_encrypted_linear_graph_replay_is_bit_exact(ring):
    """CUDA-graph replay of the online phase with triples refreshed into static buffers == eager protocol on the same triples."""
    from primia_b200.ring.resnet import EncryptedLinearGraph, SharedLinearLayers, triple_shapes

    parties = [ring.Party("model_owner", DEV), ring.Party("data_owner", DEV)]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", DEV), seed=5)
    net = SharedLinearLayers(parties, prov, 10, 16)
    xs = net.make_inputs(1)
    eg = EncryptedLinearGraph(net, xs, 1)
    eg.offline()
    torch.cuda.synchronize()
    # eager run on clones of exactly these triples
    for (_n, shapes), per_party in zip(triple_shapes(1, 3), eg.static):
        for p, tri in zip(parties, per_party):
            p.crypto_store.add_primitives("matmul", shapes, [tuple(t.clone() for t in tri)])
    ref = net.forward(xs)
    ref = {k: [s.clone() for s in v.child.child] for k, v in ref.items()}
    out = eg.online()
    torch.cuda.synchronize()
    for k in ("conv1", "layer2.0.conv1", "layer4.1.conv2", "fc"):
        for j in range(2):
            assert torch.equal(out[k].child.child[j], ref[k][j]), k


def test_newton_p2p_kernel_equals_the_single_gpu_fused_kernel(ring):
    """pm_bn_newton_p2p_i64 (one kernel per party, openings through mailboxes -- the cross-GPU placement) produces the very
    shares pm_bn_newton_fused_i64 produces (which test_batch_norm_eval_newton_bit_exact pins to the oracle).  Here both
    parties sit on one GPU, on two streams; tests/test_multigpu.py repeats it across two GPUs over NVLink."""
    from primia_b200.ring import ops

    g = torch.Generator().manual_seed(31)
    iters, scale, Cc = 80, 10 ** 4, 20
    jobs = []
    for C in (64, 64, 128, 512, 3):
        v = (torch.rand(C, generator=g) * 0.5 + 0.75)
        vq = R.encode(v, 10, 4)
        vs = R.share_from_random(vq, rnd(g, vq.shape))
        a, b = rnd(g, (3 * (iters - 1), C)), rnd(g, (3 * (iters - 1), C))
        c = a * b
        a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
        tri = [[cu(a0), cu(b0), cu(c0)], [cu(a - a0), cu(b - b0), cu(c - c0)]]
        kq = torch.full((iters,), 21 * scale, dtype=torch.int64)
        k0 = rnd(g, kq.shape)
        jobs.append(([cu(vs[0]), cu(vs[1])], tri, (cu(k0), cu(kq - k0))))
    ref = ops.bn_newton_fused(jobs, iters, scale, Cc)
    for rep in range(2):                       # twice: the epoch counter distinguishes the launches' messages
        got = ops.bn_newton_p2p(jobs, iters, scale, Cc)
        torch.cuda.synchronize()
        assert all(int(e.item()) == 0 for e in ops.bn_newton_p2p.last_err)
        for (r0, r1), (g0, g1) in zip(ref, got):
            assert torch.equal(r0, g0) and torch.equal(r1, g1)
    inv = ring.decode(ref[0][0] + ref[0][1], 10, 4).cpu()
    want = 1.0 / torch.sqrt(R.decode(R.reconstruct([t.cpu() for t in jobs[0][0]]), 10, 4))
    assert (inv - want).abs().max() < 5e-3
