"""MPC versions of torch.nn.functional ops on FixedPrecisionTensor > AdditiveSharingTensor.

Mirrors syft/frameworks/torch/nn/functional.py: conv2d :204-308 (_pre_conv :79-166, _post_conv :170-201),
batch_norm :44-75, avg_pool2d :460-525, linear :10-14."""
from __future__ import annotations

import torch

from . import ops
from .spdz import EmptyCryptoPrimitiveStoreError, _key, open_planes, open_shares, spdz_compute
from .tensors import FixedPrecisionTensor


def _pre_conv(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Party-local im2col (functional.py:79-166); kept for API parity -- conv2d itself never materialises it."""
    assert groups == 1 and input.dim() == 4 and weight.dim() == 4
    B, C, H, W = input.shape
    Co, Ck, kh, kw = weight.shape
    assert C == Ck
    im = ops.im2col(input, kh, kw, stride, padding, dilation)
    Ho = ops.conv_out_size(H, kh, stride, padding, dilation)
    Wo = ops.conv_out_size(W, kw, stride, padding, dilation)
    w_r = weight.reshape(Co, -1).t()
    return im, w_r, B, Co, Ho, Wo


def _post_conv(bias, res, batch_size, nb_channels_out, nb_rows_out, nb_cols_out):
    """functional.py:170-201 (no truncation: divisor 1)"""
    return ops.trunc_post_conv(res, 1, bias, nb_rows_out, nb_cols_out)


class WeightSide:
    """Everything of one Beaver matmul that does NOT depend on the image, per party j: the limb planes of the right operand
    b_j (+ eps on party 0) and c'_j = a_j @ eps + c_j.  Built by ``prepare_weight_side`` in the offline phase, consumed by
    ``conv2d(..., prepared=)``."""

    __slots__ = ("shapes", "pb1", "cprime")
    TENSORS = ("pb1", "cprime")

    def __init__(self, shapes, pb1, cprime):
        self.shapes, self.pb1, self.cprime = shapes, pb1, cprime


def prepare_weight_side(weight: FixedPrecisionTensor, tri, batch: int, Ho: int, Wo: int):
    """The weight half of spdz_mul for one convolution (spdz.py:22-45,162-163), hoisted out of the per-image path:
    eps = open(w_j^T - b_j); party 0's right operand b_0 + eps as limb planes; c'_j = a_j @ eps + c_j (one ring GEMM per party, so
    that the online product is the single segment  z_j = delta @ (b_j [+ eps]) + c'_j).  ``tri``: the per-party triple the layer is
    going to consume (peeked, not popped).  Returns None when the shape does not run on the tensor cores."""
    w = weight.child
    Co = w.shape[0]
    K = w.child[0].numel() // Co
    rows = batch * Ho * Wo
    if not ops.tc_supported(rows, K, Co):
        return None
    parties = w.parties
    e_sh = [ops.mask_wt(w.child[j].reshape(Co, K), tri[j][1]) for j in range(2)]
    eps = open_shares(parties, e_sh)
    pb1 = [ops.planarize_cols(ops.axpby(1, tri[0][1], 1, eps[0])), ops.planarize_cols(tri[1][1])]
    cprime = [ops.gemm_planes(ops.planarize_rows(tri[j][0]), ops.planarize_cols(eps[j]), None, None, tri[j][2].reshape(rows, Co),
                              rows, K, Co) for j in range(2)]
    return WeightSide(((batch, Ho * Wo, K), (K, Co)), pb1, cprime)


def conv2d(input: FixedPrecisionTensor, weight: FixedPrecisionTensor, bias=None, stride=1, padding=0, dilation=1,
           groups=1, prepared: WeightSide = None):
    """functional.py:204-308 + FPT.matmul truncation (precision.py:419-463), fused per party:

      delta_j = im2col(x_j) - a_j      (pm_spdz_mask_im2col_i64: no im2col tensor is materialised)
      eps_j   = w_j^T - b_j            (pm_spdz_mask_wt_i64)
      open    delta, eps               (peer read over NVLink)
      z_j     = [delta | a_j] @ [b_j (+eps) ; eps] + c_j   (pm_spdz_combine_matmul_i64)
      out_j   = NCHW( z_j / base**pf ) (pm_trunc_post_conv_i64)
    """
    assert groups == 1
    x, w = input.child, weight.child
    parties, provider = x.parties, x.provider
    B, C, H, W = x.shape
    Co, Ck, kh, kw = w.shape
    assert C == Ck
    Ho = ops.conv_out_size(H, kh, stride, padding, dilation)
    Wo = ops.conv_out_size(W, kw, stride, padding, dilation)
    M, K, N = Ho * Wo, C * kh * kw, Co
    shapes = ((B, M, K), (K, N))
    while True:
        try:
            tri = [p.crypto_store.get_keys(op="matmul", shapes=shapes, remove=False) for p in parties]
            break
        except EmptyCryptoPrimitiveStoreError as e:  # spdz.py:156-160
            if provider is None or any(p.crypto_store.force_preprocessing for p in parties):
                raise
            provider.provide_primitives(parties=parties, **e.kwargs_)
    d_sh = [ops.mask_im2col(x.child[j], tri[j][0], kh, kw, stride, padding, dilation) for j in range(2)]
    if prepared is not None:
        # online half only: delta is opened INTO its limb planes (never materialised), then one GEMM per party
        #   z_j = delta @ (b_j [+ eps]) + c'_j ,  c'_j = a_j @ eps + c_j from the offline phase
        assert prepared.shapes == _key(shapes), (prepared.shapes, shapes)
        pd = open_planes(parties, d_sh)
        z = []
        for j, p in enumerate(parties):
            p.crypto_store.get_keys(op="matmul", shapes=shapes, remove=True)      # the pop of spdz_compute (spdz.py:84)
            z.append(ops.gemm_planes(pd[j], prepared.pb1[j], None, None, prepared.cprime[j], B * M, K, N).view(B, M, N))
    else:
        e_sh = [ops.mask_wt(w.child[j].reshape(Co, K), tri[j][1]) for j in range(2)]
        delta = open_shares(parties, d_sh)
        eps = open_shares(parties, e_sh)
        z = [spdz_compute(p, j, delta[j], eps[j], "matmul") for j, p in enumerate(parties)]
    bsh = bias.child.child if bias is not None else [None, None]
    div = weight.base ** weight.precision_fractional
    out = [ops.trunc_post_conv(z[j], div, None, Ho, Wo) for j in range(2)]
    if bias is not None:  # bias is added after truncation (functional.py:185-192 runs on the truncated matmul result)
        out = [ops.axpby(1, out[j].permute(0, 2, 3, 1).contiguous(), 1, bsh[j]).permute(0, 3, 1, 2).contiguous()
               for j in range(2)]
    return FixedPrecisionTensor(x._new(out), input.base, input.precision_fractional)


class BNSide:
    """The image-independent half of one BatchNorm layer on shares, per party j (NCHW layout, see ``prepare_bn_side``)."""

    __slots__ = ("keys", "s1", "d1", "b1", "s2", "d2", "a2")
    TENSORS = ("s1", "d1", "b1", "s2", "d2", "a2")

    def __init__(self, keys, **kw):
        self.keys = keys
        for k, v in kw.items():
            setattr(self, k, v)


def prepare_bn_side(inv_std: FixedPrecisionTensor, weight: FixedPrecisionTensor, tri1, tri2, B, C, H, W):
    """batch_norm (functional.py:44-75) is two Beaver products with one operand that depends only on the model:

      normalized = inv_std * (flat - mean)      triple 1 = (a1 [C], b1 [P,C], c1 [P,C])
      result     = normalized * weight + bias   triple 2 = (a2 [P,C], b2 [C], c2 [P,C])

    With delta1 = open(inv_std - a1) and eps2 = open(weight - b2) known offline, spdz_compute (spdz.py:64-122) collapses to
      z1_j = s1_j[c] * eps1 + d1_j ,  s1_j = a1_j + [j=0] delta1 ,  d1_j = delta1 * b1_j + c1_j
      z2_j = s2_j[c] * delta2 + d2_j ,  s2_j = b2_j + [j=0] eps2 ,  d2_j = a2_j * eps2 + c2_j
    (exact in Z_2^64).  The [P,C] operands are stored NCHW so the online passes need no layout shuffle."""
    parties = inv_std.child.parties
    x, w = inv_std.child.child, weight.child.child
    delta1 = open_shares(parties, [ops.mask(x[j], tri1[j][0]) for j in range(2)])
    eps2 = open_shares(parties, [ops.mask(w[j], tri2[j][1]) for j in range(2)])
    nchw = lambda t: ops.pc_to_nchw(t, B, C, H, W)
    out = {k: [] for k in BNSide.TENSORS}
    for j in range(2):
        a1, b1, c1 = tri1[j]
        a2, b2, c2 = tri2[j]
        out["s1"].append(ops.axpby(1, a1, 1, delta1[j]) if j == 0 else a1)
        out["s2"].append(ops.axpby(1, b2, 1, eps2[j]) if j == 0 else b2)
        # combine_mul(1, delta, eps, a, b, c) = delta * b + a * eps + c with the [C] x [P,C] broadcasts
        out["d1"].append(nchw(ops.combine_mul(1, delta1[j], torch.zeros_like(b1), torch.zeros_like(a1), b1, c1)))
        out["d2"].append(nchw(ops.combine_mul(1, torch.zeros_like(a2), eps2[j], a2, torch.zeros_like(b2), c2)))
        out["b1"].append(nchw(b1))
        out["a2"].append(nchw(a2))
    P = B * H * W
    return BNSide((((C,), (P, C)), ((P, C), (C,))), **out)


def batch_norm_prepared(input: FixedPrecisionTensor, running_mean, bias, side: BNSide):
    """online half of batch_norm with ``side`` from the offline phase: three elementwise passes per party, NCHW throughout
         eps1_j   = x_j - mean_j[c] - b1_j
         delta2_j = trunc(s1_j[c] * open(eps1) + d1_j) - a2_j
         out_j    = trunc(s2_j[c] * open(delta2) + d2_j) + bias_j[c]
    -- the same shares as the op-by-op evaluation (tests/test_ring_gpu.py)."""
    from .spdz import peer_views, release_after_peer_reads

    ast = input.child
    parties = ast.parties
    B, C, H, W = ast.shape
    HW, D = H * W, input.base ** input.precision_fractional
    mean, bs = running_mean.child.child, bias.child.child
    for p in parties:                      # the two pops of spdz_compute (spdz.py:84)
        for key in side.keys:
            p.crypto_store.get_keys(op="mul", shapes=key, remove=True)
    e1 = [ops.spdz_affine(ast.child[j], None, None, None, 1, mean[j], -1, side.b1[j], -1, C, HW) for j in range(2)]
    pe = peer_views(parties, e1)
    d2 = [ops.spdz_affine(e1[j], pe[j], side.s1[j], side.d1[j], D, None, 0, side.a2[j], -1, C, HW) for j in range(2)]
    release_after_peer_reads(parties)
    pd = peer_views(parties, d2)
    out = [ops.spdz_affine(d2[j], pd[j], side.s2[j], side.d2[j], D, bs[j], 1, None, 0, C, HW) for j in range(2)]
    release_after_peer_reads(parties)
    return input._new(ast._new(out))


def batch_norm(input: FixedPrecisionTensor, running_mean, running_var, weight, bias, training=False,
               exponential_average_factor=0.0, eps=1e-5, inv_std=None):
    """functional.py:44-75 (eval branch: eps ignored, inverse sqrt by the 80-step Newton iteration).
    ``inv_std``: the result of ``running_var.reciprocal(method="newton")`` when the caller has already evaluated it
    (it depends on the model only; EncryptedResNet18 issues all layers' iterations as one launch)."""
    assert not training, "encrypted inference uses model.eval() (inference.py:288)"
    B, C, H, W = input.shape
    flat = input._new(input.child.map(ops.nchw_to_pc))
    x = inv_std if inv_std is not None else running_var.reciprocal(method="newton")
    normalized = x * (flat - running_mean)
    result = normalized * weight + bias
    return input._new(result.child.map(lambda s: ops.pc_to_nchw(s, B, C, H, W)))


def avg_pool2d(input: FixedPrecisionTensor, kernel_size, stride=None, padding=0):
    """functional.py:460-525 mode "avg" with kernel == stride, no padding (models.py:400-404)."""
    assert padding == 0 and (stride is None or stride == kernel_size)
    return input._new(input.child.map(lambda s: ops.avgpool(s, kernel_size)))


def relu(input: FixedPrecisionTensor, inplace=False):
    """torch.nn.functional.relu on FPT > AST: AdditiveSharingTensor.relu (additive_shared.py:896-897,922-925)"""
    return input.relu()


def max_pool2d(input: FixedPrecisionTensor, kernel_size=2, stride=2, padding=0, dilation=1, ceil_mode=None,
               return_indices=None):
    """functional.py:420-437 -> _pool2d :460-525, mode "max": _pre_pool per party, the max_half_split binary tree
    (one FSS comparison + one Beaver mul per step, on the AST: no truncation), _post_pool reshape."""
    from . import fss

    assert dilation == 1
    x = input.child
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * padding - kernel_size) // stride + 1, (W + 2 * padding - kernel_size) // stride + 1
    im = x.map(lambda s: fss.pre_pool(s, kernel_size, stride, padding))

    def select(left, right):
        return left + (right >= left) * (right - left)

    def max_half_split(t, L, half):
        return select(t.slice_lastdim(0, half), t.slice_lastdim(half, L - half))

    kk = kernel_size * kernel_size
    if kk == 4:
        res = max_half_split(max_half_split(im, 4, 2), 2, 1)
    elif kk == 9:
        res = select(im.slice_lastdim(0, 4), im.slice_lastdim(4, 4))    # max_half_split(im[..., :8], 4)
        res = max_half_split(res, 4, 2)
        left = max_half_split(res, 2, 1)
        res = select(left, im.slice_lastdim(8, 1))
    else:
        raise NotImplementedError("max_pool2d on shares: kernel 2 or 3 (the reference's fast paths, functional.py:501-508)")
    return input._new(res.reshape(B, C, Ho, Wo))


def linear(input: FixedPrecisionTensor, weight: FixedPrecisionTensor, bias: FixedPrecisionTensor = None):
    """functional.py:10-14 -> native_linear: input.matmul(weight.t()) + bias"""
    wt = weight._new(weight.child.map(lambda s: s.t().contiguous()))
    out = input.matmul(wt)
    if bias is not None:
        out = out + bias
    return out
