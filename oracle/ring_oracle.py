"""CPU oracle for path E (SPDZ additive-shared fixed-precision forward) -- TEST INFRASTRUCTURE ONLY.

This module is a torch-CPU int64 restatement of the reference's algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
it; the product (``primia_b200``) never does.  All file:line citations are relative to
``/root/reference``.

Parity pin: the reference ships no golden vectors.  The pieces of
the reference that are plain torch (``_pre_conv``, ``_post_conv``, ``triple_mat_mul``,
``build_triple``'s ``c = a @ b``, ``reciprocal(method="newton")`` control flow) were
executed *from the reference sources* in the build container by
``tests/golden/make_golden.py`` and their outputs are committed under ``tests/golden/``;
``tests/test_oracle_ring.py`` checks this restatement against them.  Two torch-1.4
semantics cannot be executed on torch 2.11 and are pinned by hand here (SURVEY.md
section 7 "Hard parts"): integer ``/`` == C truncation toward zero, and the fp32 scalar
promotion in ``fix_precision``.

Random streams (torch 1.4 ``random_``/``randint``) are not reproducible, so every share
and triple is an explicit input: parity == identical outputs for identical inputs.
"""
from __future__ import annotations

import torch

I64 = torch.int64
# which restatement of the FSS comparison the shared-tensor functions below evaluate with: the numpy + hashlib one
# (oracle/fss_oracle.py, default) or its C twin (oracle/fss_oracle_c.py: same results, pinned to the same fixtures, ~100x
# faster -- full-size 224 x 224 runs set ``ring_oracle.FSS = fss_oracle_c``)
FSS = None


# --------------------------------------------------------------------------- E2 / E14
def encode(x: torch.Tensor, base: int = 10, precision_fractional: int = 16) -> torch.Tensor:
    """FixedPrecisionTensor.fix_precision -- precision.py:117-132.

    ``(rational * base ** pf).long()``: a python-int scalar times an fp32 tensor is an
    fp32 multiply (the scalar is rounded to fp32 first), then ``.long()`` truncates
    toward zero.
    """
    assert x.dtype == torch.float32
    scale = torch.tensor(float(base ** precision_fractional), dtype=torch.float32)
    up = (x * scale).to(I64)
    return up


def decode(q: torch.Tensor, base: int = 10, precision_fractional: int = 16) -> torch.Tensor:
    """FixedPrecisionTensor.float_precision -- precision.py:134-144 (gate arithmetic is the identity)."""
    scale = torch.tensor(float(base ** precision_fractional), dtype=torch.float32)
    return q.to(torch.float32) / scale


# --------------------------------------------------------------------------- E3
def share_from_random(secret: torch.Tensor, s0: torch.Tensor):
    """AdditiveSharingTensor.generate_shares for n_workers == 2 -- additive_shared.py:336-365.

    ``s0`` is the explicit random tensor (reference: ``random_(min_value, max_value)``);
    shares = [s0, secret - s0] with native int64 wraparound.
    """
    return [s0.clone(), secret - s0]


def reconstruct(shares) -> torch.Tensor:
    """AdditiveSharingTensor.get -- additive_shared.py:287-301 (sum of shares, wraps mod 2**64)."""
    out = shares[0].clone()
    for s in shares[1:]:
        out = out + s
    return out


# --------------------------------------------------------------------------- E5 / E10
def pre_conv(x: torch.Tensor, w: torch.Tensor, stride=1, padding=0, dilation=1):
    """_pre_conv -- nn/functional.py:79-166 (groups == 1), vectorised.

    x: [B, C, H, W] int64, w: [Cout, C, kh, kw] int64.
    Returns (im [B, M=Ho*Wo, K=C*kh*kw], w_r [K, Cout], B, Cout, Ho, Wo).
    Column index k = ch*(kh*kw) + r*kw + c ; row index m = oh*Wo + ow (functional.py:129-149).
    """
    st = (stride, stride) if isinstance(stride, int) else tuple(stride)
    pd = (padding, padding) if isinstance(padding, int) else tuple(padding)
    dl = (dilation, dilation) if isinstance(dilation, int) else tuple(dilation)
    B, C, H, W = x.shape
    Co, Ck, kh, kw = w.shape
    assert C == Ck
    Ho = int(((H + 2 * pd[0] - dl[0] * (kh - 1) - 1) / st[0]) + 1)
    Wo = int(((W + 2 * pd[1] - dl[1] * (kw - 1) - 1) / st[1]) + 1)
    if pd != (0, 0):
        x = torch.nn.functional.pad(x, (pd[1], pd[1], pd[0], pd[0]), "constant")
        H += 2 * pd[0]
        W += 2 * pd[1]
    ch = torch.arange(C).view(C, 1, 1)
    r = torch.arange(kh).view(1, kh, 1)
    c = torch.arange(kw).view(1, 1, kw)
    # NB the reference uses dilation[0] for the row term and dilation[1] for the column term
    pattern = (r * W * dl[0] + c * dl[1] + ch * H * W).reshape(-1)  # [K]
    oh = torch.arange(Ho).view(Ho, 1)
    ow = torch.arange(Wo).view(1, Wo)
    offset = (oh * st[0] * W + ow * st[1]).reshape(-1)  # [M]
    idx = offset.view(-1, 1) + pattern.view(1, -1)  # [M, K]
    im = x.reshape(B, -1)[:, idx]  # [B, M, K]
    w_r = w.reshape(Co, -1).t()
    return im, w_r, B, Co, Ho, Wo


def post_conv(bias, res: torch.Tensor, B, Co, Ho, Wo):
    """_post_conv -- nn/functional.py:170-201."""
    if bias is not None:
        res = res + bias
    return res.permute(0, 2, 1).reshape(B, Co, Ho, Wo).contiguous()


# --------------------------------------------------------------------------- E7 / E8
def build_triple_c(a: torch.Tensor, b: torch.Tensor, op: str) -> torch.Tensor:
    """c = a (op) b from build_triple -- beaver.py:32-52 (a, b are explicit inputs)."""
    return torch.matmul(a, b) if op == "matmul" else a * b


def spdz_mask(x_j, y_j, a_j, b_j):
    """spdz_mask -- spdz.py:22-45: (x - a, y - b) on party j."""
    return x_j - a_j, y_j - b_j


def spdz_compute(j: int, delta, epsilon, a_j, b_j, c_j, op: str):
    """spdz_compute -- spdz.py:64-122: delta(op)b + a(op)eps + c (+ delta(op)eps on party 0).

    The reference evaluates ``delta_epsilon + delta_b + a_epsilon + c`` (j == 0) or
    ``delta_b + a_epsilon + c``; int64 addition is associative mod 2**64 so order is moot.
    """
    f = torch.matmul if op == "matmul" else torch.mul
    delta_b = f(delta, b_j)
    a_eps = f(a_j, epsilon)
    if j == 0:
        return f(delta, epsilon) + delta_b + a_eps + c_j
    return delta_b + a_eps + c_j


def spdz_mul(op: str, x_sh, y_sh, triple_sh):
    """spdz_mul -- spdz.py:125-197 for 2 parties.

    x_sh, y_sh: [share0, share1]; triple_sh: [(a0,b0,c0), (a1,b1,c1)].
    Returns the two output shares (before any truncation).
    """
    d, e = [], []
    for j in range(2):
        dj, ej = spdz_mask(x_sh[j], y_sh[j], triple_sh[j][0], triple_sh[j][1])
        d.append(dj)
        e.append(ej)
    delta = d[0] + d[1]  # spdz.py:162
    epsilon = e[0] + e[1]  # spdz.py:163
    return [spdz_compute(j, delta, epsilon, *triple_sh[j], op) for j in range(2)]


# --------------------------------------------------------------------------- E9
def trunc_div(share: torch.Tensor, divisor: int) -> torch.Tensor:
    """Per-share public division -- additive_shared.py:673-678 with torch-1.4 integer ``/``
    (C truncation toward zero); reached from FixedPrecisionTensor.truncate precision.py:146-154."""
    return torch.div(share, divisor, rounding_mode="trunc")


def truncate(shares, base: int, precision_fractional: int):
    return [trunc_div(s, base ** precision_fractional) for s in shares]


# --------------------------------------------------------------------------- E4 / E6
def conv2d_shared(x_sh, w_sh, triple_sh, stride, padding, base, pf, dilation=1):
    """conv2d on shares -- nn/functional.py:204-308 followed by FPT.matmul's truncate
    (precision.py:419-463).  Returns per-party NCHW output shares.
    triple_sh[j] = (a_j [B,M,K], b_j [K,N], c_j [B,M,N])."""
    pre = [pre_conv(x_sh[j], w_sh[j], stride, padding, dilation) for j in range(2)]
    z = spdz_mul("matmul", [pre[0][0], pre[1][0]], [pre[0][1], pre[1][1]], triple_sh)
    z = truncate(z, base, pf)
    return [post_conv(None, z[j], *pre[j][2:]) for j in range(2)]


def mul_shared(x_sh, y_sh, triple_sh, base, pf):
    """FPT.mul for FPT>AST operands -- precision.py:264-366 (private mul then truncate)."""
    z = spdz_mul("mul", x_sh, y_sh, triple_sh)
    return truncate(z, base, pf)


# --------------------------------------------------------------------------- E11
def public_sub_shared(x_sh, const_sh):
    """AST.sub(int) -- additive_shared.py:455-487: the public operand is *secret-shared*
    (fresh randomness => explicit ``const_sh`` input) and subtracted share-wise."""
    return [x_sh[j] - const_sh[j] for j in range(2)]


def rsub_const(x_sh, const_sh):
    """FPT.__rsub__ -- precision.py:240-241: ``(self - other) * -1``; the ``* -1`` is a public
    mul per share (additive_shared.py:561-588)."""
    return [(x_sh[j] - const_sh[j]) * -1 for j in range(2)]


def newton_inv_sqrt_like(v_sh, consts_sh, triples, base, pf, iters=80):
    """reciprocal(method="newton") -- precision.py:507-518, literally.

    x0 = (C+1 - v)/C ; then 79x:  y = C+1 - v*(x*x) ; x = y*x/C   with C = 20.
    consts_sh[i]  : explicit 2-party sharing of encode(21) used at iteration i.
    triples[i]    : for i >= 1, three elementwise triples (for x*x, v*(x*x), y*x).
    ``/ C`` is FPT.__truediv__ by a python int -> per-share trunc_div (precision.py:264-366
    cmd == "div" with int other: no rescale, no truncate; additive_shared.py:673-678).
    """
    C = 20
    x = None
    for i in range(iters):
        if x is not None:
            xx = mul_shared(x, x, triples[i][0], base, pf)
            vxx = mul_shared(v_sh, xx, triples[i][1], base, pf)
            y = rsub_const(vxx, consts_sh[i])
            yx = mul_shared(y, x, triples[i][2], base, pf)
            x = [trunc_div(s, C) for s in yx]
        else:
            y = rsub_const(v_sh, consts_sh[i])
            x = [trunc_div(s, C) for s in y]
    return x


def batch_norm_eval_shared(x_sh, mean_sh, var_sh, gamma_sh, beta_sh, consts_sh, triples, tri_norm,
                           tri_affine, base, pf, iters=80):
    """batch_norm (eval) -- nn/functional.py:44-75.  x_sh: per-party [B,C,H,W].

    input -> [B*H*W (n-major within channel), C]; x = newton(var); normalized = x*(input-mean);
    result = normalized*weight + bias; reshape back.  NB eps is ignored in eval (functional.py:62-64).
    tri_norm / tri_affine: elementwise triples with broadcast shapes ([C] x [P,C]) and ([P,C] x [C]).
    """
    B, C, H, W = x_sh[0].shape
    flat = [s.permute(1, 0, 2, 3).reshape(C, -1).t() for s in x_sh]  # [P, C]
    inv = newton_inv_sqrt_like(var_sh, consts_sh, triples, base, pf, iters)
    centered = [flat[j] - mean_sh[j] for j in range(2)]
    normalized = mul_shared(inv, centered, tri_norm, base, pf)
    res = mul_shared(normalized, gamma_sh, tri_affine, base, pf)
    res = [res[j] + beta_sh[j] for j in range(2)]
    return [r.t().reshape(C, B, H, W).permute(1, 0, 2, 3).contiguous() for r in res]


# --------------------------------------------------------------------------- E13
def avg_pool_shared(x_sh, k: int):
    """avg_pool2d with kernel == stride == k, no padding -- nn/functional.py:460-525 mode "avg":
    windows are summed per share and divided by k*k via AST.mean (additive_shared.py:720-729:
    ``share.sum(dim) / m`` with integer trunc)."""
    out = []
    for s in x_sh:
        B, C, H, W = s.shape
        w = s.reshape(B, C, H // k, k, W // k, k).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // k, W // k, k * k)
        out.append(trunc_div(w.sum(-1), k * k))
    return out


def linear_shared(x_sh, w_sh, b_sh, triple_sh, base, pf):
    """linear -- nn/functional.py:10-14 -> native_linear: x.matmul(w.t()) + b  (Beaver matmul + truncate)."""
    wt = [w.t() for w in w_sh]
    z = spdz_mul("matmul", x_sh, wt, triple_sh)
    z = truncate(z, base, pf)
    return [z[j] + b_sh[j] for j in range(2)]


# --------------------------------------------------------------------------- E12 / E13 (comparison-based ops, protocol="fss")
def fss_le_shared(x1_sh, x2_sh, fss_key, alpha_sh):
    """fss.le(x1, x2) -- syft/frameworks/torch/mpc/fss.py:97-185,279-283: shares of [x1 <= x2] (0/1, unscaled).
    fss_key / alpha_sh: explicit DIF key material for x.numel() instances (oracle/fss_oracle.py layout)."""
    if FSS is None:
        from . import fss_oracle as F
    else:
        F = FSS

    shape = x1_sh[0].shape
    out = F.fss_le([t.reshape(-1).numpy() for t in x1_sh], [t.reshape(-1).numpy() for t in x2_sh], fss_key, alpha_sh)
    return [torch.from_numpy(o.astype("int64")).reshape(shape) for o in out]


def relu_shared(x_sh, fss_key, alpha_sh, triple_sh):
    """AdditiveSharingTensor.relu, protocol "fss" -- additive_shared.py:922-925: ``zero = self - self;
    self * (self >= zero)`` with ``>=`` = fss.le(zero, self) (:950-952) and ``*`` an elementwise Beaver mul
    (no truncation: the comparison result is an unscaled 0/1)."""
    zero = [s - s for s in x_sh]
    c = fss_le_shared(zero, x_sh, fss_key, alpha_sh)
    return spdz_mul("mul", x_sh, c, triple_sh)


def pre_pool(x: torch.Tensor, k: int, stride: int, padding: int):
    """_pre_pool -- nn/functional.py:312-390 (dilation 1), vectorised: [B,C,H,W] -> [B,C,M,k*k], zero padding."""
    B, C, H, W = x.shape
    Ho = int(((H + 2 * padding - (k - 1) - 1) / stride) + 1)
    Wo = int(((W + 2 * padding - (k - 1) - 1) / stride) + 1)
    if padding:
        x = torch.nn.functional.pad(x, (padding, padding, padding, padding), "constant")
        H, W = H + 2 * padding, W + 2 * padding
    pattern = (torch.arange(k).view(k, 1) * W + torch.arange(k).view(1, k)).reshape(-1)
    offset = (torch.arange(Ho).view(Ho, 1) * stride * W + torch.arange(Wo).view(1, Wo) * stride).reshape(-1)
    idx = offset.view(-1, 1) + pattern.view(1, -1)
    return x.reshape(B, C, -1)[:, :, idx], B, C, Ho, Wo


def max_pool2d_shared(x_sh, k, stride, padding, fss_keys, alpha_shs, triples, trace=None):
    """max_pool2d on shares -- nn/functional.py:420-437,460-525 mode "max" for k*k in (4, 9): the binary-tree
    ``max_half_split`` (left + (right >= left) * (right - left)), one FSS comparison + one Beaver mul per step.
    fss_keys/alpha_shs/triples: one entry per step, in execution order."""
    pre = [pre_pool(s, k, stride, padding) for s in x_sh]
    im = [p[0] for p in pre]
    _, B, C, Ho, Wo = pre[0]
    step = iter(range(len(triples)))

    def select(left, right):
        i = next(step)
        if trace is not None:
            trace.append(tuple(left[0].shape))
        c = fss_le_shared(left, right, fss_keys[i], alpha_shs[i])          # right >= left
        diff = [right[j] - left[j] for j in range(2)]
        prod = spdz_mul("mul", c, diff, triples[i])
        return [left[j] + prod[j] for j in range(2)]

    def max_half_split(t, half):
        return select([s[..., :half] for s in t], [s[..., half:] for s in t])

    if im[0].shape[-1] == 4:
        res = max_half_split(max_half_split(im, 2), 1)
    elif im[0].shape[-1] == 9:
        res = max_half_split([s[..., :8] for s in im], 4)
        res = max_half_split(res, 2)
        left = max_half_split(res, 1)
        res = select(left, [s[..., 8:] for s in im])
    else:
        raise NotImplementedError("the reference falls back to AST.max for other kernels; ResNet-18 uses 3x3")
    return [r.reshape(B, C, Ho, Wo).contiguous() for r in res]


# --------------------------------------------------------------------------- E1: the whole encrypted forward
class Tape:
    """Explicit randomness of one encrypted forward.  ``triples`` [(op, [(a0,b0,c0),(a1,b1,c1)])] are consumed first-in
    first-out PER (op, operand shapes) -- the reference's crypto store is keyed that way (primitives.py:52-102) -- so only the
    relative order of same-shaped triples matters; ``consts`` [[s0, s1]] (sharings of public constants,
    additive_shared.py:473-487) and ``fss`` [(key dict, [alpha0, alpha1])] (oracle/fss_oracle.py layout) are plain FIFOs."""

    def __init__(self, triples, consts, fss):
        self.triples = {}
        for op, tri in triples:
            key = (op, tuple(tri[0][0].shape), tuple(tri[0][1].shape))
            self.triples.setdefault(key, []).append(tri)
        self.consts, self.fss = list(consts), list(fss)

    def triple(self, op, x_shape, y_shape):
        return self.triples[(op, tuple(x_shape), tuple(y_shape))].pop(0)

    def const(self):
        return self.consts.pop(0)

    def fss_keys(self, n):
        key, alpha = self.fss.pop(0)
        assert key["leaf"].shape[1] == n, (key["leaf"].shape, n)
        return key, alpha

    def exhausted(self):
        return not (any(self.triples.values()) or self.consts or self.fss)


def batch_norm_eval_taped(x_sh, mean_sh, var_sh, gamma_sh, beta_sh, tape: Tape, base, pf):
    """batch_norm (eval, functional.py:44-75) drawing its randomness from the tape in the order the reference's
    expression evaluation consumes it: iteration 0: const ; iterations 1..79: x*x, v*(xx), const, y*x ; then the two muls."""
    C = var_sh[0].shape[0]
    consts, triples = [], []
    for i in range(80):
        if i == 0:
            consts.append(tape.const())
            triples.append(None)
        else:
            t0, t1 = tape.triple("mul", (C,), (C,)), tape.triple("mul", (C,), (C,))
            consts.append(tape.const())
            triples.append((t0, t1, tape.triple("mul", (C,), (C,))))
    B, _, H, W = x_sh[0].shape
    P = B * H * W
    tri_norm, tri_affine = tape.triple("mul", (C,), (P, C)), tape.triple("mul", (P, C), (C,))
    return batch_norm_eval_shared(x_sh, mean_sh, var_sh, gamma_sh, beta_sh, consts, triples, tri_norm, tri_affine, base, pf)


def relu_taped(x_sh, tape: Tape):
    key, alpha = tape.fss_keys(x_sh[0].numel())
    return relu_shared(x_sh, key, alpha, tape.triple("mul", x_sh[0].shape, x_sh[0].shape))


def max_pool2d_taped(x_sh, k, stride, padding, tape: Tape):
    B, C, H, W = x_sh[0].shape
    M = ((H + 2 * padding - k) // stride + 1) * ((W + 2 * padding - k) // stride + 1)
    widths = {4: [2, 1], 9: [4, 2, 1, 1]}[k * k]
    keys, alphas, tris = [], [], []
    for w in widths:
        key, alpha = tape.fss_keys(B * C * M * w)
        keys.append(key)
        alphas.append(alpha)
        tris.append(tape.triple("mul", (B, C, M, w), (B, C, M, w)))
    return max_pool2d_shared(x_sh, k, stride, padding, keys, alphas, tris)


def resnet18_forward_shared(P, x_sh, tape: Tape, base=10, pf=16, input_size=224, taps=None):
    """inference.py:288-314 on shares: ResNet._forward_impl (torchlib/models.py:466-482) with ``model.pool`` and
    ``model.relu`` swapped (inference.py:289: the stem runs conv -> bn -> MAX-POOL -> RELU), BasicBlock.forward
    (:268-284), AvgPool2d(input_size/32), flatten, fc.  P: {state_dict key: [share0, share1]} of the encoded
    parameters and BN buffers.  taps (optional dict) receives named intermediate shares."""

    def conv(x, name, stride, pad):
        w = P[name + ".weight"]
        B, C, H, W = x[0].shape
        Co, _, kh, kw = w[0].shape
        Ho = (H + 2 * pad - kh) // stride + 1
        tri = tape.triple("matmul", (B, Ho * Ho, C * kh * kw), (C * kh * kw, Co))
        return conv2d_shared(x, w, tri, stride, pad, base, pf)

    def bn(x, name):
        return batch_norm_eval_taped(x, P[name + ".running_mean"], P[name + ".running_var"], P[name + ".weight"],
                                     P[name + ".bias"], tape, base, pf)

    def tap(name, x):
        if taps is not None:
            taps[name] = [s.clone() for s in x]
        return x

    x = tap("conv1", conv(x_sh, "conv1", 2, 3))
    x = tap("bn1", bn(x, "bn1"))
    x = tap("pool", max_pool2d_taped(x, 3, 2, 1, tape))
    x = tap("relu", relu_taped(x, tape))
    inplanes = 64
    for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi in range(2):
            st = stride if bi == 0 else 1
            pre = f"layer{li}.{bi}"
            identity = x
            out = conv(x, pre + ".conv1", st, 1)
            out = bn(out, pre + ".bn1")
            out = relu_taped(out, tape)
            out = conv(out, pre + ".conv2", 1, 1)
            out = bn(out, pre + ".bn2")
            if st != 1 or inplanes != planes:
                identity = bn(conv(x, pre + ".downsample.0", st, 0), pre + ".downsample.1")
            out = [out[j] + identity[j] for j in range(2)]
            x = tap(pre, relu_taped(out, tape))
            inplanes = planes
    x = avg_pool_shared(x, input_size // 32)
    x = [s.reshape(s.shape[0], -1) for s in x]
    B = x[0].shape[0]
    ncls = P["fc.weight"][0].shape[0]
    return linear_shared(x, P["fc.weight"], P["fc.bias"], tape.triple("matmul", (B, 512), (512, ncls)), base, pf)


class GeneratingTape(Tape):
    """a tape that draws its randomness on demand (numpy), for oracle-only end-to-end runs: the CPU tests and the CPU
    baseline leg of bench.py.  ``const_value``: the encoded public constant of the Newton iteration ((C+1) * base**pf)."""

    def __init__(self, seed=0, const_value=0):
        import numpy as np

        self.rng = np.random.default_rng(seed)
        self.const_value = const_value
        self.n_triples = self.n_consts = self.n_fss = 0
        self.gen_seconds = 0.0  # time spent producing primitives (the offline phase), so callers can time the online part

    def _r(self, shape):
        return torch.from_numpy(self.rng.integers(-2 ** 63, 2 ** 63 - 1, tuple(shape)))

    def triple(self, op, x_shape=None, y_shape=None):
        import time

        t0 = time.perf_counter()
        self.n_triples += 1
        a, b = self._r(x_shape), self._r(y_shape)
        c = torch.matmul(a, b) if op == "matmul" else a * b
        a0, b0, c0 = self._r(a.shape), self._r(b.shape), self._r(c.shape)
        out = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
        self.gen_seconds += time.perf_counter() - t0
        return out

    def const(self):
        self.n_consts += 1
        s0 = self._r((1,))
        return [s0, torch.tensor([self.const_value], dtype=torch.int64) - s0]

    def fss_keys(self, n):
        import numpy as np

        if FSS is None:
            from . import fss_oracle as F
        else:
            F = FSS

        import time

        t0 = time.perf_counter()
        self.n_fss += n
        alpha = self.rng.integers(0, 2 ** 32, n, dtype=np.uint64)
        key = F.dif_keygen(alpha, self.rng.integers(0, 2 ** 63, (2, 2, n), dtype=np.uint64))
        out = key, F.split_alpha(alpha, self.rng.integers(0, 2 ** 32, n, dtype=np.uint64))
        self.gen_seconds += time.perf_counter() - t0
        return out

    def exhausted(self):
        return True
