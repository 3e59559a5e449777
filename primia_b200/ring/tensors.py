"""AdditiveSharingTensor / FixedPrecisionTensor on GPU shares.

Mirrors the subset of syft/frameworks/torch/tensors/interpreters/{additive_shared,precision}.py that
inference.py:279-321 exercises."""
from __future__ import annotations

import os

import torch

from . import ops
from .spdz import TripleProvider, spdz_mul


class ShareRNG:
    """Source of the fresh randomness the reference draws when it secret-shares a value
    (additive_shared.py:336-365).  Tests subclass it to replay explicit shares.

    ``mode``: "live" draws a new Philox offset per call; "record" additionally keeps (q, s0, s1) of every call as static
    buffers; "replay" hands those buffers out again in order (a captured CUDA graph reads them; ``refresh_static`` rewrites
    them with fresh randomness between replays)."""

    def __init__(self, seed=0xA11CE):
        self.seed = seed
        self.counter = 0
        self.mode = "live"
        self.static, self.cursor = [], 0
        self.epoch = None   # device-side Philox epoch (ops.new_epoch), created with the first sharing

    def _ep(self, q):
        if self.epoch is None or self.epoch.device != q.device:
            self.epoch = ops.new_epoch(q.device)
        return self.epoch

    def bump_epoch(self):
        if self.epoch is not None:
            ops.bump_epoch(self.epoch)

    def share(self, q: torch.Tensor):
        if self.mode == "replay":
            _q, s0, s1 = self.static[self.cursor]
            self.cursor += 1
            return s0, s1
        self.counter += 1
        s0, s1 = ops.share_gen(q, self.seed, self.counter, epoch=self._ep(q))
        if self.mode == "record":
            self.static.append((q, s0, s1))
        return s0, s1

    def refresh_static(self):
        for q, s0, s1 in self.static:
            self.counter += 1
            ops.share_gen(q, self.seed, self.counter, out=(s0, s1), epoch=self._ep(q))   # straight into the static buffers
        self.cursor = 0


DEFAULT_RNG = ShareRNG()  # one stream per process: every fresh sharing draws a new Philox offset


def provider_of(crypto_provider, seed=None):
    """``crypto_provider=`` of the PySyft verbs is a WORKER (inference.py:281-285); the object that generates primitives on
    that worker's GPU is created once per worker and reused, so its Philox counter never restarts."""
    if crypto_provider is None or isinstance(crypto_provider, TripleProvider):
        return crypto_provider
    prov = getattr(crypto_provider, "_triple_provider", None)
    if prov is None:
        if seed is None:
            import secrets

            seed = secrets.randbits(63)
        prov = TripleProvider(crypto_provider, seed)
        crypto_provider._triple_provider = prov
    return prov


class AdditiveSharingTensor:
    """2-party additive sharing over Z_2^64; ``child[j]`` is party j's share on party j's GPU."""

    def __init__(self, shares, parties, provider: TripleProvider = None, rng: ShareRNG = None):
        self.child = list(shares)
        self.parties = list(parties)
        self.provider = provider
        self.rng = rng or DEFAULT_RNG

    # -- construction / reconstruction
    @classmethod
    def share_secret(cls, q: torch.Tensor, parties, provider=None, rng=None):
        """additive_shared.py:317-365"""
        rng = rng or DEFAULT_RNG
        s0, s1 = rng.share(q)
        shares = [s0.to(parties[0].device), s1.to(parties[1].device)]
        return cls(shares, parties, provider, rng)

    def get(self):
        """additive_shared.py:287-301"""
        a, b = self.child
        return ops.open_add(a, b.to(a.device))

    @property
    def shape(self):
        return self.child[0].shape

    def _new(self, shares):
        return AdditiveSharingTensor(shares, self.parties, self.provider, self.rng)

    # -- linear ops (additive_shared.py:455-527)
    def _shared_const(self, value: int):
        q = torch.tensor([value], dtype=torch.int64, device=self.parties[0].device)
        s0, s1 = self.rng.share(q)
        return [s0, s1.to(self.parties[1].device)]

    def add(self, other):
        if isinstance(other, int):
            other = self._new(self._shared_const(other))
        return self._new([ops.axpby(1, s, 1, o) for s, o in zip(self.child, other.child)])

    def sub(self, other):
        if isinstance(other, int):
            other = self._new(self._shared_const(other))
        return self._new([ops.axpby(1, s, -1, o) for s, o in zip(self.child, other.child)])

    __add__ = add
    __sub__ = sub

    def public_mul(self, k: int):
        """_public_mul additive_shared.py:561-588 (k != 0)"""
        assert k != 0, "multiplying by public 0 needs a zero-refresh (additive_shared.py:27-60); not on the hot path"
        return self._new([ops.axpby(k, s) for s in self.child])

    def public_div(self, d: int):
        """_public_div additive_shared.py:673-678"""
        return self._new([ops.trunc_div(s, d) for s in self.child])

    # -- private multiplications (additive_shared.py:529-557,643-654)
    def mul(self, other):
        if isinstance(other, int):
            return self.public_mul(other)
        return self._new(spdz_mul("mul", self.child, other.child, self.parties, self.provider))

    def matmul(self, other):
        return self._new(spdz_mul("matmul", self.child, other.child, self.parties, self.provider))

    def map(self, fn):
        return self._new([fn(s) for s in self.child])

    def numel(self):
        return self.child[0].numel()

    # -- comparisons, protocol "fss" (additive_shared.py:939-974): shares of an unscaled 0/1
    def __le__(self, other):
        from . import fss

        return self._new(fss.le(self.child, other.child, self.parties, self.provider))

    def __ge__(self, other):
        from . import fss

        return self._new(fss.le(other.child, self.child, self.parties, self.provider))

    def __gt__(self, other):
        return (other + 1) <= self

    def __lt__(self, other):
        return (self + 1) <= other

    def relu(self):
        """additive_shared.py:922-925"""
        from . import fss, spdz

        if spdz.FUSE_OPEN:
            # zero = self - self is, share by share, exactly 0: mask_builder takes it as the public-0 operand (fss.py:189-204)
            return self * self._new(fss.le([None, None], self.child, self.parties, self.provider))
        zero = self - self
        return self * (self >= zero)

    __mul__ = mul

    def slice_lastdim(self, start, length):
        """t[..., start:start+length] on every share"""
        from . import fss

        return self._new([fss.slice_lastdim(s, start, length) for s in self.child])

    def reshape(self, *shape):
        return self._new([s.reshape(*shape) for s in self.child])


class FixedPrecisionTensor:
    """precision.py: value = child / base**precision_fractional ; child is an AST (private) or an int64 tensor (public)."""

    def __init__(self, child, base=10, precision_fractional=16):
        self.child = child
        self.base = base
        self.precision_fractional = precision_fractional

    @property
    def scale(self):
        return self.base ** self.precision_fractional

    @classmethod
    def fix_precision(cls, x: torch.Tensor, base=10, precision_fractional=16):
        return cls(ops.encode(x, base, precision_fractional), base, precision_fractional)

    def share(self, *parties, crypto_provider=None, rng=None, **_):
        """precision.py:910-957 -> native.share native.py:887-949"""
        ast = AdditiveSharingTensor.share_secret(self.child, list(parties), provider_of(crypto_provider), rng)
        return FixedPrecisionTensor(ast, self.base, self.precision_fractional)

    def get(self):
        return FixedPrecisionTensor(self.child.get(), self.base, self.precision_fractional)

    def float_precision(self):
        return ops.decode(self.child, self.base, self.precision_fractional)

    float_prec = float_precision

    def _new(self, child):
        return FixedPrecisionTensor(child, self.base, self.precision_fractional)

    @property
    def shape(self):
        return self.child.shape

    def truncate(self, pf):
        """precision.py:146-160"""
        return self._new(self.child.public_div(self.base ** pf))

    # precision.py:180-254
    def __add__(self, other):
        if isinstance(other, int):
            return self._new(self.child.add(int(other * self.scale)))
        return self._new(self.child.add(other.child))

    def __sub__(self, other):
        if isinstance(other, int):
            return self._new(self.child.sub(int(other * self.scale)))
        return self._new(self.child.sub(other.child))

    def __rsub__(self, other):
        return (self - other) * -1

    # precision.py:264-366
    def __mul__(self, other):
        if isinstance(other, int):
            return self._new(self.child.mul(other))
        return self._new(self.child.mul(other.child)).truncate(self.precision_fractional)

    def __truediv__(self, other):
        assert isinstance(other, int)
        return self._new(self.child.public_div(other))

    # precision.py:419-463
    def matmul(self, other):
        return self._new(self.child.matmul(other.child)).truncate(other.precision_fractional)

    def relu(self):
        """F.relu on an FPT is forwarded to the child (precision.py:864-905 finds no FPT override): no rescale, no truncation"""
        return self._new(self.child.relu())

    def reciprocal(self, method="newton"):
        """precision.py:507-518 -- literally (80 iterations, C = 20).  When both share holders are resident on one GPU the
        whole iteration runs as ONE kernel (pm_bn_newton_fused_i64) with identical per-party arithmetic and the same
        consumption of triples and constant sharings; when they sit on two GPUs it is one kernel per party exchanging the
        openings over NVLink (pm_bn_newton_p2p_i64); PRIMIA_FUSE_NEWTON=0 issues the protocol op by op."""
        assert method == "newton"
        ast = self.child
        if FUSE_NEWTON and len(ast.child[0].shape) == 1:
            return self._reciprocal_fused(80, 20)
        x = None
        C = 20
        for _ in range(80):
            if x is not None:
                y = C + 1 - self * (x * x)
                x = y * x / C
            else:
                y = C + 1 - self
                x = y / C
        return x

    def _reciprocal_fused(self, iters, C):
        return reciprocal_newton_batched([self], iters, C)[0]


def reciprocal_newton_batched(fpts, iters=80, C=20):
    """reciprocal(method="newton") of several shared vectors (one per BatchNorm layer) in ONE launch.  Triples and constant
    sharings are taken from the stores / the share RNG in list order, exactly as the sequential evaluation would."""
    from .spdz import take_primitives

    jobs = []
    for f in fpts:
        ast = f.child
        n = ast.child[0].shape[0]
        tri = take_primitives("mul", ((n,), (n,)), 3 * (iters - 1), ast.parties, ast.provider)
        dev = ast.parties[0].device
        q = torch.full((iters,), int((C + 1) * f.scale), dtype=torch.int64, device=dev)
        k = ast.rng.share(q)  # the constant is freshly shared at every iteration (additive_shared.py:473-487)
        k = (k[0], k[1].to(ast.parties[1].device))
        packed = [[ops.stack([t[i] for t in tri[j]]) for i in range(3)] for j in range(2)]
        jobs.append((ast.child, packed, k))
    scales = {f.scale for f in fpts}
    assert len(scales) == 1
    scale = scales.pop()
    parties = fpts[0].child.parties
    if parties[0].device != parties[1].device:
        xs = []
        for base in range(0, len(jobs), 32):
            xs += ops.bn_newton_p2p(jobs[base:base + 32], iters, scale, C)
    else:
        xs = ops.bn_newton_fused(jobs, iters, scale, C)
    return [f._new(f.child._new(x)) for f, x in zip(fpts, xs)]


FUSE_NEWTON = os.environ.get("PRIMIA_FUSE_NEWTON", "1") != "0"
