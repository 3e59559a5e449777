"""Function-secret-sharing comparison on GPU shares (protocol="fss"): keys, store, ``le``.

Mirrors syft/frameworks/torch/mpc/fss.py (fss_op :97-185, mask_builder :189-204, evaluate :208-245, DIF :341-428) and
the FSS half of the crypto store (primitives.py:52-102 get_keys, :237-253 build_fss_keys).  One thread of
``pm_fss_dif_keygen`` / ``pm_fss_dif_eval`` owns one comparison instance (32 SHA-512 compressions per evaluation)."""
from __future__ import annotations

import ctypes

import torch

from .._lib import PrimiaError, call, ptr, stream
from . import ops

N_BITS = 32  # fss.py:27
OP = "fss_comp"


def _p(t):
    """device pointer of a (possibly column-sliced) tensor"""
    if not t.is_cuda:
        raise PrimiaError("primia_b200 FSS kernels take CUDA tensors only (no CPU fallback)")
    return ctypes.c_void_p(t.data_ptr())


class FSSKeys:
    """Party b's DIF keys for N comparison instances (structure of arrays, see include/primia_b200.h).

    alpha [N] int64 (this party's additive share mod 2^32 of the secret offsets), s0 [2,N], bits [32,N] uint8,
    sigma_cw [32,2,N], s_cw [32,2,N] (uint64 bit patterns held in int64 tensors), leaf [33,N] int32.
    ``off`` is the consumption cursor (get_keys(remove=True) burns instances from the front, primitives.py:84-100)."""

    def __init__(self, alpha, s0, bits, sigma_cw, s_cw, leaf):
        self.alpha, self.s0, self.bits, self.sigma_cw, self.s_cw, self.leaf = alpha, s0, bits, sigma_cw, s_cw, leaf
        self.N = int(alpha.shape[0])
        self.off = 0

    @property
    def available(self):
        return self.N - self.off

    @property
    def device(self):
        return self.alpha.device

    def tensors(self):
        return (self.alpha, self.s0, self.bits, self.sigma_cw, self.s_cw, self.leaf)

    def window(self, n):
        """the next n instances as column views (row stride stays N)"""
        o = self.off
        return FSSWindow(self.alpha[o:o + n], self.s0[:, o:o + n], self.bits[:, o:o + n], self.sigma_cw[:, :, o:o + n],
                         self.s_cw[:, :, o:o + n], self.leaf[:, o:o + n], n, self.N)

    def to(self, device):
        return FSSKeys(*(t.to(device, non_blocking=True) for t in self.tensors()))

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors())

    @staticmethod
    def cat(chunks):
        """concatenate the unconsumed parts of several pools (primitives.py:213-233 appends along the instance axis)"""
        parts = [[t[..., c.off:] for t in c.tensors()] for c in chunks]
        return FSSKeys(*(torch.cat([p[i] for p in parts], dim=-1).contiguous() for i in range(6)))


class FSSWindow:
    def __init__(self, alpha, s0, bits, sigma_cw, s_cw, leaf, n, stride):
        self.alpha, self.s0, self.bits, self.sigma_cw, self.s_cw, self.leaf = alpha, s0, bits, sigma_cw, s_cw, leaf
        self.n, self.stride = n, stride


# ------------------------------------------------------------------------------------------ kernels
def prg_sha512(seed):
    """H's hash (fss.py:581-586): seed [2,n] -> [8,n] little-endian digest words"""
    seed = seed.contiguous()
    n = seed.shape[1]
    out = torch.empty((8, n), dtype=torch.int64, device=seed.device)
    with torch.cuda.device(seed.device):
        call("pm_fss_prg_sha512", ptr(seed), n, ptr(out), stream())
    return out


def dif_keygen(alpha, seeds):
    """DIF.keygen (fss.py:341-399) from explicit randomness.  alpha [n] (int64 holding values < 2^32), seeds [2,2,n].
    Returns (bits [32,n] uint8, sigma_cw [32,2,n], s_cw [32,2,n], leaf [33,n] int32)."""
    alpha, seeds = alpha.contiguous(), seeds.contiguous()
    n, dev = alpha.shape[0], alpha.device
    bits = torch.empty((N_BITS, n), dtype=torch.uint8, device=dev)
    sigma_cw = torch.empty((N_BITS, 2, n), dtype=torch.int64, device=dev)
    s_cw = torch.empty((N_BITS, 2, n), dtype=torch.int64, device=dev)
    leaf = torch.empty((N_BITS + 1, n), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        call("pm_fss_dif_keygen", ptr(alpha), ptr(seeds), n, n, ptr(bits), ptr(sigma_cw), ptr(s_cw), ptr(leaf), stream())
    return bits, sigma_cw, s_cw, leaf


def dif_eval(b: int, x_masked, win: FSSWindow):
    """DIF.eval (fss.py:401-428)"""
    x = x_masked.contiguous()
    assert x.numel() == win.n, (x.shape, win.n)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        call("pm_fss_dif_eval", int(b), ptr(x), _p(win.s0), _p(win.bits), _p(win.sigma_cw), _p(win.s_cw), _p(win.leaf),
             win.n, win.stride, ptr(out), stream())
    return out


def dif_eval_open(b: int, r_own, r_peer, win: FSSWindow):
    """DIF.eval on (r_own + r_peer) mod 2^32: the opening of the masked difference (fss.py:158) happens inside the kernel"""
    x = r_own.contiguous()
    assert x.numel() == win.n == r_peer.numel(), (x.shape, win.n)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        call("pm_fss_dif_eval_open", int(b), ptr(x), ptr(r_peer), _p(win.s0), _p(win.bits), _p(win.sigma_cw), _p(win.s_cw),
             _p(win.leaf), win.n, win.stride, ptr(out), stream())
    return out


def mask_builder(x1, x2, alpha_share):
    """fss.py:189-204: x1 - x2 + alpha (either operand may be None = public 0)"""
    ref = x1 if x1 is not None else x2
    r = torch.empty_like(ref, memory_format=torch.contiguous_format)
    assert alpha_share.is_contiguous() and alpha_share.numel() == ref.numel()
    with torch.cuda.device(ref.device):
        call("pm_fss_mask_i64", ptr(x1.contiguous()) if x1 is not None else None,
             ptr(x2.contiguous()) if x2 is not None else None, _p(alpha_share), ptr(r), ref.numel(), stream())
    return r


def open_mod32(local, peer):
    out = torch.empty_like(local)
    with torch.cuda.device(local.device):
        call("pm_fss_open_mod32_i64", ptr(local), ptr(peer), ptr(out), local.numel(), stream())
    return out


def pre_pool(x, k, stride, pad):
    """_pre_pool nn/functional.py:312-390 -> [B,C,M,k*k]"""
    x = x.contiguous()
    B, C, H, W = x.shape
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    out = torch.empty((B, C, Ho * Wo, k * k), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_pre_pool_i64", ptr(x), B, C, H, W, k, stride, pad, ptr(out), stream())
    return out


def slice_lastdim(t, start, length):
    """t[..., start:start+length] as a new contiguous tensor"""
    t = t.contiguous()
    L = t.shape[-1]
    rows = t.numel() // L
    out = torch.empty((*t.shape[:-1], length), dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        call("pm_slice_lastdim_i64", ptr(t), rows, L, start, length, ptr(out), stream())
    return out


# ------------------------------------------------------------------------------------------ provider side
def build_fss_keys(n_instances: int, device, seed: int, counter: int, epoch=None):
    """build_separate_fss_keys (primitives.py:237-253): DIF.keygen on the crypto provider's GPU from Philox randomness,
    alpha split additively mod 2^32.  Returns [keys of party 0, keys of party 1] (sharing the correction words)."""
    n = int(n_instances)
    alpha = ops.random_i64((n,), seed, counter, device, epoch)
    mask = ops.random_i64((n,), seed, counter + 1, device, epoch)
    seeds = ops.random_i64((2, 2, n), seed, counter + 2, device, epoch)
    alpha0 = torch.empty((n,), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        call("pm_fss_condition_randomness", ptr(alpha), ptr(mask), ptr(seeds), ptr(alpha0), n, stream())
    bits, sigma_cw, s_cw, leaf = dif_keygen(alpha, seeds)
    return [FSSKeys(alpha0, seeds[0], bits, sigma_cw, s_cw, leaf), FSSKeys(mask, seeds[1], bits, sigma_cw, s_cw, leaf)]


# ------------------------------------------------------------------------------------------ protocol
def le(x1_shares, x2_shares, parties, provider=None):
    """fss.le -> fss_op(x1, x2, "comp") (fss.py:97-185,279-283): per-party int64 shares of [x1 <= x2]."""
    from .spdz import EmptyCryptoPrimitiveStoreError, _ensure_peer, release_after_peer_reads

    ref = x1_shares if x1_shares[0] is not None else x2_shares
    n = ref[0].numel()
    try:
        wins = [p.crypto_store.get_keys(op=OP, n_instances=n, remove=False) for p in parties]
    except EmptyCryptoPrimitiveStoreError as e:
        if provider is None or any(p.crypto_store.force_preprocessing for p in parties):
            raise
        provider.provide_primitives(parties=parties, **e.kwargs_)
        return le(x1_shares, x2_shares, parties, provider)
    r = [mask_builder(x1_shares[j], x2_shares[j], wins[j].alpha) for j in range(2)]
    from . import spdz as _spdz

    if _spdz.FUSE_OPEN:
        peers = _spdz.peer_views(parties, r)
        out = []
        for j, p in enumerate(parties):
            win = p.crypto_store.get_keys(op=OP, n_instances=n, remove=True)
            out.append(dif_eval_open(j, r[j], peers[j].contiguous(), win).view(ref[j].shape))
        release_after_peer_reads(parties)
        return out
    masked = []
    for j, p in enumerate(parties):
        peer = r[1 - j]
        if peer.device != p.device:
            torch.cuda.current_stream(p.device).wait_stream(torch.cuda.current_stream(peer.device))
            if not _ensure_peer(p.device, peer.device):
                peer = peer.to(p.device)
        masked.append(open_mod32(r[j], peer))
    release_after_peer_reads(parties)
    out = []
    for j, p in enumerate(parties):
        win = p.crypto_store.get_keys(op=OP, n_instances=n, remove=True)
        out.append(dif_eval(j, masked[j], win).view(ref[j].shape))
    return out
