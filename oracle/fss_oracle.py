"""CPU oracle for the function-secret-sharing comparison used by ReLU / max-pool on shares -- TEST INFRASTRUCTURE ONLY.

Restates ``syft/frameworks/torch/mpc/fss.py`` (DIF = distributed interval function, "x <= alpha" on n = 32 bits with
lambda = 127) in plain numpy + hashlib.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import
this module.  Citations are relative to ``/root/reference``.

Third-party dependency absent from the reference tree: ``shaloop`` (unpinned, fss.py:14) -- its
``sha512_loop_func(x[n,16] uint8, out[n,64] uint8)`` is taken to be SHA-512 of each 16-byte row (FIPS 180-4), restated
with ``hashlib``.  The pin: ``tests/golden/make_golden.py`` executes the reference's own ``DIF.keygen`` / ``DIF.eval`` /
``H`` / ``compress`` / ``uncompress`` with that hashlib stand-in for shaloop and stores keys + outputs
(``tests/golden/fss_dif.npz``); ``tests/test_oracle_fss.py`` checks this restatement against them.  Randomness
(``np.random`` for alpha, the root seeds and the alpha mask, fss.py:346,354; primitives.py:245-251) is an explicit input.

Key material of one comparison (per value), as the reference lays it out:
    alpha   uint32 (held additively shared mod 2^32 by the two parties, primitives.py:245-251)
    s0[b]   2 x uint64 root seed of party b (first word < 2^63)
    per level i < 32 (compressed correction word, fss.py:431-455):  tauL, tL, tauR, tR bits ; sigma_cw[2] ; s_cw[2]
    leaf[33] int32
"""
from __future__ import annotations

import hashlib

import numpy as np

N_BITS = 32
MASK31 = np.uint64(0x7FFFFFFF)
CLR1 = np.uint64(0xFFFFFFFFFFFFFFFE)
ONE = np.uint64(1)


def H(seed: np.ndarray) -> np.ndarray:
    """PRG lambda -> 4(lambda+1)  (fss.py:553-601).  seed [2, n] uint64 -> valuebits [2, 6, n] uint64:
    row r = (sigma[2], tau, s[2], t) for direction r (0 = left, 1 = right)."""
    n = seed.shape[1]
    msg = np.ascontiguousarray(seed.T).view(np.uint8).reshape(n, 16)
    dig = np.empty((n, 64), dtype=np.uint8)
    for i in range(n):
        dig[i] = np.frombuffer(hashlib.sha512(msg[i].tobytes()).digest(), dtype=np.uint8)
    buf = dig.view(np.uint64).T  # [8, n] little-endian words of the digest bytes
    out = np.empty((2, 6, n), dtype=np.uint64)
    for r in range(2):
        w = buf[4 * r: 4 * r + 4]
        out[r, 0], out[r, 1], out[r, 2] = w[0] & CLR1, w[1], w[0] & ONE
        out[r, 3], out[r, 4], out[r, 5] = w[2] & CLR1, w[3], w[2] & ONE
    return out


def bits_msb_first(x: np.ndarray) -> np.ndarray:
    """bit_decomposition (fss.py:487-495): [32, n], row 0 = most significant bit of the low 32 bits."""
    x = x.astype(np.uint64) & np.uint64(0xFFFFFFFF)
    return np.stack([(x >> np.uint64(N_BITS - 1 - i)) & ONE for i in range(N_BITS)])


def conv31(words: np.ndarray) -> np.ndarray:
    """convert (fss.py:655-661): the 31 low bits of the last word, as int64."""
    return (words[-1] & MASK31).astype(np.int64)


def _sel(a, b, bit):
    """(1-bit)*a + bit*b for uint64 arrays (multi_dim_filter fss.py:648-650)"""
    return np.where(bit.astype(bool), b, a)


def dif_keygen(alpha: np.ndarray, seeds: np.ndarray):
    """DIF.keygen (fss.py:341-399) with explicit randomness.

    alpha [n] uint32-valued ; seeds [2, 2, n] uint64 (party, word, value; word 0 < 2^63 as randbit produces, fss.py:498-505).
    Returns dict(alpha, s0 [2,2,n], bits [32,4,n] uint8 (tauL,tL,tauR,tR), sigma_cw [32,2,n], s_cw [32,2,n], leaf [33,n] int32)."""
    n = alpha.shape[0]
    a_bits = bits_msb_first(alpha)
    s = [seeds[0].copy(), seeds[1].copy()]
    t = [np.zeros(n, np.uint64), np.ones(n, np.uint64)]
    bits = np.empty((N_BITS, 4, n), np.uint8)
    sigma_cw = np.empty((N_BITS, 2, n), np.uint64)
    s_cw = np.empty((N_BITS, 2, n), np.uint64)
    leaf = np.empty((N_BITS + 1, n), np.int64)
    for i in range(N_BITS):
        ai = a_bits[i]
        h = [H(s[0]), H(s[1])]
        x = h[0] ^ h[1]                      # [2,6,n]
        # SwitchTableDIF (fss.py:628-645): leaf part switched by 1-alpha_i, next part by alpha_i
        s_rand = _sel(x[1, 3:5], x[0, 3:5], ai)       # (sL0^sL1)*a + (sR0^sR1)*(1-a)
        sg_rand = _sel(x[1, 0:2], x[0, 0:2], ai)
        table = np.zeros((2, 6, n), np.uint64)
        # leafTable: row0 = a*(sg_rand,1) ; row1 = (1-a)*(sg_rand,1)
        table[0, 0:2] = sg_rand * ai
        table[0, 2] = ai
        table[1, 0:2] = sg_rand * (ONE - ai)
        table[1, 2] = ONE - ai
        # nextTable: row0 = (1-a)*(s_rand,1) ; row1 = a*(s_rand,1)
        table[0, 3:5] = s_rand * (ONE - ai)
        table[0, 5] = ONE - ai
        table[1, 3:5] = s_rand * ai
        table[1, 5] = ai
        cw = table ^ x                        # CW[i] = cw_i ^ h0 ^ h1
        # compress (fss.py:431-455) then uncompress (:458-479)
        bits[i, 0], bits[i, 1], bits[i, 2], bits[i, 3] = cw[0, 2], cw[0, 5], cw[1, 2], cw[1, 5]
        sigma_cw[i] = _sel(cw[0, 0:2], cw[1, 0:2], ai)      # a*sigmaR + (1-a)*sigmaL
        s_cw[i] = _sel(cw[1, 3:5], cw[0, 3:5], ai)          # (1-a)*sR + a*sL
        cwi = np.empty((2, 6, n), np.uint64)
        for r in range(2):
            cwi[r, 0:2], cwi[r, 2] = sigma_cw[i], bits[i, 2 * r].astype(np.uint64)
            cwi[r, 3:5], cwi[r, 5] = s_cw[i], bits[i, 2 * r + 1].astype(np.uint64)
        sig, tau = [None, None], [None, None]
        for b in range(2):
            dual = h[b] ^ (t[b] * cwi)
            keep = _sel(dual[0], dual[1], ai)               # follow the special path
            anti = _sel(dual[1], dual[0], ai)               # leave it
            s[b], t[b] = keep[3:5], keep[5]
            sig[b], tau[b] = anti[0:2], anti[2]
        sign = np.where(tau[1].astype(bool), -1, 1).astype(np.int64)
        leaf[i] = sign * (1 - conv31(sig[0]) + conv31(sig[1]) - (1 - ai.astype(np.int64)))
    sign = np.where(t[1].astype(bool), -1, 1).astype(np.int64)
    leaf[N_BITS] = sign * (1 - conv31(s[0]) + conv31(s[1]))
    return {"alpha": alpha.astype(np.uint64), "s0": seeds.copy(), "bits": bits, "sigma_cw": sigma_cw, "s_cw": s_cw,
            "leaf": leaf.astype(np.int32)}


def dif_eval(b: int, x_masked: np.ndarray, key) -> np.ndarray:
    """DIF.eval (fss.py:401-428): party b's int64 share of [x_masked <= alpha] over the low 32 bits."""
    x = np.asarray(x_masked).reshape(-1)
    n = x.shape[0]
    xb = bits_msb_first(x)
    s = key["s0"][b].copy()
    t = np.full(n, b, np.uint64)
    leaf = key["leaf"].astype(np.int64)
    sign = -1 if b else 1
    acc = np.zeros(n, np.int64)
    for i in range(N_BITS):
        h = H(s)
        cwi = np.empty((2, 6, n), np.uint64)
        for r in range(2):
            cwi[r, 0:2], cwi[r, 2] = key["sigma_cw"][i], key["bits"][i, 2 * r].astype(np.uint64)
            cwi[r, 3:5], cwi[r, 5] = key["s_cw"][i], key["bits"][i, 2 * r + 1].astype(np.uint64)
        dual = h ^ (t * cwi)
        st = _sel(dual[0], dual[1], xb[i])
        sig, tau, s, t = st[0:2], st[2], st[3:5], st[5]
        with np.errstate(over="ignore"):
            acc = acc + sign * (tau.astype(np.int64) * leaf[i] + conv31(sig))
    with np.errstate(over="ignore"):
        acc = acc + sign * (t.astype(np.int64) * leaf[N_BITS] + conv31(s))
    return acc.reshape(np.asarray(x_masked).shape)


def split_alpha(alpha: np.ndarray, mask: np.ndarray):
    """build_separate_fss_keys (primitives.py:245-251): party 0 holds (alpha - mask) mod 2^32, party 1 holds mask."""
    return [(alpha.astype(np.int64) - mask.astype(np.int64)) % (1 << N_BITS), mask.astype(np.int64)]


def fss_le(x1_sh, x2_sh, key, alpha_sh):
    """fss_op(x1, x2, "comp") (fss.py:97-185): mask_builder on each party (:189-204), opening mod 2^32 (:158),
    evaluate (:208-245).  x*_sh: per-party int64 numpy arrays.  Returns per-party int64 shares of [x1 <= x2]."""
    r = [x1_sh[j] - x2_sh[j] + alpha_sh[j].reshape(x1_sh[j].shape) for j in range(2)]
    with np.errstate(over="ignore"):
        masked = (r[0] + r[1]) % (1 << N_BITS)
    return [dif_eval(j, masked, key) for j in range(2)]
