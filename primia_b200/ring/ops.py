"""Tensor-level wrappers of the ring (Z_2^64) entry points of the C ABI.

Each function takes/returns contiguous CUDA int64 tensors and launches on the current
stream of the tensor's device.  No CPU path exists."""
from __future__ import annotations

import ctypes
import os

import torch

from .._lib import PrimiaError, call, ptr, stream

I64 = torch.int64

__all__ = [
    "encode", "decode", "share_gen", "random_i64", "im2col", "mask", "mask_im2col", "mask_wt", "open_add",
    "combine_matmul", "combine_mul", "matmul", "trunc_div", "trunc_post_conv", "axpby", "avgpool",
    "nchw_to_pc", "pc_to_nchw", "conv_out_size", "stack", "bn_newton_fused", "bn_newton_p2p", "new_epoch", "bump_epoch",
]


def _chk(t, dtype=I64):
    if not t.is_cuda:
        raise PrimiaError("primia_b200 ring ops take CUDA tensors only (there is no CPU fallback)")
    if t.dtype != dtype:
        raise PrimiaError(f"expected {dtype}, got {t.dtype}")
    return t.contiguous()


def conv_out_size(H, k, stride, pad, dil=1):
    return (H + 2 * pad - dil * (k - 1) - 1) // stride + 1


def encode(x: torch.Tensor, base: int = 10, precision_fractional: int = 16, check_range: bool = True):
    """precision.py:117-132"""
    x = _chk(x, torch.float32)
    q = torch.empty(x.shape, dtype=I64, device=x.device)
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_encode_f32_i64", ptr(x), ctypes.c_float(float(base ** precision_fractional)), ptr(q), x.numel(),
             ptr(flag), stream())
    if check_range and int(flag.item()):
        raise AssertionError("tensor cannot be correctly embedded: choose bigger field or a lower precision")
    return q


def decode(q: torch.Tensor, base: int = 10, precision_fractional: int = 16):
    """precision.py:134-144"""
    q = _chk(q)
    x = torch.empty(q.shape, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        call("pm_decode_i64_f32", ptr(q), ctypes.c_float(float(base ** precision_fractional)), ptr(x), q.numel(), stream())
    return x


def new_epoch(device):
    """a device-side Philox epoch (uint64, starts at 0) owned by ONE generator: its value is added to the high half of every
    offset that generator uses, and ``bump_epoch`` advances it with a kernel -- so a generation captured in a CUDA graph draws a
    fresh stream at every replay while staying a pure function of (seed, number of bumps)"""
    return torch.zeros(1, dtype=I64, device=torch.device(device))


def bump_epoch(epoch):
    with torch.cuda.device(epoch.device):
        call("pm_epoch_bump", ptr(epoch), stream())


def share_gen(q: torch.Tensor, seed: int, offset: int, out=None, epoch=None):
    """additive_shared.py:336-365 (2 parties); ``out`` = (s0, s1) writes into existing tensors"""
    q = _chk(q)
    s0, s1 = out if out is not None else (torch.empty_like(q), torch.empty_like(q))
    with torch.cuda.device(q.device):
        call("pm_share_gen_i64", ptr(q), seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), ptr(epoch), ptr(s0), ptr(s1), q.numel(), stream())
    return s0, s1


def random_i64(shape, seed: int, offset: int, device, epoch=None):
    out = torch.empty(shape, dtype=I64, device=device)
    with torch.cuda.device(out.device):
        call("pm_random_i64", seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), ptr(epoch), ptr(out), out.numel(), stream())
    return out


def im2col(x, kh, kw, stride, pad, dil=1):
    x = _chk(x)
    B, C, H, W = x.shape
    Ho, Wo = conv_out_size(H, kh, stride, pad, dil), conv_out_size(W, kw, stride, pad, dil)
    out = torch.empty((B, Ho * Wo, C * kh * kw), dtype=I64, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_im2col_i64", ptr(x), B, C, H, W, kh, kw, stride, pad, dil, ptr(out), stream())
    return out


def mask(x, a):
    x, a = _chk(x), _chk(a)
    assert x.shape == a.shape, (x.shape, a.shape)
    d = torch.empty_like(x)
    with torch.cuda.device(x.device):
        call("pm_spdz_mask_i64", ptr(x), ptr(a), ptr(d), x.numel(), stream())
    return d


def mask2(x, a, y, b):
    """spdz_mask of both operands of a same-shape product in one pass: (x - a, y - b)"""
    x, a, y, b = map(_chk, (x, a, y, b))
    assert x.shape == a.shape == y.shape == b.shape, (x.shape, a.shape, y.shape, b.shape)
    d, e = torch.empty_like(x), torch.empty_like(y)
    with torch.cuda.device(x.device):
        call("pm_spdz_mask2_i64", ptr(x), ptr(a), ptr(y), ptr(b), ptr(d), ptr(e), x.numel(), stream())
    return d, e


def combine_mul_open(j, d_own, d_peer, e_own, e_peer, a, b, c):
    """spdz_compute (same-shape "mul") with both openings fused: delta = d_own + d_peer, eps = e_own + e_peer in registers"""
    d_own, d_peer, e_own, e_peer, a, b, c = map(_chk, (d_own, d_peer, e_own, e_peer, a, b, c))
    z = torch.empty_like(c)
    with torch.cuda.device(d_own.device):
        call("pm_spdz_combine_mul_open_i64", j, ptr(d_own), ptr(d_peer), ptr(e_own), ptr(e_peer), ptr(a), ptr(b), ptr(c),
             c.numel(), ptr(z), stream())
    return z


def mask_im2col(x, a, kh, kw, stride, pad, dil=1):
    x, a = _chk(x), _chk(a)
    B, C, H, W = x.shape
    Ho, Wo = conv_out_size(H, kh, stride, pad, dil), conv_out_size(W, kw, stride, pad, dil)
    assert tuple(a.shape) == (B, Ho * Wo, C * kh * kw), (a.shape, (B, Ho * Wo, C * kh * kw))
    d = torch.empty_like(a)
    with torch.cuda.device(x.device):
        call("pm_spdz_mask_im2col_i64", ptr(x), B, C, H, W, kh, kw, stride, pad, dil, ptr(a), ptr(d), stream())
    return d


def mask_wt(w2d, b):
    """eps = w2d.t() - b with w2d [N,K] contiguous, b [K,N]"""
    w2d, b = _chk(w2d), _chk(b)
    N, K = w2d.shape
    assert tuple(b.shape) == (K, N)
    e = torch.empty_like(b)
    with torch.cuda.device(b.device):
        call("pm_spdz_mask_wt_i64", ptr(w2d), N, K, ptr(b), ptr(e), stream())
    return e


def open_add(local, peer):
    """out = local + peer ; ``peer`` may live on another GPU (peer-mapped pointer, NVLink load)."""
    local, peer = _chk(local), _chk(peer)
    out = torch.empty_like(local)
    with torch.cuda.device(local.device):
        call("pm_open_add_i64", ptr(local), ptr(peer), ptr(out), local.numel(), stream())
    return out


USE_TENSOR_CORES = os.environ.get("PRIMIA_RING_TC", "1") != "0"


def gemm2_tc(A1, B1, A2, B2, Cinit):
    """C = Cinit + A1@B1 + A2@B2 over Z_2^64 on the int8 tensor cores (limb decomposition, exact). A: [rows,K], B: [K,N]."""
    from .._lib import lib

    rows, K = A1.shape
    N = B1.shape[1]
    l = lib()
    l.pm_ring_tc_ws_bytes.restype = ctypes.c_size_t
    nseg = 2 if A2 is not None else 1
    ws = torch.empty(int(l.pm_ring_tc_ws_bytes(rows, K, N, nseg)), dtype=torch.uint8, device=A1.device)
    C = torch.empty((rows, N), dtype=I64, device=A1.device)
    with torch.cuda.device(A1.device):
        call("pm_ring_gemm2_tc_i64", ptr(A1), ptr(B1), ptr(A2) if A2 is not None else None, ptr(B2) if B2 is not None else None,
             ptr(Cinit) if Cinit is not None else None, rows, K, N, ptr(ws), ptr(C), stream())
    return C


def _planes(n, K, device):
    from .._lib import lib

    l = lib()
    l.pm_ring_planes_bytes.restype = ctypes.c_size_t
    return torch.empty(int(l.pm_ring_planes_bytes(int(n), int(K))), dtype=torch.uint8, device=device)


def planarize_rows(A, peer=None):
    """limb planes of the left operand A [rows,K] -- or of (A + peer) mod 2^64 when ``peer`` is given: the OPENING of a masked
    share fused into the planarisation (``peer`` may be a peer-mapped tensor on another GPU)."""
    A = _chk(A)
    K = A.shape[-1]
    rows = A.numel() // K
    out = _planes(rows, K, A.device)
    with torch.cuda.device(A.device):
        call("pm_ring_planarize_rows_i64", ptr(A), ptr(_chk(peer)) if peer is not None else None, rows, K, ptr(out), stream())
    return out


def planarize_cols(Bm):
    """limb planes of the right operand B [K,N] (stored transposed, K contiguous)"""
    Bm = _chk(Bm)
    K, N = Bm.shape
    out = _planes(N, K, Bm.device)
    with torch.cuda.device(Bm.device):
        call("pm_ring_planarize_cols_i64", ptr(Bm), K, N, ptr(out), stream())
    return out


def gemm_planes(pa1, pb1, pa2, pb2, Cinit, rows, K, N):
    """C = Cinit + A1@B1 (+ A2@B2) from limb planes built by planarize_rows / planarize_cols"""
    C = torch.empty((rows, N), dtype=I64, device=pa1.device)
    with torch.cuda.device(pa1.device):
        call("pm_ring_gemm_planes_i64", ptr(pa1), ptr(pb1), ptr(pa2) if pa2 is not None else None,
             ptr(pb2) if pb2 is not None else None, ptr(Cinit) if Cinit is not None else None, rows, K, N, ptr(C), stream())
    return C


def tc_supported(rows, K, N):
    from .._lib import lib

    return USE_TENSOR_CORES and bool(lib().pm_ring_tc_supported(int(rows), int(K), int(N)))


def combine_matmul(j, delta, eps, a, b, c):
    delta, eps, a, b, c = map(_chk, (delta, eps, a, b, c))
    if tc_supported(delta.numel() // delta.shape[-1], delta.shape[-1], eps.shape[1]):
        # delta@b + delta@eps == delta@(b + eps) exactly in the ring (party 0); then one 2-segment limb GEMM
        b_eff = axpby(1, b, 1, eps) if j == 0 else b
        K = delta.shape[-1]
        z = gemm2_tc(delta.reshape(-1, K), b_eff, a.reshape(-1, K), eps, c.reshape(-1, c.shape[-1]))
        return z.view(c.shape)
    if delta.dim() == 2:  # plain [M,K] @ [K,N] (linear, functional.py:10-14)
        Bt, (M, K) = 1, delta.shape
        assert tuple(c.shape) == (M, eps.shape[1])
    else:
        Bt, M, K = delta.shape
        assert tuple(c.shape) == (Bt, M, eps.shape[1])
    K2, N = eps.shape
    assert K == K2 and a.shape == delta.shape and b.shape == eps.shape
    z = torch.empty_like(c)
    ws = torch.empty_like(b) if j == 0 else None
    with torch.cuda.device(delta.device):
        call("pm_spdz_combine_matmul_i64", j, ptr(delta), ptr(eps), ptr(a), ptr(b), ptr(c), Bt, M, K, N, ptr(ws), ptr(z),
             stream())
    return z


def combine_mul(j, delta, eps, a, b, c):
    delta, eps, a, b, c = map(_chk, (delta, eps, a, b, c))
    if delta.shape == eps.shape:
        mode, P, C = 0, 1, delta.numel()
    elif delta.dim() == 1 and eps.dim() == 2 and eps.shape[1] == delta.shape[0]:
        mode, (P, C) = 1, eps.shape
    elif eps.dim() == 1 and delta.dim() == 2 and delta.shape[1] == eps.shape[0]:
        mode, (P, C) = 2, delta.shape
    else:
        raise PrimiaError(f"unsupported broadcast {tuple(delta.shape)} x {tuple(eps.shape)}")
    z = torch.empty_like(c)
    with torch.cuda.device(delta.device):
        call("pm_spdz_combine_mul_i64", j, ptr(delta), ptr(eps), ptr(a), ptr(b), ptr(c), mode, P, C, ptr(z), stream())
    return z


def matmul(A, Bm):
    A, Bm = _chk(A), _chk(Bm)
    if tc_supported(A.numel() // A.shape[-1], A.shape[-1], Bm.shape[1]):
        return gemm2_tc(A.reshape(-1, A.shape[-1]), Bm, None, None, None).view(*A.shape[:-1], Bm.shape[1])
    squeeze = A.dim() == 2
    if squeeze:
        A = A.unsqueeze(0)
    Bt, M, K = A.shape
    K2, N = Bm.shape
    assert K == K2
    C = torch.empty((Bt, M, N), dtype=I64, device=A.device)
    with torch.cuda.device(A.device):
        call("pm_matmul_i64", ptr(A), ptr(Bm), Bt, M, K, N, ptr(C), stream())
    return C[0] if squeeze else C


def trunc_div(x, divisor: int, out=None):
    x = _chk(x)
    out = torch.empty_like(x) if out is None else out
    with torch.cuda.device(x.device):
        call("pm_trunc_div_i64", ptr(x), int(divisor), ptr(out), x.numel(), stream())
    return out


def trunc_post_conv(z, divisor: int, bias, Ho, Wo):
    z = _chk(z)
    Bt, M, N = z.shape
    out = torch.empty((Bt, N, Ho, Wo), dtype=I64, device=z.device)
    with torch.cuda.device(z.device):
        call("pm_trunc_post_conv_i64", ptr(z), int(divisor), ptr(bias) if bias is not None else None, Bt, M, N, ptr(out),
             stream())
    return out


def axpby(alpha: int, x, beta: int = 0, y=None):
    x = _chk(x)
    if y is None:
        ybcast, P, C = 0, 1, x.numel()
    else:
        y = _chk(y)
        if y.shape == x.shape:
            ybcast, P, C = 0, 1, x.numel()
        elif y.numel() == 1:
            ybcast, P, C = 2, 1, x.numel()
        elif y.dim() == 1 and x.shape[-1] == y.shape[0]:
            ybcast, C = 1, y.shape[0]
            P = x.numel() // C
        else:
            raise PrimiaError(f"unsupported broadcast {tuple(x.shape)} , {tuple(y.shape)}")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        call("pm_axpby_i64", int(alpha), ptr(x), int(beta), ptr(y), ybcast, P, C, ptr(out), stream())
    return out


def spdz_affine(u, peer, sc, add, div: int, chan, cs: int, elem, es: int, C: int, HW: int):
    """out[i] = T(sc[c] * (u[i] + peer[i]) + add[i]) + cs * chan[c] + es * elem[i] on NCHW shares, c = (i // HW) % C; T = C-style
    division by ``div`` when div > 1.  None operands drop out.  One pass of a hoisted Beaver product (functional.batch_norm)."""
    u = _chk(u)
    out = torch.empty_like(u)
    o = lambda t: ptr(_chk(t)) if t is not None else None
    with torch.cuda.device(u.device):
        call("pm_spdz_affine_i64", ptr(u), o(peer), o(sc), o(add), int(div), o(chan), int(cs), o(elem), int(es), int(C), int(HW),
             u.numel(), ptr(out), stream())
    return out


def avgpool(x, k: int):
    x = _chk(x)
    B, C, H, W = x.shape
    out = torch.empty((B, C, H // k, W // k), dtype=I64, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_avgpool_i64", ptr(x), B, C, H, W, k, ptr(out), stream())
    return out


def nchw_to_pc(x):
    x = _chk(x)
    B, C, H, W = x.shape
    out = torch.empty((B * H * W, C), dtype=I64, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_nchw_to_pc_i64", ptr(x), B, C, H * W, ptr(out), stream())
    return out


def pc_to_nchw(x, B, C, H, W):
    x = _chk(x)
    out = torch.empty((B, C, H, W), dtype=I64, device=x.device)
    with torch.cuda.device(x.device):
        call("pm_pc_to_nchw_i64", ptr(x), B, C, H * W, ptr(out), stream())
    return out


def stack(tensors):
    """[n] same-shape tensors -> one contiguous [n, ...] buffer (a no-copy view when they already are consecutive slices of
    one allocation, as bulk-generated triples are)."""
    t0 = tensors[0]
    step = t0.numel() * t0.element_size()
    if all(t.is_contiguous() and t.data_ptr() == t0.data_ptr() + i * step and t.untyped_storage().data_ptr() ==
           t0.untyped_storage().data_ptr() for i, t in enumerate(tensors)):
        return torch.as_strided(t0, (len(tensors), *t0.shape), (t0.numel(), *t0.stride()))
    return torch.stack([_chk(t) for t in tensors])


def bn_newton_fused(jobs, iters: int, divisor: int, newton_c: int):
    """precision.py:507-518 for both co-resident parties, several vectors (BatchNorm layers) per launch.
    jobs: list of (v, tri, k) with v = [v0, v1] ([C]); tri[j] = (a, b, c) each [3*(iters-1), C]; k = (k0, k1) each [iters].
    Returns [[x0, x1], ...]."""
    from .._lib import NewtonJob, lib

    outs = []
    dev = jobs[0][0][0].device
    MAXJ = 32  # PM_NEWTON_MAX_JOBS
    for base in range(0, len(jobs), MAXJ):
        chunk = jobs[base:base + MAXJ]
        arr = (NewtonJob * len(chunk))()
        keep = []
        for i, (v, tri, k) in enumerate(chunk):
            v0, v1 = _chk(v[0]), _chk(v[1])
            C = v0.shape[0]
            x0, x1 = torch.empty_like(v0), torch.empty_like(v1)
            (a0, b0, c0), (a1, b1, c1) = [[_chk(t) for t in tri[j]] for j in range(2)]
            assert tuple(a0.shape) == (3 * (iters - 1), C), (a0.shape, iters, C)
            k0, k1 = _chk(k[0]), _chk(k[1])
            for name, t in zip(("v0", "v1", "a0", "b0", "c0", "a1", "b1", "c1", "k0", "k1", "x0", "x1"),
                               (v0, v1, a0, b0, c0, a1, b1, c1, k0, k1, x0, x1)):
                setattr(arr[i], name, t.data_ptr())
            arr[i].C = C
            keep.append((v0, v1, a0, b0, c0, a1, b1, c1, k0, k1))
            outs.append([x0, x1])
        with torch.cuda.device(dev):
            call("pm_bn_newton_fused_i64", ctypes.cast(arr, ctypes.c_void_p), len(chunk), iters, int(divisor), int(newton_c),
                 stream())
    return outs


_P2P_MAILBOX = {}


def _p2p_mailbox(dev0, dev1, n_jobs, iters, max_c):
    """per GPU pair: (inbox on dev0, inbox on dev1, epoch0, epoch1, err0, err1), allocated and zeroed once"""
    from .._lib import lib
    from .spdz import _ensure_peer

    key = (dev0, dev1, n_jobs, iters, max_c)
    st = _P2P_MAILBOX.get(key)
    if st is None:
        if dev0 != dev1 and not (_ensure_peer(dev0, dev1) and _ensure_peer(dev1, dev0)):
            raise PrimiaError(f"no peer access between {dev0} and {dev1}: the cross-GPU Newton kernel needs NVLink / PCIe P2P")
        fn = lib().pm_bn_newton_p2p_mailbox_bytes
        fn.restype = ctypes.c_size_t
        nbytes = int(fn(n_jobs, iters, max_c))
        st = tuple(torch.zeros(nbytes // 8, dtype=I64, device=d) for d in (dev0, dev1)) + tuple(
            torch.zeros(1, dtype=I64, device=d) for d in (dev0, dev1)) + tuple(
            torch.zeros(1, dtype=torch.int32, device=d) for d in (dev0, dev1))
        for d in (dev0, dev1):
            torch.cuda.synchronize(d)
        _P2P_MAILBOX[key] = st
    return st


def bn_newton_p2p(jobs, iters: int, divisor: int, newton_c: int):
    """precision.py:507-518 with the two share holders on DIFFERENT GPUs: one kernel per party, launched back to back on the two
    devices' current streams, exchanging the masked operands of every Beaver product through peer-mapped mailboxes over NVLink
    (pm_bn_newton_p2p_i64).  Same job layout as ``bn_newton_fused``; party j's tensors live on party j's GPU."""
    from .._lib import NewtonP2PJob

    assert len(jobs) <= 32
    devs = [jobs[0][0][j].device for j in range(2)]
    max_c = max(int(v[0].shape[0]) for v, _t, _k in jobs)
    box0, box1, ep0, ep1, err0, err1 = _p2p_mailbox(devs[0], devs[1], len(jobs), iters, max_c)
    outs = [[None, None] for _ in jobs]
    keep = []
    # the two kernels talk to each other while running: neither launch may be stream-ordered after the other.  On two GPUs
    # the devices' current streams are independent; with both parties on ONE GPU (tests) party 1 gets a side stream that is
    # forked BEFORE party 0's kernel is enqueued.
    side = cur = None
    if devs[0] == devs[1]:
        with torch.cuda.device(devs[0]):
            cur = torch.cuda.current_stream()
            side = _P2P_MAILBOX.setdefault(("side", devs[0]), torch.cuda.Stream(devs[0]))
            side.wait_stream(cur)
    for j in range(2):
        arr = (NewtonP2PJob * len(jobs))()
        for i, (v, tri, k) in enumerate(jobs):
            vj, kj = _chk(v[j]), _chk(k[j])
            a, b, c = (_chk(t) for t in tri[j])
            C = vj.shape[0]
            assert tuple(a.shape) == (3 * (iters - 1), C) and vj.device == devs[j] and a.device == devs[j] and kj.device == devs[j]
            x = torch.empty_like(vj)
            for name, t in zip(("v", "a", "b", "c", "k", "x"), (vj, a, b, c, kj, x)):
                setattr(arr[i], name, t.data_ptr())
            arr[i].C = C
            keep.append((vj, kj, a, b, c))
            outs[i][j] = x
        inbox, peer = (box0, box1) if j == 0 else (box1, box0)
        with torch.cuda.device(devs[j]):
            st = side if (j == 1 and side is not None) else torch.cuda.current_stream()
            call("pm_bn_newton_p2p_i64", j, ctypes.cast(arr, ctypes.c_void_p), len(jobs), iters, int(divisor), int(newton_c),
                 ptr(inbox), ctypes.c_void_p(peer.data_ptr()), ptr(ep0 if j == 0 else ep1), max_c, ptr(err0 if j == 0 else err1),
                 ctypes.c_void_p(st.cuda_stream))
    if side is not None:
        cur.wait_stream(side)
        for o in outs:
            o[1].record_stream(cur)
    bn_newton_p2p.last_err = (err0, err1)
    return outs
