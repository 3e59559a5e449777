"""SPDZ Beaver multiplication on GPU-resident shares.

Mirrors syft/frameworks/torch/mpc/{spdz,beaver,primitives}.py:
  spdz_mask     spdz.py:22-45      party-local  delta_j = x_j - a_j , eps_j = y_j - b_j
  spdz_compute  spdz.py:64-122     party-local  z_j = delta(op)b_j + a_j(op)eps + c_j (+ delta(op)eps, j==0)
  spdz_mul      spdz.py:125-197    the protocol round (mask -> open -> compute)
  PrimitiveStorage primitives.py:52-102,161-213,255-286 ; build_triple beaver.py:7-63
  EmptyCryptoPrimitiveStoreError syft/exceptions.py:343-356

B200 mapping: a Party is a GPU (``torch.device``) + its crypto store.  Opening delta/eps is a
symmetric peer read over NVLink (``pm_open_add_i64`` reads the other party's buffer through its
peer-mapped pointer) instead of the reference's star through the orchestrator (spdz.py:162-163).
Triples are produced on the crypto provider's GPU (Philox) and copied to the two parties.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Tuple

import os

import torch

from . import ops


class EmptyCryptoPrimitiveStoreError(Exception):
    """syft/exceptions.py:343-356: carries the kwargs needed to build the missing primitives."""

    def __init__(self, crypto_store=None, available_instances=-1, n_instances=1, op="", **kwargs):
        self.kwargs_ = {"op": op, "n_instances": n_instances, **kwargs}
        super().__init__(
            f"not enough '{op}' primitives in the store of {getattr(crypto_store, 'owner_id', '?')}: "
            f"asked {n_instances}, available {available_instances}; shapes={kwargs.get('shapes')}"
        )


def _key(shapes):
    return tuple(tuple(int(d) for d in s) for s in shapes)


class PrimitiveStorage:
    """Per-party store of Beaver triples keyed by (op, operand shapes) -- primitives.py:52-102.

    ``get_keys(remove=False)`` peeks (spdz_mask), ``remove=True`` pops (spdz_compute)."""

    def __init__(self, owner_id):
        self.owner_id = owner_id
        self.force_preprocessing = False
        self._stacks: Dict[str, Dict[tuple, List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]]] = {
            "mul": defaultdict(list),
            "matmul": defaultdict(list),
        }
        self._fss = []  # pools of DIF keys (ring/fss.py FSSKeys), consumed from the front

    def _get_fss_keys(self, n_instances: int, remove: bool):
        """FSS half of get_keys (primitives.py:52-102): the next ``n_instances`` comparison keys; ``remove`` burns them."""
        from .fss import OP, FSSKeys

        self._fss = [c for c in self._fss if c.available > 0]
        avail = sum(c.available for c in self._fss)
        if avail < n_instances:
            raise EmptyCryptoPrimitiveStoreError(self, avail if avail else -1, n_instances, op=OP)
        if self._fss[0].available < n_instances:  # the request spans pools: the reference keeps one concatenated array
            self._fss = [FSSKeys.cat(self._fss)]
        head = self._fss[0]
        win = head.window(n_instances)
        if remove:
            head.off += n_instances
            if head.available == 0:
                self._fss.pop(0)  # the window keeps the tensors alive until the evaluation that uses it has been issued
        return win

    def add_fss_keys(self, keys):
        self._fss.append(keys)

    def fss_available(self):
        return sum(c.available for c in self._fss)

    def get_keys(self, op: str, shapes=None, n_instances: int = 1, remove: bool = True, **kwargs):
        if op == "fss_comp":
            return self._get_fss_keys(n_instances, remove)
        stack = self._stacks[op][_key(shapes)]
        if len(stack) < n_instances:
            raise EmptyCryptoPrimitiveStoreError(self, len(stack) if stack else -1, n_instances, op=op, shapes=_key(shapes))
        assert n_instances == 1
        return stack.pop(0) if remove else stack[0]

    # ---- static-buffer support for CUDA-graph replay of the online phase (ring/resnet.py EncryptedInferenceGraph)
    def export_state(self):
        """the current content, by reference: ({op: {key: [triples]}}, [fss pools])"""
        return ({op: {k: list(v) for k, v in st.items()} for op, st in self._stacks.items()}, list(self._fss))

    def import_state(self, state):
        stacks, fss = state
        for op in self._stacks:
            self._stacks[op] = defaultdict(list, {k: list(v) for k, v in stacks.get(op, {}).items()})
        self._fss = list(fss)
        for c in self._fss:
            c.off = 0

    def clear(self):
        self.import_state(({}, []))

    @staticmethod
    def state_tensors(state):
        """every tensor of an exported state, in a deterministic order"""
        stacks, fss = state
        out = []
        for op in ("mul", "matmul"):
            for key, lst in stacks.get(op, {}).items():
                for tri in lst:
                    out.extend(tri)
        for c in fss:
            out.extend(c.tensors())
        return out

    def add_primitives(self, op: str, shapes, triples):
        """primitives.py:194-213"""
        self._stacks[op][_key(shapes)].extend(triples)

    def count(self, op, shapes):
        return len(self._stacks[op][_key(shapes)])

    def nbytes(self):
        return (sum(t.numel() * 8 for st in self._stacks.values() for lst in st.values() for tri in lst for t in tri)
                + sum(c.nbytes() for c in self._fss))


class Party:
    """A share holder / crypto provider pinned to one GPU (the VirtualWorker of path E)."""

    def __init__(self, id, device):
        self.id = id
        self.device = torch.device(device)
        self.crypto_store = PrimitiveStorage(id)

    def __repr__(self):
        return f"<Party {self.id} on {self.device}>"


class TripleProvider:
    """provide_primitives / build_triples / build_triple -- primitives.py:161-192,255-286, beaver.py:7-63.

    a, b ~ Philox over [-2^63, 2^63-2]; c = a (op) b computed with the ring GEMM on the provider's GPU;
    each of a,b,c is split into 2 additive shares which are copied to the parties' GPUs."""

    def __init__(self, provider: Party, seed: int = 0x5EED):
        self.provider = provider
        self.seed = seed
        self.counter = 0
        self.generated_bytes = 0
        self.request_log = []  # (op, shapes, n_instances) of every provide_primitives call, in order
        self._epoch = None

    @property
    def epoch(self):
        """device-side Philox epoch of THIS provider (ops.new_epoch): bumped by the captured offline graph of
        EncryptedInferenceGraph so that every replay generates fresh primitives"""
        if self._epoch is None:
            self._epoch = ops.new_epoch(self.provider.device)
        return self._epoch

    def bump_epoch(self):
        ops.bump_epoch(self.epoch)

    def _rand(self, shape):
        self.counter += 1
        return ops.random_i64(shape, self.seed, self.counter, self.provider.device, self.epoch)

    def build_triple(self, op: str, shapes):
        ls, rs = shapes
        a, b = self._rand(ls), self._rand(rs)
        if op == "matmul":
            c = ops.matmul(a, b)
        else:
            c = _mul_bcast(a, b)
        out = [[None] * 3, [None] * 3]
        for i, t in enumerate((a, b, c)):
            self.counter += 1
            s0, s1 = ops.share_gen(t, self.seed, self.counter, epoch=self.epoch)
            out[0][i], out[1][i] = s0, s1
        return out

    def on_triple(self, op, shapes, tri):
        """hook: called once per generated triple instance, in store order (tests record the randomness here)"""

    def build_fss_keys(self, n_instances: int):
        """build_fss_keys / build_separate_fss_keys -- primitives.py:237-253 (DIF.keygen on the provider's GPU)"""
        from .fss import build_fss_keys

        self.counter += 3
        return build_fss_keys(n_instances, self.provider.device, self.seed, self.counter - 2, self.epoch)

    def provide_primitives(self, op: str, shapes=None, parties=None, n_instances: int = 1, **_):
        self.request_log.append((op, shapes, n_instances))
        if op == "fss_comp":
            keys = self.build_fss_keys(n_instances)
            for j, p in enumerate(parties):
                k = keys[j] if p.device == self.provider.device else keys[j].to(p.device)
                self.generated_bytes += k.nbytes()
                p.crypto_store.add_fss_keys(k)
            self._settle(parties)
            return
        if op == "mul" and n_instances > 1 and tuple(shapes[0]) == tuple(shapes[1]):
            # build_triple with a leading n_instances axis (beaver.py:23-31), split into per-instance views
            bulk = ((n_instances, *shapes[0]), (n_instances, *shapes[1]))
            tri = self.build_triple(op, bulk)
            for j, p in enumerate(parties):
                moved = tuple(t.to(p.device, non_blocking=True) for t in tri[j])
                self.generated_bytes += sum(t.numel() * 8 for t in moved)
                p.crypto_store.add_primitives(op, shapes, [tuple(t[i] for t in moved) for i in range(n_instances)])
            for i in range(n_instances):
                self.on_triple(op, shapes, [tuple(t[i] for t in tri[j]) for j in range(2)])
            n_instances = 0
        for _i in range(n_instances):
            tri = self.build_triple(op, shapes)
            self.on_triple(op, shapes, tri)
            for j, p in enumerate(parties):
                moved = tuple(t.to(p.device, non_blocking=True) for t in tri[j])
                self.generated_bytes += sum(t.numel() * 8 for t in moved)
                p.crypto_store.add_primitives(op, shapes, [moved])
        self._settle(parties)

    def _settle(self, parties):
        """the freshly generated tensors die on the provider's GPU as soon as this returns; the copies to the parties must have
        read them first.  (Inside a CUDA-graph capture the copies are graph nodes ordered before anything that reuses the
        memory, and a device synchronisation would be illegal.)"""
        if any(p.device != self.provider.device for p in parties) and not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize(self.provider.device)


def take_primitives(op: str, shapes, n: int, parties, provider=None):
    """pop the next ``n`` triples of one kind from every party's store (n consecutive spdz_compute pops, spdz.py:84),
    asking the provider for the missing ones first as spdz_mul does (spdz.py:156-160)."""
    have = min(p.crypto_store.count(op, shapes) for p in parties)
    if have < n:
        if provider is None or any(p.crypto_store.force_preprocessing for p in parties):
            raise EmptyCryptoPrimitiveStoreError(parties[0].crypto_store, have, n, op=op, shapes=_key(shapes))
        provider.provide_primitives(op, shapes, parties, n - have)
    return [[p.crypto_store.get_keys(op=op, shapes=shapes, remove=True) for _ in range(n)] for p in parties]


def _mul_bcast(a, b):
    """elementwise ring product with the [C] x [P,C] broadcasts batch_norm needs (exact in int64)."""
    # use the combine kernel with delta=a, eps=b, a=0, b=0, c=0 on "party 0": z = delta*eps
    if a.shape == b.shape:
        zl, zr, zc = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(a)
    elif a.dim() == 1:
        zl, zr, zc = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(b)
    else:
        zl, zr, zc = torch.zeros_like(a), torch.zeros_like(b), torch.zeros_like(a)
    return ops.combine_mul(0, a, b, zl, zr, zc)


# ------------------------------------------------------------------ share level (run on each party)
def spdz_mask(party: Party, x, y, op: str):
    """spdz.py:22-45"""
    a, b, _c = party.crypto_store.get_keys(op=op, shapes=(x.shape, y.shape), n_instances=1, remove=False)
    return ops.mask(x, a), ops.mask(y, b)


def spdz_compute(party: Party, j: int, delta, epsilon, op: str):
    """spdz.py:64-122"""
    a, b, c = party.crypto_store.get_keys(op=op, shapes=(delta.shape, epsilon.shape), n_instances=1, remove=True)
    if op == "matmul":
        return ops.combine_matmul(j, delta, epsilon, a, b, c)
    return ops.combine_mul(j, delta, epsilon, a, b, c)


_PEER_READY = set()


def _ensure_peer(dev_a: torch.device, dev_b: torch.device) -> bool:
    """Make dev_b's memory dereferenceable from kernels running on dev_a (cudaDeviceEnablePeerAccess, which PyTorch
    performs on the first peer copy between the pair)."""
    key = (dev_a.index, dev_b.index)
    if key in _PEER_READY:
        return True
    if not torch.cuda.can_device_access_peer(dev_a.index, dev_b.index):
        return False
    torch.zeros(1, device=dev_b).to(dev_a)  # triggers peer-access enablement for the pair
    torch.zeros(1, device=dev_a).to(dev_b)
    torch.cuda.synchronize(dev_a)
    torch.cuda.synchronize(dev_b)
    _PEER_READY.add(key)
    return True


def open_shares(parties, shares):
    """delta = sum(shares) made available on every party (spdz.py:162-163), via peer reads."""
    outs = []
    for j, p in enumerate(parties):
        peer = shares[1 - j]
        if peer.device != p.device:
            # symmetric NVLink exchange: the add kernel on party j's GPU loads the peer's share through its peer-mapped
            # device pointer; the producer's stream must have finished writing it first
            torch.cuda.current_stream(p.device).wait_stream(torch.cuda.current_stream(peer.device))
            if not _ensure_peer(p.device, peer.device):
                peer = peer.to(p.device)  # staged copy when no P2P mapping exists
        outs.append(ops.open_add(shares[j], peer))
    release_after_peer_reads(parties)
    return outs


def open_planes(parties, shares):
    """``open_shares`` for a value that is only ever used as the LEFT operand of the limb-plane ring GEMM: each party's kernel
    adds the peer's masked share (peer-mapped pointer when it sits on another GPU) and writes the byte planes of the sum."""
    outs = []
    for j, p in enumerate(parties):
        peer = shares[1 - j]
        if peer.device != p.device:
            torch.cuda.current_stream(p.device).wait_stream(torch.cuda.current_stream(peer.device))
            if not _ensure_peer(p.device, peer.device):
                peer = peer.to(p.device)
        outs.append(ops.planarize_rows(shares[j], peer))
    release_after_peer_reads(parties)
    return outs


def release_after_peer_reads(parties):
    """After a symmetric exchange each GPU has just read the other's buffer through a peer pointer.  The caching allocator
    only orders a block's reuse on the OWNER's stream, so the owner's stream is made to wait for the reader's kernel: without
    this edge a later kernel of the owner could overwrite a freed share while the peer is still loading it."""
    devs = []
    for p in parties:
        if p.device not in devs:
            devs.append(p.device)
    if len(devs) < 2:
        return
    evs = {}
    for d in devs:
        with torch.cuda.device(d):
            evs[d] = torch.cuda.current_stream(d).record_event()
    for d in devs:
        for o in devs:
            if o != d:
                torch.cuda.current_stream(d).wait_event(evs[o])


# PRIMIA_FUSE_OPEN=0: every opening is its own launch (mask, mask, open, open, combine -- the reference's message pattern one to
# one); 1: same-shape products mask both operands in one pass and open them inside the combine kernel; DIF.eval opens its input
FUSE_OPEN = os.environ.get("PRIMIA_FUSE_OPEN", "1") != "0"


def peer_views(parties, shares):
    """the event edges of ``open_shares`` for kernels that read the peer's masked share themselves: per party, the peer's
    tensor (peer-mapped pointer, or a staged copy when no P2P mapping exists)"""
    peers = []
    for j, p in enumerate(parties):
        peer = shares[1 - j]
        if peer.device != p.device:
            torch.cuda.current_stream(p.device).wait_stream(torch.cuda.current_stream(peer.device))
            if not _ensure_peer(p.device, peer.device):
                peer = peer.to(p.device)
        peers.append(peer)
    return peers


def spdz_mul(op: str, x_shares, y_shares, parties, provider: TripleProvider = None):
    """spdz.py:125-197.  x_shares / y_shares: per-party tensors. Returns per-party output shares."""
    if FUSE_OPEN and op == "mul" and x_shares[0].shape == y_shares[0].shape:
        shapes = (x_shares[0].shape, y_shares[0].shape)
        try:
            tri = [p.crypto_store.get_keys(op=op, shapes=shapes, n_instances=1, remove=False) for p in parties]
        except EmptyCryptoPrimitiveStoreError as e:
            if provider is None or any(p.crypto_store.force_preprocessing for p in parties):
                raise
            provider.provide_primitives(parties=parties, **e.kwargs_)
            return spdz_mul(op, x_shares, y_shares, parties, provider)
        masked = [ops.mask2(x_shares[j], tri[j][0], y_shares[j], tri[j][1]) for j in range(2)]
        pd = peer_views(parties, [m[0] for m in masked])
        pe = peer_views(parties, [m[1] for m in masked])
        out = []
        for j, p in enumerate(parties):
            a, b, c = p.crypto_store.get_keys(op=op, shapes=shapes, n_instances=1, remove=True)
            out.append(ops.combine_mul_open(j, masked[j][0], pd[j], masked[j][1], pe[j], a, b, c))
        release_after_peer_reads(parties)
        return out
    try:
        masked = [spdz_mask(p, x_shares[j], y_shares[j], op) for j, p in enumerate(parties)]
    except EmptyCryptoPrimitiveStoreError as e:
        if provider is None or any(p.crypto_store.force_preprocessing for p in parties):
            raise
        provider.provide_primitives(parties=parties, **e.kwargs_)
        return spdz_mul(op, x_shares, y_shares, parties, provider)
    deltas = open_shares(parties, [m[0] for m in masked])
    epsilons = open_shares(parties, [m[1] for m in masked])
    return [spdz_compute(p, j, deltas[j], epsilons[j], op) for j, p in enumerate(parties)]
