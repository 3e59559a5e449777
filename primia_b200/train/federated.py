"""Federated round on GPUs: one hospital (PySyft VirtualWorker) == one ResNet18Engine pinned to one GPU.

Mirrors torchlib/utils.py:
  secure_aggregation_epoch :1108-1233  -> federated_round
  aggregation              :1000-1092  -> aggregation  (sum over hospitals of the flat state, / n or weighted)
  send_new_models          :1095-1105  -> the all-reduce result is already resident on every GPU
  optimizer reset          :1131-1145,1209-1218 -> ResNet18Engine.reset_optimizer
Two deployments:
  * multi-process (one rank per GPU, torch.distributed NCCL): FedAvg = all_reduce(sum) of ``engine.flat``
    over NVLink followed by ``pm_scale_f32`` (weighted averaging pre-scales by w_i, utils.py:953-957,1051-1055);
  * single-process (several hospitals time-sharing one GPU, as the reference's VirtualWorkers share one
    process): the same sum is formed locally with ``pm_scale_f32`` + in-place adds.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from .._lib import call, ptr, stream
from .resnet18 import ResNet18Engine


class HospitalWorker:
    """A data owner: its engine (model + optimizer state) and its local batches."""

    def __init__(self, id: str, engine: ResNet18Engine, dp: Optional[dict] = None):
        self.id = id
        self.engine = engine
        self.batches: List = []
        self.dp = dp  # {"noise_multiplier": .., "max_grad_norm": ..}: the PrivacyEngine attached to this hospital's optimizer

    def local_step(self, data, target):
        """utils.py:1168-1174 (with a privacy engine attached, train.py:326-334: the DP-SGD step of primia_b200/train/dp.py)"""
        if self.dp is not None:
            from .dp import dp_train_step

            return dp_train_step(self.engine, data, target, **self.dp)
        return self.engine.train_step(data, target)

    def local_step_and_fedavg(self, data, target, group, host=False):
        """utils.py:1168-1201 for the reference's default schedule (sync after every batch, unweighted, plain aggregation) with the
        all-reduce OVERLAPPED with the backward pass: the engine's two-graph step (``capture_graph_overlap``) finishes layer4's
        gradients and optimizer step first; that bucket (75 % of the 44.75 MB state, plus the BatchNorm running statistics) is
        averaged over NVLink (ncclAvg) on NCCL's stream while the rest of the backward runs; the small remainder follows."""
        import torch.distributed as dist

        eng = self.engine
        off = eng.split_offset
        works = []
        if host:
            self._slots(data, target)
            if self._pending is None or self._pending[1] is not data:
                self.prefetch_host(data, target)
            slot = self._pending[0]
            self._pending = None
            torch.cuda.current_stream(eng.device).wait_event(self._ready[slot])
            data, target = self._stage[slot]

            def staged():
                ev = torch.cuda.Event()
                ev.record()
                self._consumed[slot] = ev

            eng.on_inputs_staged = staged
        with torch.cuda.device(eng.device):
            try:
                loss = eng.train_step_overlapped(
                    data, target,
                    lambda: works.append(dist.all_reduce(eng.flat[off:], op=dist.ReduceOp.AVG, group=group, async_op=True)),
                    lambda: works.append(dist.all_reduce(eng.flat[:off], op=dist.ReduceOp.AVG, group=group, async_op=True)))
            finally:
                eng.on_inputs_staged = None
            for w in works:
                w.wait()   # stream-level wait: the next kernels on this stream see the averaged state
        return loss

    def _slots(self, host_data, host_target):
        eng = self.engine
        if getattr(self, "_stage", None) is None or self._stage[0][0].shape != host_data.shape:
            mk = lambda h: torch.empty(h.shape, dtype=h.dtype, device=eng.device)
            self._stage = [(mk(host_data), mk(host_target)) for _ in range(2)]
            self._copy_stream = torch.cuda.Stream(eng.device)
            self._ready = [None, None]     # event: slot filled
            self._consumed = [None, None]  # event: the step that read the slot was enqueued and finished reading
            self._pending = None           # (slot, host_data, host_target) staged ahead of time
            self._next_slot = 0
        return self._stage

    def prefetch_host(self, host_data, host_target):
        """Start the host->device copy of the NEXT batch (pinned memory) on a copy stream, so it overlaps the step that is
        running -- what a data loader with one batch of look-ahead does."""
        self._slots(host_data, host_target)
        slot = self._next_slot
        self._next_slot ^= 1
        with torch.cuda.device(self.engine.device):
            if self._consumed[slot] is not None:
                self._copy_stream.wait_event(self._consumed[slot])
            with torch.cuda.stream(self._copy_stream):
                self._stage[slot][0].copy_(host_data, non_blocking=True)
                self._stage[slot][1].copy_(host_target, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            self._ready[slot] = ev
        self._pending = (slot, host_data, host_target)

    def local_step_host(self, host_data, host_target):
        """The same local step fed from HOST (pinned) memory: the batch crosses PCIe inside the call unless it was already
        prefetched with ``prefetch_host``; the returned loss is a device scalar the caller reads back."""
        eng = self.engine
        self._slots(host_data, host_target)
        if self._pending is None or self._pending[1] is not host_data:
            self.prefetch_host(host_data, host_target)
        slot = self._pending[0]
        self._pending = None
        with torch.cuda.device(eng.device):
            torch.cuda.current_stream().wait_event(self._ready[slot])
            self._consumed[slot] = None

            def staged():   # graph replays copy the batch into their static input first: the slot is free from then on
                ev = torch.cuda.Event()
                ev.record()
                self._consumed[slot] = ev

            eng.on_inputs_staged = staged
            try:
                loss = self.local_step(self._stage[slot][0], self._stage[slot][1])
            finally:
                eng.on_inputs_staged = None
            if self._consumed[slot] is None:   # eager step: the stem reads the batch until the end of the backward pass
                staged()
        return loss


def fedavg_scales(worker_id, n_workers: int, weights: Optional[Dict[str, float]]):
    """(pre, post) factors around the SUM all-reduce: unweighted FedAvg divides by n afterwards (utils.py:1090);
    weighted averaging multiplies each hospital's state by w_i before the sum (utils.py:1051-1055)."""
    if weights is not None:
        return float(weights[worker_id]), 1.0
    return 1.0, 1.0 / n_workers


def _scale(t: torch.Tensor, scale: float):
    if scale == 1.0:
        return
    with torch.cuda.device(t.device):
        call("pm_scale_f32", ptr(t), ctypes.c_float(scale), t.numel(), stream())


def _telescoping_shares(q: torch.Tensor, n: int, seed: int, counter: int):
    """AdditiveSharingTensor.generate_shares for n workers (additive_shared.py:336-365): r_0..r_{n-2} random,
    shares = [r_0, r_1 - r_0, ..., r_{n-2} - r_{n-3}, q - r_{n-2}]  (sum telescopes to q mod 2^64)."""
    from ..ring import ops

    if n == 1:
        return [q]
    rs = [ops.random_i64(q.shape, seed, counter + i, q.device) for i in range(n - 1)]
    shares = [rs[0]]
    for i in range(1, n - 1):
        shares.append(ops.axpby(1, rs[i], -1, rs[i - 1]))
    shares.append(ops.axpby(1, q, -1, rs[n - 2]))
    return shares


def _mask_stream(worker: HospitalWorker, n: int, seed: Optional[int]):
    """(seed, first counter) of the Philox stream this hospital masks its state with in ONE aggregation.  The seed is a
    per-hospital secret drawn from the OS (never derivable by a peer) and the counter only moves forward, so no pad is ever
    reused across rounds -- a reused pad would hand the receiving hospital the difference of two rounds' weights, a public
    one the weights themselves.  ``seed`` (tests only) pins the stream."""
    if seed is not None:
        return seed, 1
    if getattr(worker, "_mask_seed", None) is None:
        import secrets

        worker._mask_seed, worker._mask_counter = secrets.randbits(63), 1
    c = worker._mask_counter
    worker._mask_counter += max(n - 1, 1)
    return worker._mask_seed, c


def secure_aggregation(workers: List[HospitalWorker], weights: Optional[Dict[str, float]] = None, group=None,
                       precision_fractional: int = 16, base: int = 10, seed: Optional[int] = None):
    """The reference's DEFAULT aggregation (``secure=True``, utils.py:1045-1060,1078-1090) on GPUs:

        (param * w_i).fix_prec(pf).share(*workers).get()  ->  share-wise sum over hospitals  ->  .get()  ->  .float_prec()  [-> / n]

    Every hospital fixed-point-encodes its flat state (pm_encode_f32_i64), splits it into one additive share per hospital
    (Philox), share j travels to hospital j (all-to-all over NVLink when one hospital runs per rank), hospital j adds the
    shares it received -- a hospital only ever sees uniformly random shares of another hospital's weights, masked with pads
    from that hospital's own secret Philox stream (``_mask_stream``) -- and the per-hospital sums are
    reconstructed with an int64 SUM all-reduce (two's-complement wraparound == arithmetic mod 2^64) and decoded.
    Result (identical on every hospital) == decode(sum_i encode(theta_i * w_i)) [/ n], the reference's arithmetic."""
    import torch.distributed as dist

    from ..ring import ops

    distributed = group is not None
    n = dist.get_world_size(group) if distributed else len(workers)
    rank = dist.get_rank(group) if distributed else 0
    encoded = []
    for w in workers:
        eng = w.engine
        x = eng.flat
        if weights is not None:
            x = x.clone()
            _scale(x, float(weights[w.id]))  # (param * weight) in fp32, then encoded
        with torch.cuda.device(eng.device):
            encoded.append(ops.encode(x, base, precision_fractional))
    if distributed:
        q = encoded[0]
        shares = _telescoping_shares(q, n, *_mask_stream(workers[0], n, None if seed is None else seed + rank))
        send = torch.stack(shares)                      # [n, len]: row j goes to hospital j
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        mine = recv[0]
        for j in range(1, n):
            mine = ops.axpby(1, mine, 1, recv[j])        # this hospital's share of the sum
        dist.all_reduce(mine, op=dist.ReduceOp.SUM, group=group)  # reconstruction (mod 2^64)
        total = mine
    else:
        dev = workers[0].engine.device
        held = [None] * n                                # held[j]: what hospital j accumulates
        for i, q in enumerate(encoded):
            for j, sh in enumerate(_telescoping_shares(q, n, *_mask_stream(workers[i], n, None if seed is None else seed + i))):
                sh = sh.to(workers[j].engine.device)
                held[j] = sh if held[j] is None else ops.axpby(1, held[j], 1, sh)
        total = held[0].to(dev)
        for j in range(1, n):
            total = ops.axpby(1, total, 1, held[j].to(dev))
    for w in workers:
        eng = w.engine
        with torch.cuda.device(eng.device):
            out = ops.decode(total.to(eng.device), base, precision_fractional)
            eng.flat.copy_(out)
            if weights is None:
                # sumstacked / len(workers), utils.py:1090: IEEE fp32 division (a tensor divisor; a python-scalar divisor would
                # be turned into a multiplication by the rounded reciprocal on CUDA and differ in the last bit for n = 3, 5, ...)
                eng.flat.div_(torch.full((), float(n), dtype=torch.float32, device=eng.device))


def aggregation(workers: List[HospitalWorker], weights: Optional[Dict[str, float]] = None, group=None, secure: bool = False,
                precision_fractional: int = 16):
    """FedAvg of every state entry except num_batches_tracked (utils.py:1027-1092).

    With ``group`` (torch.distributed process group, one hospital per rank) this is one NCCL all-reduce of the
    flat state; otherwise the hospitals of this process are reduced locally.  Returns nothing: every engine ends
    with the averaged state (== aggregation + send_new_models)."""
    import torch.distributed as dist

    if group is None and dist.is_available() and dist.is_initialized() and len(workers) == 1:
        group = dist.group.WORLD
    if secure:
        return secure_aggregation(workers, weights, group, precision_fractional)
    if group is not None:
        eng = workers[0].engine
        pre, post = fedavg_scales(workers[0].id, dist.get_world_size(group), weights)
        _scale(eng.flat, pre)
        if weights is None and dist.get_backend(group) == "nccl":
            # sum / n inside the collective (ncclAvg): one pass less over the 44.75 MB state than all-reduce + pm_scale_f32
            dist.all_reduce(eng.flat, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(eng.flat, op=dist.ReduceOp.SUM, group=group)
            _scale(eng.flat, post)
        return
    n = len(workers)
    acc = workers[0].engine.flat
    if weights is not None:
        _scale(acc, float(weights[workers[0].id]))
    for w in workers[1:]:
        src = w.engine.flat
        if weights is not None:
            _scale(src, float(weights[w.id]))
        acc.add_(src.to(acc.device))  # ring-free local reduction; torch add is plumbing, not the hot path
    if weights is None:
        _scale(acc, 1.0 / n)
    for w in workers[1:]:
        w.engine.flat.copy_(acc)


def federated_round(workers: List[HospitalWorker], sync_every_n_batch: int = 1, weights=None, keep_optim_dict=False,
                    group=None, secure: bool = False, precision_fractional: int = 16, opt_lr: Optional[float] = None):
    """secure_aggregation_epoch (utils.py:1108-1233).  Hospitals of this process are visited in order (utils.py:1160); with
    one hospital per rank they run concurrently on their own GPUs.

    Reference semantics kept: (1) optimizers are re-created at the start of the epoch and after every mid-epoch aggregation
    unless ``keep_optim_dict`` -- WITH ``lr = args.lr`` (utils.py:1131-1145,1209-1218; pass it as ``opt_lr``), which discards
    whatever the epoch's learning-rate schedule had set; (2) a hospital that has run out of batches is skipped
    (:1166-1167), still contributes its (stale) model to every later aggregation, but receives the new model only at the end
    of the epoch (send_new_models is restricted to ``num_batches[w] > batch_idx``, :1187-1194); (3) the returned loss is the
    mean over every local step of every hospital.  With one hospital per rank the number of rounds is the maximum over ranks
    and exhausted ranks keep taking part in the collectives."""
    import torch.distributed as dist

    def reset(w):
        w.engine.reset_optimizer()
        if opt_lr is not None:
            w.engine.lr = opt_lr

    if not keep_optim_dict:
        for w in workers:
            reset(w)
    losses = []
    nb = {w.id: len(w.batches) for w in workers}
    max_b = max(nb.values())
    dev = workers[0].engine.device
    if group is not None:
        t = torch.tensor([max_b], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        max_b = int(t.item())
    for batch_idx in range(max_b):
        for w in workers:
            if batch_idx >= nb[w.id]:
                continue
            d, t = w.batches[batch_idx]
            losses.append(w.local_step(d, t).clone())  # engine.loss is a reused device buffer
        if batch_idx > 0 and batch_idx % sync_every_n_batch == 0:
            done = [w for w in workers if nb[w.id] <= batch_idx]
            stale = {w.id: w.engine.flat.clone() for w in done}
            aggregation(workers, weights, group, secure, precision_fractional)
            for w in done:  # not in send_new_models' recipient list: keeps the model it finished with
                w.engine.flat.copy_(stale[w.id])
            if not keep_optim_dict:
                for w in workers:
                    reset(w)
    aggregation(workers, weights, group, secure, precision_fractional)
    tot = torch.zeros(2, dtype=torch.float64, device=dev)
    if losses:
        tot[0] = torch.stack([l.reshape(()).to(dev).double() for l in losses]).sum()
        tot[1] = len(losses)
    if group is not None:
        dist.all_reduce(tot, group=group)
    return (tot[0] / tot[1].clamp_min(1)).float()
