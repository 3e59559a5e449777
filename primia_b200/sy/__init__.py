"""PySyft-shaped surface for the two hot paths (SURVEY.md section 8b "verbs the entry points actually use").

    import primia_b200.sy as sy
    hook = sy.TorchHook(torch)
    alice = sy.VirtualWorker(hook, id="alice")            # pinned to one GPU
    ptr = tensor.tag("#traindata").send(alice)            # PointerTensor: the data lives on alice's GPU
    grid = sy.PrivateGridNetwork(alice, bob).search("#traindata")
    loader = sy.FederatedDataLoader(sy.FederatedDataset([sy.BaseDataset(data_ptr, target_ptr)]), batch_size=64, shuffle=True)
    enc = tensor.fix_precision(precision_fractional=16).share(model_owner, data_owner, crypto_provider=cp, protocol="fss")

The reference implements these verbs by monkey-patching every torch function and shipping msgpack messages between
in-process workers (syft/frameworks/torch/hook/hook.py, syft/workers/base.py:290-363).  Here a worker is a GPU: ``send``
is a device copy, pointer ops run on the owner's GPU, and the arithmetic behind ``.fix_precision().share()`` and the
training step is the primia_b200 C ABI.  Only the verbs train.py / inference.py use are provided.
"""
from __future__ import annotations

import torch

from .. import ring
from ..ring.spdz import Party, TripleProvider
from ..ring.tensors import AdditiveSharingTensor, FixedPrecisionTensor  # noqa: F401

hook = None
local_worker = None


class _ObjectStore:
    def __init__(self):
        self.objects = {}
        self.garbage_delay = 0

    def clear_objects(self):
        self.objects.clear()

    def register(self, obj):
        self.objects[id(obj)] = obj
        return obj


class VirtualWorker(Party):
    """syft/workers/virtual.py:8-22 -> one GPU.  ``device`` defaults to round-robin over the visible GPUs."""

    _next = 0

    def __init__(self, hook=None, id=None, verbose=False, device=None, **_):
        if device is None:
            n = max(torch.cuda.device_count(), 1)
            device = f"cuda:{VirtualWorker._next % n}"
            VirtualWorker._next += 1
        super().__init__(id, device)
        self.verbose = verbose
        self.object_store = _ObjectStore()
        self.clients = []

    def load_data(self, tensors):
        """base.py:229 -- register tagged tensors so PrivateGridNetwork.search finds them"""
        for t in tensors:
            self.object_store.register(t)

    def search(self, tag):
        return [PointerTensor(self, t) for t in self.object_store.objects.values() if tag in getattr(t, "_sy_tags", ())]


class PointerTensor:
    """syft/generic/pointers/pointer_tensor.py: a handle to a tensor living on ``location``'s GPU.  The verbs the entry points
    call on pointers (inference.py:298-311: shape / squeeze / unsqueeze / to / fix_precision / share / copy / get) run on the
    owner's GPU and return pointers; ``get`` hands the object over."""

    def __init__(self, location: VirtualWorker, tensor):
        self.location = location
        self._t = tensor

    @property
    def shape(self):
        return self._t.shape

    def get(self):
        return self._t

    def copy(self):
        return PointerTensor(self.location, self._t.clone())

    def __len__(self):
        return self._t.shape[0]

    def __getitem__(self, idx):
        return PointerTensor(self.location, self._t[idx])

    def squeeze(self, *a):
        return PointerTensor(self.location, self._t.squeeze(*a))

    def unsqueeze(self, dim):
        return PointerTensor(self.location, self._t.unsqueeze(dim))

    def to(self, *_a, **_k):
        return self  # the data stays on its owner's GPU

    def fix_precision(self, **kw):
        return PointerTensor(self.location, _fix_precision(self._t, **kw))

    fix_prec = fix_precision

    def share(self, *workers, **kw):
        return PointerTensor(self.location, self._t.share(*workers, **kw))


def _tag(self, *tags):
    self._sy_tags = tuple(getattr(self, "_sy_tags", ())) + tags
    return self


def _send(self, worker):
    t = self.to(worker.device, non_blocking=True)
    t._sy_tags = getattr(self, "_sy_tags", ())
    worker.object_store.register(t)
    return PointerTensor(worker, t)


def _fix_precision(self, precision_fractional=3, dtype="long", base=10, **_):
    """native.fix_prec native.py:835-864 -> FixedPrecisionTensor.fix_precision precision.py:117-132"""
    if dtype != "long":
        raise NotImplementedError("the ring is Z_2^64 (dtype='long'), as inference.py:280 uses")
    x = self if self.is_cuda else self.cuda()
    return FixedPrecisionTensor.fix_precision(x.detach().float().contiguous(), base, precision_fractional)


def _share(self, *workers, crypto_provider=None, protocol="fss", requires_grad=False, **_):
    """native.share native.py:887-949 on a raw tensor: only integer tensors can be additively shared (:931-932)"""
    if self.is_floating_point():
        raise TypeError("FloatTensor cannot be additively shared, Use fix_precision.")
    if protocol != "fss":
        raise NotImplementedError("protocol 'fss' (function secret sharing) is the built comparison protocol; 'snn' is not")
    x = (self if self.is_cuda else self.cuda()).to(torch.int64).contiguous()
    return AdditiveSharingTensor.share_secret(x, list(workers), ring.tensors.provider_of(crypto_provider))


# ---- torch.nn.Module verbs (hook.py:613-807).  PySyft replaces every parameter AND buffer in place (tensor_iterator :626-632
# iterates ``parameters`` and ``buffers``); torch 2 Parameters cannot hold our share objects, so the converted tensors live in
# ``module._sy_shared`` (state_dict key -> FixedPrecisionTensor) beside the untouched fp32 Parameters.
def _named_tensors(module):
    for k, v in module.named_parameters():
        yield k, v
    for k, v in module.named_buffers():
        yield k, v


def _module_fix_precision(self, **kw):
    """module_fix_precision_ hook.py:738-765"""
    self._sy_shared = {k: _fix_precision(v.data, **kw) for k, v in _named_tensors(self)
                       if v.is_floating_point()}  # num_batches_tracked (int64 counter) is never read by the eval forward
    self._sy_state = "fixed"
    return self


def _module_share(self, *workers, **kw):
    """module_share_ hook.py:767-782: every parameter and buffer .share()d in place"""
    if getattr(self, "_sy_state", None) != "fixed":
        raise RuntimeError("call model.fix_precision(...) before model.share(...) (FloatTensor cannot be additively shared)")
    self._sy_shared = {k: v.share(*workers, **kw) for k, v in self._sy_shared.items()}
    self._sy_state = "shared"
    return self


def _module_float_precision(self):
    """module_float_precision_ hook.py:784-796 (after .get())"""
    for k, v in _named_tensors(self):
        if k in getattr(self, "_sy_shared", {}):
            v.data.copy_(self._sy_shared[k].float_precision().reshape(v.shape))
    self._sy_shared, self._sy_state = None, None
    return self


def _module_send(self, *dest, **_):
    """module_send_ hook.py:650-664: parameters and buffers move to the worker's GPU; the module remembers its location"""
    (worker,) = dest
    if any(True for _ in self.parameters()) or any(True for _ in self.buffers()):
        self.to(worker.device)
    self.location = worker
    worker.object_store.register(self)
    return self


def _module_get(self):
    """module_get_ hook.py:693-706: for shared modules reconstruct every tensor, otherwise just drop the location"""
    if getattr(self, "_sy_state", None) == "shared":
        self._sy_shared = {k: v.get() for k, v in self._sy_shared.items()}
        self._sy_state = "fixed"
    self.location = None
    return self


def _module_copy(self):
    """module_copy hook.py:798-807"""
    import copy

    # engines / encrypted evaluators are caches bound to device buffers: the copy starts without them
    held = {k: self.__dict__.pop(k) for k in ("_engines", "_encrypted") if k in self.__dict__}
    try:
        new = copy.deepcopy(self)
    finally:
        self.__dict__.update(held)
    if "_engines" in held:
        new._engines, new._encrypted = {}, None
    return new


class TorchHook:
    """syft/frameworks/torch/hook/hook.py: grafts the PySyft verbs this port supports onto torch.Tensor."""

    def __init__(self, torch_module=torch, **_):
        global hook, local_worker
        T = torch_module.Tensor
        T.tag = _tag
        T.send = _send
        T.fix_precision = _fix_precision
        T.fix_prec = _fix_precision
        T.share = _share
        M = torch_module.nn.Module
        M.fix_precision = M.fix_prec = _module_fix_precision
        M.share = _module_share
        M.float_precision = M.float_prec = _module_float_precision
        M.send = _module_send
        M.get = _module_get
        M.copy = _module_copy
        if not hasattr(M, "location"):
            M.location = None
        hook = self
        local_worker = VirtualWorker(self, id="me", device="cuda:0" if torch.cuda.is_available() else "cpu")
        self.local_worker = local_worker


class PrivateGridNetwork:
    """syft/grid/private_grid.py:24"""

    def __init__(self, *workers):
        self.workers = workers

    def search(self, *query):
        out = {}
        for w in self.workers:
            found = [p for tag in query for p in w.search(tag)]
            if found:
                out[w.id] = found
        return out


class BaseDataset:
    """syft/frameworks/torch/fl/dataset.py:15"""

    def __init__(self, data, targets, transform=None):
        self.data, self.targets = data, targets
        self.location = data.location

    def __len__(self):
        return len(self.data)


class FederatedDataset:
    """syft/frameworks/torch/fl/dataset.py:288"""

    def __init__(self, datasets):
        self.datasets = {d.location.id: d for d in datasets}
        self.workers = list(self.datasets)

    def __len__(self):
        return sum(len(d) for d in self.datasets.values())


class FederatedDataLoader:
    """syft/frameworks/torch/fl/dataloader.py:159-258 for the single-worker datasets PriMIA builds
    (utils.py:753-763): RandomSampler on the owner's GPU, batches gathered there (no per-sample messages)."""

    def __init__(self, federated_dataset, batch_size=8, shuffle=False, drop_last=False, seed=0, **_):
        self.federated_dataset = federated_dataset
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self._gen = None
        self._seed = seed

    def __len__(self):
        n = len(self.federated_dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        (ds,) = self.federated_dataset.datasets.values()
        data, targets = ds.data.get(), ds.targets.get()
        n = data.shape[0]
        if self._gen is None:
            self._gen = torch.Generator(device=data.device).manual_seed(self._seed)
        order = torch.randperm(n, device=data.device, generator=self._gen) if self.shuffle else torch.arange(n, device=data.device)
        for i in range(len(self)):
            idx = order[i * self.batch_size:(i + 1) * self.batch_size]
            yield PointerTensor(ds.location, data[idx]), PointerTensor(ds.location, targets[idx])


class RemoteTensorDataset:
    """torchlib/dataloader.py RemoteTensorDataset (inference.py:229): items are pointers to single images on the data owner"""

    def __init__(self, tensor: PointerTensor):
        self.tensor = tensor

    def __len__(self):
        return len(self.tensor)

    def __getitem__(self, i):
        return self.tensor[i].copy()


class serde:  # inference.py:37-39 touches sy.serde.compression.* : accepted and ignored (no bytes are serialised here)
    class compression:
        NO_COMPRESSION = 40
        default_compress_scheme = 40


def make_crypto_provider(worker: VirtualWorker, seed=None) -> TripleProvider:
    """the TripleProvider that generates Beaver triples / FSS keys on ``worker``'s GPU (one per worker, created on first use)"""
    return ring.tensors.provider_of(worker, seed)
