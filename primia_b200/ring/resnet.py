"""Encrypted ResNet-18 forward on shares (inference.py:279-321): layer schedule, geometry, offline / online split.

``EncryptedResNet18`` is the whole forward of the reference's encrypted inference -- 20 Beaver convolutions + fc (rows E4-E10 of
SURVEY.md section 8a), 20 BatchNorms on shares with the 80-step Newton inverse square root (E11), 17 ReLUs and the 3x3 max-pool
through function secret sharing (E12, E13), average pool, reconstruction and decoding (E14) -- share for share equal to the
oracle (tests/test_fss_gpu.py, at 32x32 / pf 4 and at the reference's 224x224 / pf 16).  ``EncryptedInferenceGraph`` captures
its online phase in a CUDA graph (single- or multi-GPU placement).  ``SharedLinearLayers`` / ``EncryptedLinearGraph`` run the
21 linear layers alone on synthetic activation shares: the Beaver-matmul micro-benchmark bench.py reports beside the full
forward."""
from __future__ import annotations

import os

import torch

from . import functional as F
from .spdz import TripleProvider
from .tensors import FixedPrecisionTensor

# PRIMIA_HOIST_WEIGHT_SIDE=0: the online graph runs the whole protocol per layer (mask and open the weights, planarise all four
# operands, Newton iterations, op-by-op BatchNorm) as the reference does per call; 1: everything that depends only on the model
# and the primitives moves to the offline graph (EncryptedResNet18.prepare_offline_side)
HOIST_WEIGHT_SIDE = os.environ.get("PRIMIA_HOIST_WEIGHT_SIDE", "1") != "0"
HOIST_PARTS = set(os.environ.get("PRIMIA_HOIST_PARTS", "newton,conv,bn").split(","))   # bisecting aid: which halves are hoisted

# (name, Cin, H_in, Cout, k, stride, pad) for a 224x224 input -- torchlib/models.py:379-405,425-464
RESNET18_CONVS = [("conv1", 3, 224, 64, 7, 2, 3)]
_h, _c = 56, 64
for _li, (_planes, _stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
    for _bi in range(2):
        _st = _stride if _bi == 0 else 1
        RESNET18_CONVS.append((f"layer{_li}.{_bi}.conv1", _c, _h, _planes, 3, _st, 1))
        _ho = (_h + 2 - 3) // _st + 1
        RESNET18_CONVS.append((f"layer{_li}.{_bi}.conv2", _planes, _ho, _planes, 3, 1, 1))
        if _st != 1 or _c != _planes:
            RESNET18_CONVS.append((f"layer{_li}.{_bi}.downsample.0", _c, _h, _planes, 1, _st, 0))
        _c, _h = _planes, _ho


def conv_geometry(input_size=224):
    """RESNET18_CONVS for any input size: (name, Cin, H_in, Cout, k, stride, pad), walking the forward's own size arithmetic"""
    h = (input_size + 2 * 3 - 7) // 2 + 1
    out = [("conv1", 3, input_size, 64, 7, 2, 3)]
    h = (h + 2 - 3) // 2 + 1                      # max-pool 3x3 s2 p1
    c = 64
    for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for bi in range(2):
            st = stride if bi == 0 else 1
            out.append((f"layer{li}.{bi}.conv1", c, h, planes, 3, st, 1))
            ho = (h + 2 - 3) // st + 1
            out.append((f"layer{li}.{bi}.conv2", planes, ho, planes, 3, 1, 1))
            if st != 1 or c != planes:
                out.append((f"layer{li}.{bi}.downsample.0", c, h, planes, 1, st, 0))
            c, h = planes, ho
    return out


def triple_shapes(batch=1, num_classes=3):
    """Beaver-triple shapes one encrypted image consumes (reference im2col-shaped: spdz.py:36-38)."""
    out = []
    for name, C, H, Co, k, s, p in RESNET18_CONVS:
        Ho = (H + 2 * p - k) // s + 1
        out.append((name, ((batch, Ho * Ho, C * k * k), (C * k * k, Co))))
    out.append(("fc", ((batch, 512), (512, num_classes))))
    return out


class SharedLinearLayers:
    """Holds the 2-party sharing of the 21 weight tensors and runs every linear layer's protocol."""

    def __init__(self, parties, provider: TripleProvider, base=10, precision_fractional=16, num_classes=3, seed=42):
        self.parties, self.provider = parties, provider
        self.base, self.pf, self.ncls = base, precision_fractional, num_classes
        dev = parties[0].device
        g = torch.Generator().manual_seed(seed)
        self.weights = {}
        for name, C, H, Co, k, s, p in RESNET18_CONVS:
            w = (torch.randn(Co, C, k, k, generator=g) * (2.0 / (Co * k * k)) ** 0.5).to(dev)
            self.weights[name] = FixedPrecisionTensor.fix_precision(w, base, precision_fractional).share(
                *parties, crypto_provider=provider)
        self.fc_w = FixedPrecisionTensor.fix_precision((torch.randn(num_classes, 512, generator=g) * 0.04).to(dev), base,
                                                       precision_fractional).share(*parties, crypto_provider=provider)
        self.fc_b = FixedPrecisionTensor.fix_precision(torch.zeros(num_classes, device=dev), base,
                                                       precision_fractional).share(*parties, crypto_provider=provider)

    def make_inputs(self, batch=1, seed=7):
        """synthetic activation shares of every layer's input shape (encode + share on GPU)."""
        dev = self.parties[0].device
        g = torch.Generator().manual_seed(seed)
        xs = {}
        for name, C, H, Co, k, s, p in RESNET18_CONVS:
            x = torch.randn(batch, C, H, H, generator=g).to(dev)
            xs[name] = FixedPrecisionTensor.fix_precision(x, self.base, self.pf).share(*self.parties,
                                                                                       crypto_provider=self.provider)
        f = torch.randn(batch, 512, generator=g).to(dev)
        xs["fc"] = FixedPrecisionTensor.fix_precision(f, self.base, self.pf).share(*self.parties, crypto_provider=self.provider)
        return xs

    def preprocess(self, batch=1, n_images=1):
        """offline phase: build and distribute the triples n_images forward passes will consume."""
        for _ in range(n_images):
            for _name, shapes in triple_shapes(batch, self.ncls):
                self.provider.provide_primitives("matmul", shapes, self.parties, 1)

    def prepare_weight_side(self, triples, batch=1):
        """image-independent half of every layer from ``triples`` (per layer, per party), see EncryptedResNet18"""
        ws = {}
        for (name, C, H, Co, k, s, p), tri in zip(RESNET18_CONVS, triples):
            Ho = (H + 2 * p - k) // s + 1
            w = F.prepare_weight_side(self.weights[name], tri, batch, Ho, Ho)
            if w is not None:
                ws[name] = w
        return ws

    wside = {}

    def forward(self, xs):
        """online phase of all 21 linear layers; returns {name: FixedPrecisionTensor}."""
        out = {}
        for name, C, H, Co, k, s, p in RESNET18_CONVS:
            out[name] = F.conv2d(xs[name], self.weights[name], None, s, p, prepared=self.wside.get(name))
        out["fc"] = F.linear(xs["fc"], self.fc_w, self.fc_b)
        return out


class EncryptedLinearGraph:
    """The online phase of ``SharedLinearLayers.forward`` captured once in a CUDA graph (the ~214 launches of one encrypted
    image are latency-bound at batch 1).  The Beaver triples live in static buffers: the offline phase generates fresh
    triples (crypto provider, Philox + ring GEMM) and copies them into those buffers, the online phase is one graph replay.
    The crypto-store bookkeeping (peek in spdz_mask / pop in spdz_compute, primitives.py:52-102) runs at capture time."""

    def __init__(self, net: SharedLinearLayers, xs, batch=1):
        self.net, self.xs, self.batch = net, xs, batch
        net.preprocess(batch, 1)
        # the triples the capture is going to consume, in consumption order, per party
        self.static = []
        for _name, shapes in triple_shapes(batch, net.ncls):
            self.static.append([p.crypto_store.get_keys(op="matmul", shapes=shapes, remove=False) for p in net.parties])
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            net.forward(xs)  # warm-up: loads kernels, sizes the caching allocator
        torch.cuda.current_stream().wait_stream(side)
        self._refill_store()
        self.wside = net.prepare_weight_side(self.static, batch) if HOIST_WEIGHT_SIDE else {}
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from .. import _lib

        net.wside = self.wside
        l0 = _lib.launch_counter
        try:
            with torch.cuda.graph(self.graph):
                self.out = net.forward(xs)
        finally:
            net.wside = {}
        self.launches = _lib.launch_counter - l0     # kernels (and memsets) of one online replay

    def _refill_store(self):
        for (_name, shapes), per_party in zip(triple_shapes(self.batch, self.net.ncls), self.static):
            for p, tri in zip(self.net.parties, per_party):
                p.crypto_store.add_primitives("matmul", shapes, [tri])

    def offline(self):
        """fresh triples for the next image, written into the static buffers"""
        prov, parties = self.net.provider, self.net.parties
        for (_name, shapes), per_party in zip(triple_shapes(self.batch, self.net.ncls), self.static):
            fresh = prov.build_triple("matmul", shapes)
            for j in range(len(parties)):
                for dst, src in zip(per_party[j], fresh[j]):
                    dst.copy_(src, non_blocking=True)
        if self.wside:    # ... and the image-independent half of every layer, into the planes the graph reads
            fresh = self.net.prepare_weight_side(self.static, self.batch)
            for name, st in self.wside.items():
                for attr in F.WeightSide.TENSORS:
                    for dst, src in zip(getattr(st, attr), getattr(fresh[name], attr)):
                        dst.copy_(src, non_blocking=True)

    def online(self):
        self.graph.replay()
        return self.out


class EncryptedResNet18:
    """The reference's encrypted-inference model (inference.py:279-321): every parameter AND buffer of the ResNet-18
    fix-precision-encoded and secret-shared between model_owner and data_owner (hook.py:738-765 iterates parameters and
    buffers), evaluated layer by layer on shares.  ``forward`` is ResNet._forward_impl (torchlib/models.py:466-482) with
    ``model.pool`` and ``model.relu`` swapped as inference.py:289 does (stem: conv -> bn -> max-pool -> relu).

    ``shared`` maps state_dict keys to FixedPrecisionTensor > AdditiveSharingTensor; build it with ``from_state_dict``
    (encode + share on the GPU) or pass explicit shares (parity tests)."""

    def __init__(self, shared, parties, provider, base=10, precision_fractional=16, input_size=224, rng=None):
        self.P, self.parties, self.provider = shared, parties, provider
        self.base, self.pf, self.input_size = base, precision_fractional, input_size
        self.taps = None
        self.rng = rng
        # image-independent halves prepared by the offline phase (prepare_offline_side); empty = every layer runs the whole protocol
        self.wside = {}          # conv name -> functional.WeightSide
        self.bnside = {}         # BatchNorm name -> functional.BNSide
        self.hoisted_inv = None  # {BatchNorm name: inverse standard deviation} when the Newton iterations ran offline

    @classmethod
    def from_state_dict(cls, state_dict, parties, provider, base=10, precision_fractional=16, input_size=224, rng=None):
        """model.fix_precision(**kw).share(*workers, **kw) -- inference.py:280-286"""
        from .tensors import ShareRNG

        dev = parties[0].device
        rng = rng or ShareRNG(seed=0xA11CE ^ getattr(provider, "seed", 0))
        shared = {}
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue  # an integer counter: never read by the eval forward
            t = v.detach().to(dev, torch.float32).contiguous()
            shared[k] = FixedPrecisionTensor.fix_precision(t, base, precision_fractional).share(
                *parties, crypto_provider=provider, rng=rng)
        return cls(shared, parties, provider, base, precision_fractional, input_size, rng)

    def share_input(self, x: torch.Tensor, rng=None):
        """data.fix_precision(**kw).share(*workers, **kw) -- inference.py:307-311"""
        x = x.to(self.parties[0].device, torch.float32).contiguous()
        return FixedPrecisionTensor.fix_precision(x, self.base, self.pf).share(*self.parties, crypto_provider=self.provider,
                                                                               rng=rng or self.rng)

    def _tap(self, name, x):
        if self.taps is not None:
            self.taps[name] = [s.clone() for s in x.child.child]
        return x

    BN_ORDER = ["bn1"] + [f"layer{li}.{bi}.{n}" for li in range(1, 5) for bi in range(2)
                          for n in (("bn1", "bn2", "downsample.1") if (bi == 0 and li > 1) else ("bn1", "bn2"))]

    def _newton_all(self):
        """inverse standard deviations of all 20 BatchNorm layers (functional.py:62-64 -> precision.py:507-518) in one launch,
        consuming triples and constant sharings in forward order (one kernel for co-resident parties, one kernel per party
        exchanging openings over NVLink when they sit on two GPUs)."""
        from . import tensors as T

        if not T.FUSE_NEWTON:
            return {}
        inv = T.reciprocal_newton_batched([self.P[n + ".running_var"] for n in self.BN_ORDER])
        return dict(zip(self.BN_ORDER, inv))

    def _conv(self, x, name, stride, pad):
        return F.conv2d(x, self.P[name + ".weight"], None, stride, pad, prepared=self.wside.get(name))

    def prepare_weight_side(self, batch: int = 1):
        """Offline half of the 20 Beaver convolutions: for every layer, in forward order, peek the triple it is going to consume
        and run everything that does not involve the image -- mask + open the weights, build the limb planes of b (+ eps), eps
        and a (functional.prepare_weight_side).  The online forward of a layer is then mask -> open+planarise -> GEMM -> truncate.
        Needs the stores filled for exactly one forward (``preprocess(1)``)."""
        from .spdz import _key

        self.wside = {}
        seen = {}
        for name, C, Hin, Co, k, s, pd in conv_geometry(self.input_size):
            Ho = (Hin + 2 * pd - k) // s + 1
            shapes = ((batch, Ho * Ho, C * k * k), (C * k * k, Co))
            i = seen.get(_key(shapes), 0)
            seen[_key(shapes)] = i + 1
            tri = [p.crypto_store._stacks["matmul"][_key(shapes)][i] for p in self.parties]
            ws = F.prepare_weight_side(self.P[name + ".weight"], tri, batch, Ho, Ho)
            if ws is not None:
                self.wside[name] = ws

    NEWTON_ITERS = 80

    def prepare_offline_side(self, batch: int = 1):
        """Everything of one forward that depends on the model and the primitives but not on the image, run once per image in
        the OFFLINE phase: the 80-step Newton inverse square roots of all BatchNorm layers, the weight half of the 20 Beaver
        convolutions (prepare_weight_side) and the model-only operands of the 40 BatchNorm products (prepare_bn_side).  The stores
        must hold the primitives of exactly one forward.  Nothing is consumed: the online forward still pops every primitive (and
        advances the share RNG) where the op-by-op protocol would, so the two stay interchangeable primitive for primitive."""
        from . import tensors as T

        pdev = self.provider.provider.device
        for p in self.parties:   # the primitives were copied from the provider's GPU on ITS stream
            if p.device != pdev:
                torch.cuda.current_stream(p.device).wait_stream(torch.cuda.current_stream(pdev))
        self.hoisted_inv, self.bnside = None, {}
        self.wside = {}
        if "newton" in HOIST_PARTS and T.FUSE_NEWTON and self.rng is not None and self.rng.static:
            snaps = [p.crypto_store.export_state() for p in self.parties]
            mode, cur = self.rng.mode, self.rng.cursor
            self.rng.mode, self.rng.cursor = "replay", 0       # the constants' sharings are the first entries of a forward
            inv = self._newton_all()
            self.rng.mode, self.rng.cursor = mode, cur
            for p, st in zip(self.parties, snaps):
                p.crypto_store.import_state(st)
            self.hoisted_inv = inv
        if "conv" in HOIST_PARTS:
            self.prepare_weight_side(batch)
        if self.hoisted_inv is not None and "bn" in HOIST_PARTS:
            self.prepare_bn_side(batch)

    def prepare_bn_side(self, batch: int = 1):
        from .spdz import _key

        geo = {name: (Co, (Hin + 2 * pd - k) // s + 1) for name, _C, Hin, Co, k, s, pd in conv_geometry(self.input_size)}
        seen = {}
        for bn in self.BN_ORDER:
            conv = "conv1" if bn == "bn1" else bn.replace("downsample.1", "downsample.0").replace(".bn", ".conv")
            C, H = geo[conv]
            P = batch * H * H
            tris = []
            for key in (((C,), (P, C)), ((P, C), (C,))):
                i = seen.get(_key(key), 0)
                seen[_key(key)] = i + 1
                tris.append([p.crypto_store._stacks["mul"][_key(key)][i] for p in self.parties])
            self.bnside[bn] = F.prepare_bn_side(self.hoisted_inv[bn], self.P[bn + ".weight"], tris[0], tris[1], batch, C, H, H)

    def _newton_bookkeeping(self):
        """the pops and share-RNG draws of ``_newton_all`` without its arithmetic (the result came from the offline phase)"""
        from .spdz import take_primitives

        assert self.rng.mode == "replay"
        for bn in self.BN_ORDER:
            ast = self.P[bn + ".running_var"].child
            n = ast.child[0].shape[0]
            take_primitives("mul", ((n,), (n,)), 3 * (self.NEWTON_ITERS - 1), ast.parties, None)
            self.rng.cursor += 1

    def _bn(self, x, name):
        P = self.P
        if name in self.bnside:
            return F.batch_norm_prepared(x, P[name + ".running_mean"], P[name + ".bias"], self.bnside[name])
        return F.batch_norm(x, P[name + ".running_mean"], P[name + ".running_var"], P[name + ".weight"], P[name + ".bias"],
                            inv_std=self._inv.get(name))

    def forward(self, x: FixedPrecisionTensor) -> FixedPrecisionTensor:
        P = self.P
        if self.hoisted_inv is not None:
            self._newton_bookkeeping()
            self._inv = self.hoisted_inv
        else:
            self._inv = self._newton_all()
        x = self._tap("conv1", self._conv(x, "conv1", 2, 3))
        x = self._tap("bn1", self._bn(x, "bn1"))
        x = self._tap("pool", F.max_pool2d(x, 3, 2, 1))       # model.relu <- model.pool (inference.py:289)
        x = self._tap("relu", F.relu(x))                       # model.pool <- model.relu
        inplanes = 64
        for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
            for bi in range(2):
                st = stride if bi == 0 else 1
                pre = f"layer{li}.{bi}"
                identity = x
                out = self._conv(x, pre + ".conv1", st, 1)
                out = F.relu(self._bn(out, pre + ".bn1"))
                out = self._conv(out, pre + ".conv2", 1, 1)
                out = self._bn(out, pre + ".bn2")
                if st != 1 or inplanes != planes:
                    identity = self._bn(self._conv(x, pre + ".downsample.0", st, 0), pre + ".downsample.1")
                out = out + identity
                x = self._tap(pre, F.relu(out))
                inplanes = planes
        x = F.avg_pool2d(x, self.input_size // 32)
        x = x._new(x.child.reshape(x.shape[0], -1))            # torch.flatten(x, 1)
        return F.linear(x, P["fc.weight"], P["fc.bias"])

    __call__ = forward

    # ---- offline / online split (triple and key generation is reported separately)
    def trace(self, x: FixedPrecisionTensor):
        """run one forward with on-demand primitives and remember which primitives it asked for, in order"""
        start = len(self.provider.request_log)
        out = self.forward(x)
        self.schedule = list(self.provider.request_log[start:])
        return out

    def preprocess(self, n_images: int = 1):
        """offline phase: the crypto provider generates and distributes every Beaver triple and FSS key that ``n_images``
        forward passes will consume (beaver.py:7-63, primitives.py:237-286), in consumption order."""
        assert getattr(self, "schedule", None), "call trace(x) once first: it records the primitive schedule"
        for _ in range(n_images):
            for op, shapes, n in self.schedule:
                self.provider.provide_primitives(op, shapes, self.parties, n)

    def predict(self, x: torch.Tensor):
        """one pass of the loop body of inference.py:292-317: share the image, forward, reconstruct, decode, argmax"""
        out = self.forward(self.share_input(x)).get().float_prec()
        return out, out.argmax(dim=1)


class _MultiDeviceCapture:
    """Capture ``fn()`` into ONE CUDA graph that spans several GPUs.  The capture runs on ``devices[0]``; every other device
    gets (a) a private memory pool that lives as long as the graph (replays reuse the addresses) and (b) a capturable side
    stream, forked from the capture stream by an event and made that device's current stream for the duration, joined again
    before the capture ends.  Cross-device edges inside ``fn`` are ordinary event waits between the devices' current streams."""

    def __init__(self, devices):
        self.devices = list(devices)
        self.graph = torch.cuda.CUDAGraph()
        # a MemPool registers its reference with the allocator of the device that is CURRENT when it is constructed: built under
        # another device, the pool's use count on `d` drops to zero when the capture ends and the next empty_cache() -- every
        # torch.cuda.graph capture starts with one -- cudaFree()s the segments that hold only capture-time temporaries, which the
        # graph still writes on replay (found as an illegal address in the second replay of the 3-GPU offline graph)
        self.pools = {}
        for d in self.devices[1:]:
            with torch.cuda.device(d):
                self.pools[d] = torch.cuda.MemPool()
        self.streams = {d: torch.cuda.Stream(d) for d in self.devices[1:]}
        # torch.cuda.graph's default capture stream is a process-wide singleton living on whichever device captured first:
        # a capture on another GPU must bring its own stream or nothing is recorded
        self.capture_stream = torch.cuda.Stream(self.devices[0])

    def capture(self, fn):
        import contextlib

        dev, others = self.devices[0], self.devices[1:]
        with contextlib.ExitStack() as es:
            for d in others:
                es.enter_context(torch.cuda.use_mem_pool(self.pools[d], device=d))
            with torch.cuda.device(dev), torch.cuda.graph(self.graph, stream=self.capture_stream,
                                                          capture_error_mode="relaxed" if others else "global"):
                main = torch.cuda.current_stream(dev)
                prev = {}
                for d in others:
                    prev[d] = torch.cuda.current_stream(d)
                    self.streams[d].wait_stream(main)          # fork: the stream joins the capture
                    with torch.cuda.device(d):
                        torch.cuda.set_stream(self.streams[d])
                try:
                    out = fn()
                finally:
                    for d in others:
                        main.wait_stream(self.streams[d])      # join
                        with torch.cuda.device(d):
                            torch.cuda.set_stream(prev[d])
        return out

    def replay(self):
        dev = self.devices[0]
        for d in self.devices[1:]:                              # the replay (launched on dev) must see the other GPUs' pending work
            torch.cuda.current_stream(dev).wait_stream(torch.cuda.current_stream(d))
        with torch.cuda.device(dev):
            self.graph.replay()
        for d in self.devices[1:]:                              # ... and their later work must see the replay's results
            torch.cuda.current_stream(d).wait_stream(torch.cuda.current_stream(dev))


class EncryptedInferenceGraph:
    """One encrypted image as TWO CUDA-graph replays:

      offline()    the crypto provider regenerates every primitive the forward consumes -- 8.85 GB of Beaver triples, FSS keys
                   and sharings of the Newton constant -- IN PLACE: the generating launches (Philox, ring GEMM c = a@b, DIF
                   keygen) were captured once, their output tensors are the graph's own allocations, and a device-side epoch
                   added to every Philox offset (pm_epoch_bump is the graph's first node) makes each replay draw fresh
                   randomness.  No host bookkeeping, no staging copy.  The same graph then runs everything of the forward that
                   depends on the model and the primitives but not on the image (EncryptedResNet18.prepare_offline_side):
                   the Newton inverse square roots, the weight half of every Beaver convolution, the model-only operands of
                   the BatchNorm products.
      online(img)  share input -> forward on shares -> reconstruct -> decode: 554 launches (1497 without the hoisting and the
                   fused openings) of mostly tiny kernels that are launch-bound when issued eagerly from Python.  The
                   crypto-store bookkeeping (peek / pop, primitives.py:52-102) ran on the host at capture time, primitive for
                   primitive as the op-by-op protocol would.

    Placement: with the two share holders on different GPUs (model_owner cuda:0, data_owner cuda:1,
    crypto provider cuda:2) both graphs are multi-device graphs: every opening is a kernel on the consuming GPU that loads the
    peer's share through its peer-mapped pointer behind a cross-stream event edge, the provider's primitives travel to the
    parties as peer copies, and each party's FSS evaluation -- the dominant cost -- runs on its own GPU concurrently with the
    other's."""

    def __init__(self, net: EncryptedResNet18, example: torch.Tensor):
        self.net = net
        dev = net.parties[0].device
        others = []
        for p in net.parties[1:]:
            if p.device != dev and p.device not in others:
                others.append(p.device)
        self.devices = [dev] + others
        pdev = net.provider.provider.device
        self.off_devices = [pdev] + [d for d in self.devices if d != pdev]
        assert net.rng is not None, "build the model with from_state_dict (it owns the share RNG)"
        x = net.share_input(example)
        net.trace(x)                                   # warm-up + primitive schedule (primitives on demand, consumed)
        self.x_static = x
        # a recording pass: which constants get freshly shared during a forward (additive_shared.py:473-487), into static buffers
        net.preprocess(1)
        self._sync()
        net.rng.mode, net.rng.static = "record", []
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            net.forward(x)
        torch.cuda.current_stream(dev).wait_stream(side)
        self._sync()
        net.rng.mode = "live"
        for p in net.parties:
            p.crypto_store.clear()
        # ---- offline graph: generation straight into the tensors the online graph will read
        b0 = net.provider.generated_bytes
        self._off = _MultiDeviceCapture(self.off_devices)

        def generate():
            net.provider.bump_epoch()      # graph nodes: every replay draws from a fresh Philox stream
            net.rng.bump_epoch()
            net.preprocess(1)
            net.rng.refresh_static()
            if HOIST_WEIGHT_SIDE:
                net.prepare_offline_side(example.shape[0])   # Newton, the weight half of every Beaver convolution, BatchNorm operands

        self._off.capture(generate)
        self.bytes_per_image = net.provider.generated_bytes - b0
        self.static_state = [p.crypto_store.export_state() for p in net.parties]
        self._off.replay()                             # the capture itself executed nothing: fill the buffers once
        self._sync()
        # ---- online graph
        net.rng.mode, net.rng.cursor = "replay", 0
        from .. import _lib

        l0 = _lib.launch_counter
        self._on = _MultiDeviceCapture(self.devices)

        def forward():
            self.out_shares = net.forward(x)
            self.logits = self.out_shares.get().float_prec()

        self._on.capture(forward)
        self.graph = self._on.graph
        self.kernels_in_graph = _lib.launch_counter - l0
        net.rng.mode = "live"
        # the hoisted operands belong to the graphs; an eager net.forward runs the whole protocol op by op
        self.wside, self.bnside, self.hoisted_inv = net.wside, net.bnside, net.hoisted_inv
        net.wside, net.bnside, net.hoisted_inv = {}, {}, None
        for p in net.parties:
            p.crypto_store.clear()

    def _sync(self):
        for d in set(self.devices) | set(self.off_devices):
            torch.cuda.synchronize(d)

    def offline(self):
        """fresh triples, FSS keys and constant sharings for the next image: one replay of the generation graph"""
        self._off.replay()
        for d in self.devices:                          # the online replay (launched on devices[0]) must see the new primitives
            if d != self.off_devices[0]:
                torch.cuda.current_stream(d).wait_stream(torch.cuda.current_stream(self.off_devices[0]))
        self.net.provider.generated_bytes += self.bytes_per_image

    def online(self, img: torch.Tensor):
        """inference.py:307-317 for one image: returns (logits, argmax) -- device tensors owned by the graph"""
        s = self.net.share_input(img)
        for dst, src in zip(self.x_static.child.child, s.child.child):
            dst.copy_(src, non_blocking=True)
        self._on.replay()
        return self.logits, self.logits.argmax(dim=1)
