"""ResNet-18 (torchlib/models.py:345-516) forward / backward / optimizer step as a fixed schedule of
primia_b200 kernels.

The engine owns ONE flat fp32 buffer ``flat`` = [all parameters in ``model.parameters()`` order |
all BN running_mean/running_var], so that
  * the optimizer is a single kernel over ``flat[:n_params]``            (train.py:280-303),
  * FedAvg is a single all-reduce over ``flat``                          (torchlib/utils.py:1000-1092;
    ``num_batches_tracked`` is skipped exactly as the reference does, utils.py:1040).
Activations are NHWC, conv weights KRSC; ``state_dict()/load_state_dict()`` speak the reference's
torch layout (NCHW/KCRS, torchvision key names) so checkpoints interchange (utils.py:1470-1493).

mode "f32": fp32 activations, FFMA implicit-GEMM convs (the 1e-5 parity gate runs here).
mode "bf16": bf16 activations, tcgen05 tensor-core convs with fp32 accumulation (throughput mode).
There is no PyTorch fallback: every arithmetic op is a C-ABI call.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict

import torch

from .._lib import ConvDesc, PrimiaError, call, ptr, stream


def _align(n, a=4):
    return (n + a - 1) // a * a


class _Conv:
    def __init__(self, name, B, H, W, C, K, R, stride, pad):
        self.name = name
        self.Ho = (H + 2 * pad - R) // stride + 1
        self.Wo = (W + 2 * pad - R) // stride + 1
        self.desc = ConvDesc(B, H, W, C, K, R, R, stride, pad, self.Ho, self.Wo)
        self.B, self.H, self.W, self.C, self.K, self.R, self.stride, self.pad = B, H, W, C, K, R, stride, pad
        self.P = B * self.Ho * self.Wo  # output rows

    @property
    def wshape(self):  # KRSC
        return (self.K, self.R, self.R, self.C)


class ResNet18Engine:
    BN_EPS = 1e-5
    BN_MOMENTUM = 0.1
    SPLIT_BLOCK = 6  # blocks[6:] = layer4: 75 % of the parameters, and the FIRST gradients the backward pass finishes

    def __init__(self, batch, num_classes=3, in_channels=3, input_size=224, pooling="max", device="cuda:0", mode="f32",
                 optimizer="Adam", lr=1e-4, betas=(0.5, 0.99), weight_decay=5e-4, eps=1e-8, class_weights=None, adptpool=False):
        if pooling not in ("max", "avg"):
            raise NotImplementedError("pooling type unknown: {:s}".format(str(pooling)))  # models.py:388-389
        if mode not in ("f32", "bf16"):
            raise ValueError(mode)
        self.B, self.ncls, self.cin, self.size = batch, num_classes, in_channels, input_size
        self.pooling, self.adptpool = pooling, adptpool
        self.device = torch.device(device)
        self.mode = mode
        self.adt = torch.float32 if mode == "f32" else torch.bfloat16
        self.sfx = "_f32" if mode == "f32" else "_bf16"
        self.opt_name, self.lr, self.betas, self.wd, self.opt_eps = optimizer, lr, betas, weight_decay, eps
        self.step_count = 0
        self.training = True
        self._prof = None
        self._graph = None
        self.fuse_stats = True  # BN batch statistics accumulated in the conv epilogue (bf16 mode)
        self.fuse_bn_bwd = True  # BN backward as one launch with a grid barrier
        self.fuse_head = os.environ.get("PRIMIA_FUSE_HEAD", "1") != "0"    # pool + Linear + CE + their gradients as one launch per step (batch reductions on the side stream)
        self.overlap_wgrad = mode == "bf16"
        # bf16 throughput mode: the stem's BN + ReLU + max-pool run as one pass (the 112x112 activation is never written) and
        # BN backward recomputes the ReLU decision from x instead of reading the stored activation where no residual is added
        self.fuse_stem_pool = mode == "bf16" and pooling == "max"
        self.bn_xmask = mode == "bf16"
        # bf16 stem as a direct implicit GEMM over the fp32 NCHW input (conv_stem.cu): no im2col matrix.  The im2col + dense
        # GEMM route stays for other channel counts and as the cross-check (PRIMIA_NO_DIRECT_STEM=1).
        self.direct_stem = (mode == "bf16" and in_channels == 3 and input_size % 4 == 0
                            and os.environ.get("PRIMIA_NO_DIRECT_STEM", "0") != "1")
        self._x_in = None
        self._side = None
        self._split_cb = None
        self._graph2 = None
        self.on_inputs_staged = None
        self._mv_zero = True
        self._build_graph()
        self._alloc()
        self.class_weights = None
        if class_weights is not None:
            self.class_weights = torch.as_tensor(class_weights, dtype=torch.float32, device=self.device).contiguous()

    # ------------------------------------------------------------------ graph
    def _build_graph(self):
        B, S = self.B, self.size
        convs, bns = OrderedDict(), OrderedDict()
        c1 = _Conv("conv1", B, S, S, self.cin, 64, 7, 2, 3)
        convs["conv1"] = c1
        # bf16 mode runs the stem as a dense 1x1 problem over a bf16 im2col tensor [B,Ho,Wo,KPAD] (k = (r*7+s)*Cin + c)
        self.stem_kpad = (7 * 7 * self.cin + 63) // 64 * 64
        self.c1_gemm = _Conv("conv1", B, c1.Ho, c1.Wo, self.stem_kpad, 64, 1, 1, 0)
        bns["bn1"] = 64
        H = (c1.Ho + 2 - 3) // 2 + 1  # maxpool 3,2,1
        self.pool_in, self.pool_out = c1.Ho, H
        blocks = []
        inpl = 64
        for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
            for bi in range(2):
                st = stride if bi == 0 else 1
                pre = f"layer{li}.{bi}"
                ca = _Conv(pre + ".conv1", B, H, H, inpl, planes, 3, st, 1)
                cb = _Conv(pre + ".conv2", B, ca.Ho, ca.Ho, planes, planes, 3, 1, 1)
                convs[ca.name], convs[cb.name] = ca, cb
                bns[pre + ".bn1"], bns[pre + ".bn2"] = planes, planes
                ds = None
                if st != 1 or inpl != planes:
                    ds = _Conv(pre + ".downsample.0", B, H, H, inpl, planes, 1, st, 0)
                    convs[ds.name] = ds
                    bns[pre + ".downsample.1"] = planes
                blocks.append((pre, ca, cb, ds))
                inpl, H = planes, ca.Ho
        self.convs, self.bns, self.blocks = convs, bns, blocks
        self.final_hw = H
        if not self.adptpool and int(S / 32) != H:
            # AvgPool2d(int(input_size / 32)) (models.py:400-404) would leave more than one pixel: the reference's fc then fails
            raise ValueError(f"input_size {S}: final feature map {H}x{H} != AvgPool2d({int(S / 32)}); use adptpool=True")
        self.feat_dim = 512
        # parameter order == model.parameters() order of the reference module
        order = [("conv1.weight", c1.wshape), ("bn1.weight", (64,)), ("bn1.bias", (64,))]
        for pre, ca, cb, ds in blocks:
            order += [(ca.name + ".weight", ca.wshape), (pre + ".bn1.weight", (ca.K,)), (pre + ".bn1.bias", (ca.K,)),
                      (cb.name + ".weight", cb.wshape), (pre + ".bn2.weight", (cb.K,)), (pre + ".bn2.bias", (cb.K,))]
            if ds is not None:
                order += [(ds.name + ".weight", ds.wshape), (pre + ".downsample.1.weight", (ds.K,)),
                          (pre + ".downsample.1.bias", (ds.K,))]
        order += [("fc.weight", (self.ncls, 512)), ("fc.bias", (self.ncls,))]
        self.param_order = order
        off = 0
        self.offsets = OrderedDict()
        for name, shape in order:
            n = 1
            for d in shape:
                n *= d
            self.offsets[name] = (off, n, shape)
            off = _align(off + n)
        self.n_param_flat = off
        for bn, C in bns.items():
            for stat in ("running_mean", "running_var"):
                self.offsets[f"{bn}.{stat}"] = (off, C, (C,))
                off = _align(off + C)
        self.n_flat = off

    # ------------------------------------------------------------------ memory
    def _alloc(self):
        dev, f32 = self.device, torch.float32
        self.flat = torch.zeros(self.n_flat, dtype=f32, device=dev)
        self.grads = torch.zeros(self.n_param_flat, dtype=f32, device=dev)
        self.adam_m = torch.zeros(self.n_param_flat, dtype=f32, device=dev)
        self.adam_v = torch.zeros(self.n_param_flat, dtype=f32, device=dev)
        self.p = {k: self.flat[o:o + n].view(shape) for k, (o, n, shape) in self.offsets.items()}
        self.g = {k: self.grads[o:o + n].view(shape) for k, (o, n, shape) in self.offsets.items() if o < self.n_param_flat}
        for bn in self.bns:
            self.p[bn + ".running_var"].fill_(1.0)
            self.p[bn + ".weight"].fill_(1.0)
        B, adt = self.B, self.adt
        A = lambda *s: torch.empty(s, dtype=adt, device=dev)
        c1 = self.convs["conv1"]
        if self.mode == "f32":
            self.x0 = A(B, self.size, self.size, self.cin)
        else:
            if self.direct_stem:
                self.x0 = None
                self.w_stem = torch.zeros((64, 192), dtype=torch.bfloat16, device=dev)  # k = (r*3 + c)*8 + s
            else:
                self.x0 = A(B, c1.Ho, c1.Wo, self.stem_kpad)  # im2col of the input
            self.dw_stem = torch.zeros((64, self.stem_kpad), dtype=f32, device=dev)
        self.act = {}   # forward tensors
        self.grad = {}  # backward tensors
        self.act["conv1"] = A(B, c1.Ho, c1.Wo, 64)
        self.act["a1"] = A(B, c1.Ho, c1.Wo, 64)
        self.act["p1"] = A(B, self.pool_out, self.pool_out, 64)
        self.pool_idx = torch.empty((B, self.pool_out, self.pool_out, 64), dtype=torch.uint8, device=dev)
        # raw conv1 output at each pooling window's argmax (fused stem pool, bf16 mode): the stem's BN backward sums read it
        self.stem_xmax = torch.empty((B, self.pool_out, self.pool_out, 64), dtype=adt, device=dev) if self.mode == "bf16" else None
        for pre, ca, cb, ds in self.blocks:
            self.act[ca.name] = A(B, ca.Ho, ca.Wo, ca.K)
            self.act[pre + ".a"] = A(B, ca.Ho, ca.Wo, ca.K)
            self.act[cb.name] = A(B, cb.Ho, cb.Wo, cb.K)
            self.act[pre + ".out"] = A(B, cb.Ho, cb.Wo, cb.K)
            if ds is not None:
                self.act[ds.name] = A(B, ds.Ho, ds.Wo, ds.K)
                self.act[pre + ".idn"] = A(B, ds.Ho, ds.Wo, ds.K)
        self.bn_mean = {bn: torch.empty(C, dtype=f32, device=dev) for bn, C in self.bns.items()}
        self.bn_invstd = {bn: torch.empty(C, dtype=f32, device=dev) for bn, C in self.bns.items()}
        self.stats = torch.zeros(len(self.bns) * 2 * 2 * 512, dtype=torch.float64, device=dev)  # fwd + bwd slots
        from .._lib import lib as _l
        _l().pm_bn_bwd_fused_ws_doubles.restype = ctypes.c_size_t
        self.bn_bwd_ws = torch.zeros(int(_l().pm_bn_bwd_fused_ws_doubles(512)), dtype=torch.float64, device=dev)
        self.feat = torch.empty((B, 512), dtype=f32, device=dev)
        self.dfeat = torch.empty((B, 512), dtype=f32, device=dev)
        self.logits = torch.empty((B, self.ncls), dtype=f32, device=dev)
        self.loss = torch.zeros(1, dtype=f32, device=dev)
        self.head_ws = torch.empty(B * (self.ncls + 1) + 4, dtype=f32, device=dev)
        # gradient buffers: two ping-pong buffers per spatial stage + one for conv-output grads
        self._gbufs = {}
        ws_bytes = 0
        for c in self.convs.values():
            ws_bytes = max(ws_bytes, int(self._lib_size("pm_conv_wgrad_ws_bytes", c)))
        self.wgrad_ws = torch.empty(max(ws_bytes, 16) // 4 + 4, dtype=f32, device=dev)
        if self.mode == "bf16":
            from .._lib import WCvt

            self.wbf = {}
            entries = []
            for name, c in self.convs.items():
                if name == "conv1":
                    wf = torch.empty((64, 1, 1, self.stem_kpad), dtype=torch.bfloat16, device=dev)
                    self.wbf[name] = (wf, None)
                    entries.append(WCvt(self.p[name + ".weight"].data_ptr(), wf.data_ptr(), None, 64, 7 * 7 * self.cin, 1, self.stem_kpad))
                else:
                    wf = torch.empty(c.wshape, dtype=torch.bfloat16, device=dev)
                    wd = torch.empty((c.C, c.R, c.R, c.K), dtype=torch.bfloat16, device=dev)
                    self.wbf[name] = (wf, wd)
                    entries.append(WCvt(self.p[name + ".weight"].data_ptr(), wf.data_ptr(), wd.data_ptr(), c.K, c.C, c.R * c.R, c.C))
            arr = (WCvt * len(entries))(*entries)
            self.wcvt_n = len(entries)
            self.wcvt_tiles = max(e.RS * ((e.K + 31) // 32) * ((e.Cpad + 31) // 32) for e in entries)
            self.wcvt_total = sum(e.RS * ((e.K + 31) // 32) * ((e.Cpad + 31) // 32) for e in entries)
            self.wcvt_table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)

    def _lib_size(self, fn, conv):
        from .._lib import lib

        return getattr(lib(), fn)(ctypes.byref(conv.desc))

    def _gbuf(self, key, like):
        t = self._gbufs.get(key)
        if t is None or t.shape != like.shape:
            t = torch.empty_like(like)
            self._gbufs[key] = t
        return t

    # ------------------------------------------------------------------ state (reference layout)
    def load_state_dict(self, sd):
        """Accepts the reference's ``model.state_dict()`` (NCHW/KCRS fp32, torchvision names)."""
        with torch.cuda.device(self.device):
            for name, (o, n, shape) in self.offsets.items():
                if name not in sd:
                    raise KeyError(name)
                src = sd[name].detach().to(self.device, torch.float32).contiguous()
                if name.endswith("weight") and src.dim() == 4:
                    K, C, R, S_ = src.shape
                    call("pm_kcrs_to_krsc_f32", ptr(src), K, C, R, S_, ptr(self.p[name]), stream())
                else:
                    self.p[name].copy_(src.view(shape))

    def state_dict(self):
        out = OrderedDict()
        with torch.cuda.device(self.device):
            for name, (o, n, shape) in self.offsets.items():
                t = self.p[name]
                if len(shape) == 4:
                    K, R, S_, C = shape
                    dst = torch.empty((K, C, R, S_), dtype=torch.float32, device=self.device)
                    call("pm_krsc_to_kcrs_f32", ptr(t.contiguous()), K, C, R, S_, ptr(dst), stream())
                    out[name] = dst
                else:
                    out[name] = t.clone()
        # reference key order (bn: weight, bias, running_mean, running_var, num_batches_tracked)
        ordered = OrderedDict()
        for name, _ in self.param_order:
            ordered[name] = out[name]
            if name.endswith(".bias") and name[:-5] in self.bns:
                bn = name[:-5]
                ordered[bn + ".running_mean"] = out[bn + ".running_mean"]
                ordered[bn + ".running_var"] = out[bn + ".running_var"]
                ordered[bn + ".num_batches_tracked"] = torch.tensor(self.step_count, dtype=torch.int64)
        return ordered

    def flat_to_torch_layout(self, flat):
        """a parameter-shaped flat buffer (gradients, Adam moments) as {name: tensor} in the reference layout (KCRS)"""
        out = OrderedDict()
        with torch.cuda.device(self.device):
            for name, _ in self.param_order:
                o, n, shape = self.offsets[name]
                t = flat[o:o + n].view(shape)
                if t.dim() == 4:
                    K, R, S_, C = t.shape
                    dst = torch.empty((K, C, R, S_), dtype=torch.float32, device=self.device)
                    call("pm_krsc_to_kcrs_f32", ptr(t), K, C, R, S_, ptr(dst), stream())
                    out[name] = dst
                else:
                    out[name] = t.clone()
        return out

    def torch_layout_to_flat(self, tensors, flat):
        """inverse of ``flat_to_torch_layout``: {name: tensor in the reference layout} -> the flat buffer (in place)"""
        with torch.cuda.device(self.device):
            for name, _ in self.param_order:
                o, n, shape = self.offsets[name]
                src = tensors[name].detach().to(self.device, torch.float32).contiguous()
                dst = flat[o:o + n].view(shape)
                if src.dim() == 4:
                    K, C, R, S_ = src.shape
                    call("pm_kcrs_to_krsc_f32", ptr(src), K, C, R, S_, ptr(dst), stream())
                else:
                    dst.copy_(src.view(shape))

    def grad_dict(self):
        """gradients in the reference layout (KCRS) -- used by the parity tests"""
        return self.flat_to_torch_layout(self.grads)

    # ------------------------------------------------------------------ kernels
    def _prof_begin(self, tag="conv"):
        """Profiling pass only: CUDA-event bracket around ONE conv launch.  The launch takes the host longer (ctypes call, two
        tensor-map encodes) than the kernel takes the GPU, so a ~100 us spin kernel is queued first: the host runs ahead and the
        events bracket device time, not launch latency."""
        if self._prof is None:
            return None
        torch.cuda._sleep(200000)
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return (e, tag)

    def _prof_end(self, e0):
        if e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self._prof.append((e0[0], e1, e0[1]))

    def _conv_fwd(self, c, x, y, stats=None):
        e0 = self._prof_begin("fwd")
        self._conv_fwd_impl(c, x, y, stats)
        self._prof_end(e0)

    def _conv_dgrad(self, c, dy, dx, accumulate):
        e0 = self._prof_begin("dgrad")
        self._conv_dgrad_impl(c, dy, dx, accumulate)
        self._prof_end(e0)

    def _conv_wgrad(self, c, x, dy):
        """Weight gradients are off the backward critical path (nothing downstream reads them before the optimizer),
        so in bf16 mode they are issued on a side stream and overlap the BN / dgrad chain (the tensor-core wgrad CTAs
        co-reside with the memory-bound BN kernels).  Each conv has its own dy buffer, so there is no reuse hazard."""
        if self._side is None or self._prof is not None:
            e0 = self._prof_begin("wgrad")
            self._conv_wgrad_impl(c, x, dy)
            self._prof_end(e0)
            return
        ev = torch.cuda.Event()
        ev.record()
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            self._conv_wgrad_impl(c, x, dy)

    def _conv_fwd_impl(self, c, x, y, stats=None):
        if self.mode == "f32":
            call("pm_conv_fwd_f32", ctypes.byref(c.desc), ptr(x), ptr(self.p[c.name + ".weight"]), ptr(y), stream())
        else:
            call("pm_conv_fwd_bf16", ctypes.byref(c.desc), ptr(x), ptr(self.wbf[c.name][0]), ptr(y),
                 ptr(stats) if stats is not None else None, stream())

    def _conv_dgrad_impl(self, c, dy, dx, accumulate):
        if self.mode == "f32":
            call("pm_conv_dgrad_f32", ctypes.byref(c.desc), ptr(dy), ptr(self.p[c.name + ".weight"]), ptr(dx),
                 int(accumulate), stream())
        else:
            call("pm_conv_dgrad_bf16", ctypes.byref(c.desc), ptr(dy), ptr(self.wbf[c.name][1]), ptr(dx), int(accumulate),
                 stream())

    def _conv_wgrad_impl(self, c, x, dy):
        fn = "pm_conv_wgrad_f32" if self.mode == "f32" else "pm_conv_wgrad_bf16"
        call(fn, ctypes.byref(c.desc), ptr(x), ptr(dy), ptr(self.g[c.name + ".weight"]), ptr(self.wgrad_ws), stream())

    def _stat_slot(self, idx):
        return self.stats[idx * 1024:(idx + 1) * 1024]

    def _bn_fwd(self, bn, idx, x, y, residual, relu, P, stats_done=False):
        e0 = self._prof_begin("bn_fwd")
        self._bn_fwd_impl(bn, idx, x, y, residual, relu, P, stats_done)
        self._prof_end(e0)

    def _bn_bwd(self, bn, idx, dy, y_out, x, dx, P, g_out=None):
        e0 = self._prof_begin("bn_bwd")
        self._bn_bwd_impl(bn, idx, dy, y_out, x, dx, P, g_out)
        self._prof_end(e0)

    def _bn_fwd_impl(self, bn, idx, x, y, residual, relu, P, stats_done=False):
        C = self.bns[bn]
        st = self._stat_slot(idx)
        if self.training:
            if not stats_done:
                call("pm_bn_stats" + self.sfx, ptr(x), P, C, ptr(st), stream())
            call("pm_bn_fwd_fused" + self.sfx, ptr(x), ptr(st), P, C, ctypes.c_float(self.BN_EPS), ctypes.c_float(self.BN_MOMENTUM),
                 ptr(self.p[bn + ".weight"]), ptr(self.p[bn + ".bias"]), ptr(residual) if residual is not None else None,
                 int(relu), ptr(y), ptr(self.bn_mean[bn]), ptr(self.bn_invstd[bn]), ptr(self.p[bn + ".running_mean"]),
                 ptr(self.p[bn + ".running_var"]), stream())
            return
        else:
            # eval: normalise with the running statistics (F.batch_norm(training=False))
            self.bn_mean[bn].copy_(self.p[bn + ".running_mean"])
            torch.rsqrt(self.p[bn + ".running_var"] + self.BN_EPS, out=self.bn_invstd[bn])
        call("pm_bn_apply" + self.sfx, ptr(x), ptr(self.bn_mean[bn]), ptr(self.bn_invstd[bn]), ptr(self.p[bn + ".weight"]),
             ptr(self.p[bn + ".bias"]), ptr(residual) if residual is not None else None, int(relu), P, C, ptr(y), stream())

    def _bn_bwd_impl(self, bn, idx, dy, y_out, x, dx, P, g_out=None):
        C = self.bns[bn]
        if self.fuse_bn_bwd and self.bn_xmask and self.mode == "bf16" and y_out is not None and g_out is None and self.training:
            # BN -> ReLU with no residual: the mask is recomputed from x (two fewer passes over an activation-sized tensor)
            call("pm_bn_bwd_fused_xmask_bf16", ptr(dy), ptr(x), ptr(self.bn_mean[bn]), ptr(self.bn_invstd[bn]),
                 ptr(self.p[bn + ".weight"]), ptr(self.p[bn + ".bias"]), P, C, ptr(self.bn_bwd_ws), ptr(dx),
                 ptr(self.g[bn + ".weight"]), ptr(self.g[bn + ".bias"]), stream())
            return
        if self.fuse_bn_bwd:
            call("pm_bn_bwd_fused" + self.sfx, ptr(dy), ptr(y_out) if y_out is not None else None, ptr(x), ptr(self.bn_mean[bn]),
                 ptr(self.bn_invstd[bn]), ptr(self.p[bn + ".weight"]), P, C, ptr(self.bn_bwd_ws), ptr(g_out) if g_out is not None else None,
                 ptr(dx), ptr(self.g[bn + ".weight"]), ptr(self.g[bn + ".bias"]), stream())
            return
        st = self._stat_slot(len(self.bns) + idx)
        call("pm_bn_bwd_reduce" + self.sfx, ptr(dy), ptr(y_out) if y_out is not None else None, ptr(x),
             ptr(self.bn_mean[bn]), ptr(self.bn_invstd[bn]), P, C, ptr(st), ptr(g_out) if g_out is not None else None, stream())
        src = g_out if g_out is not None else dy
        call("pm_bn_bwd_apply" + self.sfx, ptr(src), None if g_out is not None else (ptr(y_out) if y_out is not None else None),
             ptr(x), ptr(self.bn_mean[bn]), ptr(self.bn_invstd[bn]), ptr(self.p[bn + ".weight"]), ptr(st), P, C, ptr(dx),
             ptr(self.g[bn + ".weight"]), ptr(self.g[bn + ".bias"]), stream())

    # ------------------------------------------------------------------ forward / backward
    def set_input(self, x_nchw):
        """x: [B,C,H,W] fp32 CUDA (the reference's loader layout) -> internal NHWC."""
        if tuple(x_nchw.shape) != (self.B, self.cin, self.size, self.size):
            raise PrimiaError(f"expected input {(self.B, self.cin, self.size, self.size)}, got {tuple(x_nchw.shape)}")
        x_nchw = x_nchw.contiguous()
        if x_nchw.dtype == torch.bfloat16:
            # a loader may stage the batch as bf16 (half the PCIe bytes): in bf16 mode the stem rounds every pixel to bf16 anyway,
            # so the step is bit-identical to shipping fp32; the expansion is a device-side copy
            x_nchw = x_nchw.float()
        if self.mode == "f32":
            call("pm_nchw_to_nhwc_f32", ptr(x_nchw), self.B, self.cin, self.size, self.size, ptr(self.x0), stream())
        elif self.direct_stem:
            if x_nchw.dtype != torch.float32:
                raise PrimiaError("the stem kernel reads the fp32 NCHW batch the reference's loader produces")
            self._x_in = x_nchw  # read in place by the stem forward and (later) its weight gradient
        else:
            call("pm_im2col_stem_bf16", ptr(x_nchw), self.B, self.cin, self.size, self.size, 7, 2, 3, self.stem_kpad, ptr(self.x0),
                 stream())

    def refresh_bf16_weights(self):
        """bf16 copies of the fp32 master weights for this step.  The stem's operand first (its conv is the first kernel); the
        44.7 MB batched cast of every other layer -- and the clearing of the flat gradient buffer the split-K weight gradients
        accumulate into -- run on the side stream underneath the stem conv / BN / pool and join before layer1."""
        if self.direct_stem:
            call("pm_stem_prep_w_bf16", ptr(self.p["conv1.weight"]), ptr(self.w_stem), stream())
        side = self._get_side() if (self.overlap_wgrad and self._prof is None and self.direct_stem) else None
        if side is None:
            call("pm_krsc_to_bf16_batched_exact", ptr(self.wcvt_table), self.wcvt_n, self.wcvt_total, stream())
            self._cast_done = None
            return
        ev = torch.cuda.Event()
        ev.record()
        side.wait_event(ev)
        with torch.cuda.stream(side):
            call("pm_krsc_to_bf16_batched_exact", ptr(self.wcvt_table), self.wcvt_n, self.wcvt_total, stream())
            if self.training:
                self.grads.zero_()
                self._grads_cleared = True
            self._cast_done = torch.cuda.Event()
            self._cast_done.record()

    def _get_side(self):
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        return self._side

    def forward(self, x_nchw=None):
        with torch.cuda.device(self.device):
            if x_nchw is not None:
                self.set_input(x_nchw)
            if self.training:
                self.stats.zero_()
            if self.mode == "bf16":
                e0 = self._prof_begin("weight_cast")
                self.refresh_bf16_weights()
                self._prof_end(e0)
            bn_ids = {bn: i for i, bn in enumerate(self.bns)}
            fuse = self.mode == "bf16" and self.training and self.fuse_stats
            c1 = self.convs["conv1"]
            if self.direct_stem:
                e0 = self._prof_begin("fwd")
                call("pm_stem_conv_fwd_bf16", ptr(self._x_in), ptr(self.w_stem), self.B, self.size, self.size, ptr(self.act["conv1"]),
                     ptr(self._stat_slot(bn_ids["bn1"])) if fuse else None, stream())
                self._prof_end(e0)
            else:
                self._conv_fwd(c1 if self.mode == "f32" else self.c1_gemm, self.x0, self.act["conv1"],
                               self._stat_slot(bn_ids["bn1"]) if fuse else None)
            self._stem_pool_fused = bool(self.mode == "bf16" and self.training and self.fuse_stem_pool and c1.Ho % 2 == 0
                                         and c1.Wo % 2 == 0)
            if self._stem_pool_fused:
                if not fuse:
                    call("pm_bn_stats_bf16", ptr(self.act["conv1"]), c1.P, 64, ptr(self._stat_slot(bn_ids["bn1"])), stream())
                e0 = self._prof_begin("stem_bn_pool")
                call("pm_bn_relu_maxpool_fwd_bf16", ptr(self.act["conv1"]), ptr(self._stat_slot(bn_ids["bn1"])), self.B, c1.Ho,
                     c1.Wo, 64, ctypes.c_float(self.BN_EPS), ctypes.c_float(self.BN_MOMENTUM), ptr(self.p["bn1.weight"]),
                     ptr(self.p["bn1.bias"]), ptr(self.act["p1"]), ptr(self.pool_idx), ptr(self.stem_xmax), ptr(self.bn_mean["bn1"]),
                     ptr(self.bn_invstd["bn1"]), ptr(self.p["bn1.running_mean"]), ptr(self.p["bn1.running_var"]), stream())
                self._prof_end(e0)
            else:
                self._bn_fwd("bn1", bn_ids["bn1"], self.act["conv1"], self.act["a1"], None, True, c1.P, fuse)
                if self.pooling == "avg":
                    call("pm_avgpool3s2_fwd" + self.sfx, ptr(self.act["a1"]), self.B, c1.Ho, c1.Wo, 64, ptr(self.act["p1"]), stream())
                else:
                    call("pm_maxpool3s2_fwd" + self.sfx, ptr(self.act["a1"]), self.B, c1.Ho, c1.Wo, 64, ptr(self.act["p1"]),
                         ptr(self.pool_idx), stream())
            xin = self.act["p1"]
            if getattr(self, "_cast_done", None) is not None:
                torch.cuda.current_stream().wait_event(self._cast_done)   # layer1 onwards reads the freshly cast weights
                self._cast_done = None
            for pre, ca, cb, ds in self.blocks:
                ia, ib = bn_ids[pre + ".bn1"], bn_ids[pre + ".bn2"]
                self._conv_fwd(ca, xin, self.act[ca.name], self._stat_slot(ia) if fuse else None)
                self._bn_fwd(pre + ".bn1", ia, self.act[ca.name], self.act[pre + ".a"], None, True, ca.P, fuse)
                self._conv_fwd(cb, self.act[pre + ".a"], self.act[cb.name], self._stat_slot(ib) if fuse else None)
                idn = xin
                if ds is not None:
                    idd = bn_ids[pre + ".downsample.1"]
                    self._conv_fwd(ds, xin, self.act[ds.name], self._stat_slot(idd) if fuse else None)
                    self._bn_fwd(pre + ".downsample.1", idd, self.act[ds.name], self.act[pre + ".idn"], None, False, ds.P, fuse)
                    idn = self.act[pre + ".idn"]
                self._bn_fwd(pre + ".bn2", ib, self.act[cb.name], self.act[pre + ".out"], idn, True, cb.P, fuse)
                xin = self.act[pre + ".out"]
            hw = self.final_hw * self.final_hw
            if not (self.training and self.fuse_head):   # a training step pools inside its fused head kernel (loss_and_backward)
                call("pm_gap_fwd" + self.sfx, ptr(xin), self.B, hw, 512, ptr(self.feat), stream())
            return xin

    def logits_only(self):
        with torch.cuda.device(self.device):
            call("pm_linear_fwd_f32", ptr(self.feat), ptr(self.p["fc.weight"]), ptr(self.p["fc.bias"]), self.B, 512, self.ncls,
                 ptr(self.logits), stream())
        return self.logits

    def loss_and_backward(self, target):
        """target: int64 [B] hard labels, or float [B,ncls] soft targets (Cross_entropy_one_hot, utils.py:404-441)."""
        with torch.cuda.device(self.device):
            hard = target.dtype == torch.int64
            target = target.contiguous()
            if self.mode == "bf16":
                if not getattr(self, "_grads_cleared", False):
                    self.grads.zero_()  # the tensor-core wgrad accumulates (split over pixels) into a cleared buffer
                self._grads_cleared = False
                if not self.direct_stem:
                    self.dw_stem.zero_()
                if self.overlap_wgrad:
                    self._get_side()
            e0 = self._prof_begin("head")
            bn_ids = {bn: i for i, bn in enumerate(self.bns)}
            last = self.act[self.blocks[-1][0] + ".out"]
            d_out = self._gbuf(("d", last.shape, 0), last)
            hw = self.final_hw * self.final_hw
            cwp = ptr(self.class_weights) if self.class_weights is not None else None
            if self.training and self.fuse_head:
                # ONE launch on the critical path: average pool, logits, per-sample CE gradient, dfeat, gradient of the pool;
                # the batch reductions (dW, db, loss) are not needed by the backward chain and go to the side stream
                call("pm_head_fused" + self.sfx, ptr(last), hw, ptr(self.p["fc.weight"]), ptr(self.p["fc.bias"]),
                     ptr(target) if hard else None, None if hard else ptr(target), cwp, self.B, 512, self.ncls, ptr(self.feat),
                     ptr(self.logits), ptr(self.head_ws), ptr(self.dfeat), ptr(d_out), stream())

                def head_grads():
                    call("pm_head_grads_f32", ptr(self.feat), self.B, 512, self.ncls, ptr(self.head_ws), ptr(self.loss),
                         ptr(self.g["fc.weight"]), ptr(self.g["fc.bias"]), stream())

                if self._side is not None and self._prof is None:
                    ev = torch.cuda.Event()
                    ev.record()
                    self._side.wait_event(ev)
                    with torch.cuda.stream(self._side):
                        head_grads()
                else:
                    head_grads()
            else:
                call("pm_linear_ce_f32", ptr(self.feat), ptr(self.p["fc.weight"]), ptr(self.p["fc.bias"]),
                     ptr(target) if hard else None, None if hard else ptr(target), cwp, self.B, 512, self.ncls,
                     ptr(self.logits), ptr(self.loss), ptr(self.dfeat), ptr(self.g["fc.weight"]), ptr(self.g["fc.bias"]),
                     ptr(self.head_ws), stream())
                call("pm_gap_bwd" + self.sfx, ptr(self.dfeat), self.B, hw, 512, ptr(d_out), stream())
            self._prof_end(e0)
            for bi in range(len(self.blocks) - 1, -1, -1):
                pre, ca, cb, ds = self.blocks[bi]
                if self._split_cb is not None and bi == self.SPLIT_BLOCK - 1:
                    # every gradient of layer4 + fc exists once the side stream has joined: the caller takes over here
                    # (optimizer step on that segment, start of its all-reduce, graph boundary)
                    if self._side is not None:
                        torch.cuda.current_stream().wait_stream(self._side)
                    self._split_cb()
                xin = self.act[self.blocks[bi - 1][0] + ".out"] if bi > 0 else self.act["p1"]
                out = self.act[pre + ".out"]
                # out = relu(bn2(cB) + idn).  g = d_out * (out > 0) may alias d_out (element-wise in place).
                g = self._gbuf(("g", out.shape), out)
                dcb = self._gbuf(("dc", cb.name), out)
                self._bn_bwd(pre + ".bn2", bn_ids[pre + ".bn2"], d_out, out, self.act[cb.name], dcb, cb.P, g_out=g)
                if ds is not None:
                    d_xin = self._gbuf(("dx", xin.shape), xin)
                    dcd = self._gbuf(("dc", ds.name), out)
                    self._bn_bwd(pre + ".downsample.1", bn_ids[pre + ".downsample.1"], g, None, self.act[ds.name], dcd, ds.P)
                    self._conv_wgrad(ds, xin, dcd)
                    self._conv_dgrad(ds, dcd, d_xin, False)
                else:
                    d_xin = g  # identity branch: the masked gradient flows straight through; conv1's dgrad accumulates
                self._conv_wgrad(cb, self.act[pre + ".a"], dcb)
                d_a = self._gbuf(("da", out.shape), out)
                self._conv_dgrad(cb, dcb, d_a, False)
                dca = self._gbuf(("dc", ca.name), out)
                self._bn_bwd(pre + ".bn1", bn_ids[pre + ".bn1"], d_a, self.act[pre + ".a"], self.act[ca.name], dca, ca.P)
                self._conv_wgrad(ca, xin, dca)
                self._conv_dgrad(ca, dca, d_xin, True)
                d_out = d_xin
            c1 = self.convs["conv1"]
            dc1 = self._gbuf(("dc1",), self.act["conv1"])
            if getattr(self, "_stem_pool_fused", False):
                # the argmax table already encodes the ReLU decision (255 = no gradient): neither the 112x112 activation nor a
                # full-resolution gradient is ever materialised; reduce + apply launches over 2x2 input blocks
                e0 = self._prof_begin("stem_bn_pool")
                call("pm_stem_pool_bn_bwd_bf16", ptr(d_out), ptr(self.pool_idx), ptr(self.stem_xmax), self.B, c1.Ho, c1.Wo,
                     ptr(self.act["conv1"]),
                     ptr(self.bn_mean["bn1"]), ptr(self.bn_invstd["bn1"]), ptr(self.p["bn1.weight"]), 64,
                     ptr(self._stat_slot(len(self.bns) + bn_ids["bn1"])), ptr(dc1), ptr(self.g["bn1.weight"]),
                     ptr(self.g["bn1.bias"]), stream())
                self._prof_end(e0)
            elif self.pooling == "avg":
                d_a1 = self._gbuf(("da1",), self.act["a1"])
                call("pm_avgpool3s2_bwd" + self.sfx, ptr(d_out), self.B, c1.Ho, c1.Wo, 64, ptr(d_a1), stream())
                self._bn_bwd("bn1", bn_ids["bn1"], d_a1, self.act["a1"], self.act["conv1"], dc1, c1.P)
            elif self.fuse_bn_bwd:
                # max-pool backward is gathered inside the stem's BN backward (no full-resolution gradient round trip)
                call("pm_bn_bwd_fused_pool" + self.sfx, ptr(d_out), ptr(self.pool_idx), self.B, c1.Ho, c1.Wo, ptr(self.act["a1"]),
                     ptr(self.act["conv1"]), ptr(self.bn_mean["bn1"]), ptr(self.bn_invstd["bn1"]), ptr(self.p["bn1.weight"]), 64,
                     ptr(self.bn_bwd_ws), ptr(self._gbuf(("da1",), self.act["a1"])), ptr(dc1), ptr(self.g["bn1.weight"]),
                     ptr(self.g["bn1.bias"]), stream())
            else:
                d_a1 = self._gbuf(("da1",), self.act["a1"])
                call("pm_maxpool3s2_bwd" + self.sfx, ptr(d_out), ptr(self.pool_idx), self.B, c1.Ho, c1.Wo, 64, ptr(d_a1), stream())
                self._bn_bwd("bn1", bn_ids["bn1"], d_a1, self.act["a1"], self.act["conv1"], dc1, c1.P)
            if self.mode == "f32":
                self._conv_wgrad(c1, self.x0, dc1)
            else:
                def stem_wgrad():
                    if self.direct_stem:  # accumulates straight into the (cleared) KRSC gradient
                        call("pm_stem_conv_wgrad_bf16", ptr(self._x_in), ptr(dc1), self.B, self.size, self.size,
                             ptr(self.g["conv1.weight"]), stream())
                    else:
                        call("pm_conv_wgrad_bf16", ctypes.byref(self.c1_gemm.desc), ptr(self.x0), ptr(dc1), ptr(self.dw_stem), None, stream())
                        self.g["conv1.weight"].view(64, -1).copy_(self.dw_stem[:, : 7 * 7 * self.cin])

                if self._side is not None and self._prof is None:
                    ev = torch.cuda.Event()
                    ev.record()
                    self._side.wait_event(ev)
                    with torch.cuda.stream(self._side):
                        stem_wgrad()
                else:
                    e0 = self._prof_begin("wgrad")
                    stem_wgrad()
                    self._prof_end(e0)
            if self._side is not None:
                torch.cuda.current_stream().wait_stream(self._side)  # join: the optimizer needs every weight gradient
        return self.loss

    def optimizer_step(self, lo=0, hi=None, bump=True):
        """the optimizer on flat[lo:hi] (default: all parameters); ``bump`` advances the step index (once per step)"""
        with torch.cuda.device(self.device):
            if bump:
                self.step_count += 1
                if self.step_count > 1:
                    self._mv_zero = False   # the moments now hold what the first step wrote
            hi = self.n_param_flat if hi is None else hi
            n = hi - lo
            sl = lambda t: ctypes.c_void_p(t.data_ptr() + 4 * lo)
            e0 = self._prof_begin("optimizer")
            if self.opt_name == "Adam" and self._mv_zero and self.step_count == 1:
                # freshly (re-)created optimizer: the moments are logically zero and are not even read
                call("pm_adam_first_step_f32", sl(self.flat), sl(self.grads), sl(self.adam_m), sl(self.adam_v), n,
                     ctypes.c_float(self.lr), ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                     ctypes.c_float(self.opt_eps), ctypes.c_float(self.wd), stream())
            elif self.opt_name == "Adam":
                call("pm_adam_step_f32", sl(self.flat), sl(self.grads), sl(self.adam_m), sl(self.adam_v), n,
                     ctypes.c_float(self.lr), ctypes.c_float(self.betas[0]), ctypes.c_float(self.betas[1]),
                     ctypes.c_float(self.opt_eps), ctypes.c_float(self.wd), self.step_count, stream())
            elif self.opt_name == "SGD":
                call("pm_sgd_step_f32", sl(self.flat), sl(self.grads), n, ctypes.c_float(self.lr), ctypes.c_float(self.wd),
                     stream())
            else:
                raise NotImplementedError("only Adam or SGD supported.")  # utils.py:1141
            self._prof_end(e0)

    def reset_optimizer(self):
        """utils.py:1131-1145,1209-1218: optimizers are re-created (state zeroed) after every aggregation.  Lazily: the next
        step runs pm_adam_first_step_f32, which treats the moments as zero without reading them, so nothing is cleared here."""
        self._mv_zero = True
        self.step_count = 0

    def adam_moments(self):
        """(m, v) as the optimizer state really is (zeros right after a reset)"""
        if self._mv_zero and self.step_count == 0:
            return torch.zeros_like(self.adam_m), torch.zeros_like(self.adam_v)
        return self.adam_m, self.adam_v

    def _train_step_eager(self, x_nchw, target):
        self.forward(x_nchw)
        loss = self.loss_and_backward(target)
        self.optimizer_step()
        return loss

    def train_step(self, x_nchw, target):
        """utils.py:1168-1174: zero_grad; pred = model(data); loss; backward; step.
        Replays the captured CUDA graph when one exists for this (optimizer step index, target kind)."""
        gr = self._graph
        if (gr is not None and gr["step"] == self.step_count + 1 and gr["tdtype"] == target.dtype
                and gr["hyper"] == self._hyper_key()):
            gr["x"].copy_(x_nchw, non_blocking=True)   # (also expands a bf16-staged batch to the fp32 the stem's TMA boxes read)
            gr["y"].copy_(target, non_blocking=True)
            if self.on_inputs_staged is not None:
                self.on_inputs_staged()                # the caller's staging slot is free again: the next H2D copy may start
            gr["graph"].replay()
            self.step_count += 1
            return self.loss
        return self._train_step_eager(x_nchw, target)

    def _hyper_key(self):
        """everything a captured step bakes into kernel arguments besides the Adam step index: a replay is legal only while
        none of it changed (train.py:433-440 adjusts lr per epoch; class weights / eval mode switch kernels)"""
        cw = None if self.class_weights is None else self.class_weights.data_ptr()
        return (self.opt_name, float(self.lr), tuple(float(b) for b in self.betas), float(self.wd), float(self.opt_eps), cw,
                bool(self.training), bool(self._mv_zero))

    def capture_graph(self, x_nchw, target):
        """Capture one local step (all ~190 kernel launches) in a CUDA graph.  Adam's bias correction bakes the step
        index into the graph, so the replay is used only when the optimizer is at the captured step (always true for
        the reference's default FedAvg-after-every-step + optimizer-reset schedule, utils.py:1175-1218)."""
        from .. import _lib

        with torch.cuda.device(self.device):
            snap = (self.flat.clone(), self.adam_m.clone(), self.adam_v.clone(), self.step_count, self._mv_zero)
            gx, gy = x_nchw.clone(), target.clone()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._train_step_eager(gx, gy)  # warm-up on the capture stream: allocates every lazily created buffer
            torch.cuda.current_stream().wait_stream(side)
            self.flat.copy_(snap[0]); self.adam_m.copy_(snap[1]); self.adam_v.copy_(snap[2]); self.step_count = snap[3]
            self._mv_zero = snap[4]
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            n0 = _lib.launch_counter
            with torch.cuda.graph(graph):
                self._train_step_eager(gx, gy)
            launches = _lib.launch_counter - n0 + 1  # + stats.zero_()
            self.step_count = snap[3]
            self._mv_zero = snap[4]
            self._graph = {"graph": graph, "x": gx, "y": gy, "step": snap[3] + 1, "tdtype": target.dtype, "launches": launches,
                           "hyper": self._hyper_key()}
        return launches

    # ------------------------------------------------------------------ FedAvg overlapped with the backward pass
    @property
    def split_offset(self):
        """flat[split_offset:] = layer4 + fc parameters followed by ALL BatchNorm running statistics (33.6 MB of the 44.75 MB
        state); flat[:split_offset] = stem + layer1-3 parameters"""
        return self.offsets[self.blocks[self.SPLIT_BLOCK][1].name + ".weight"][0]

    def capture_graph_overlap(self, x_nchw, target):
        """The local step as TWO CUDA graphs, so that the all-reduce of most of the state overlaps the rest of the backward:

          graph A: forward, backward of layer4, optimizer step on flat[split_offset:n_param]    -> all-reduce flat[split_offset:]
          graph B: backward of layer3..1 + stem, optimizer step on flat[:split_offset]          -> all-reduce flat[:split_offset]

        The running statistics (end of the flat buffer) are final after the forward, so they travel with the first bucket.
        Same replay conditions as ``capture_graph`` (optimizer step index, hyper-parameters)."""
        with torch.cuda.device(self.device):
            snap = (self.flat.clone(), self.adam_m.clone(), self.adam_v.clone(), self.step_count, self._mv_zero)
            gx, gy = x_nchw.clone(), target.clone()
            off = self.split_offset
            cap = torch.cuda.Stream(self.device)
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                self._train_step_eager(gx, gy)  # warm-up: allocates every lazily created buffer
            torch.cuda.current_stream().wait_stream(cap)
            self.flat.copy_(snap[0]); self.adam_m.copy_(snap[1]); self.adam_v.copy_(snap[2]); self.step_count = snap[3]
            self._mv_zero = snap[4]
            torch.cuda.synchronize(self.device)
            gA, gB = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()

            def boundary():
                self.optimizer_step(off, self.n_param_flat, bump=True)
                gA.capture_end()
                gB.capture_begin(pool=gA.pool())

            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                self._split_cb = boundary
                try:
                    gA.capture_begin()
                    self.forward(gx)
                    self.loss_and_backward(gy)
                    self.optimizer_step(0, off, bump=False)
                    gB.capture_end()
                finally:
                    self._split_cb = None
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize(self.device)
            self.step_count = snap[3]
            self._mv_zero = snap[4]
            self._graph2 = {"A": gA, "B": gB, "x": gx, "y": gy, "step": snap[3] + 1, "tdtype": target.dtype,
                            "hyper": self._hyper_key(), "off": off}

    def train_step_overlapped(self, x_nchw, target, after_a, after_b):
        """replay graph A, call ``after_a()`` (start the first bucket's all-reduce), replay graph B, call ``after_b()``"""
        gr = self._graph2
        if not (gr is not None and gr["step"] == self.step_count + 1 and gr["tdtype"] == target.dtype
                and gr["hyper"] == self._hyper_key()):
            raise PrimiaError("no overlap graphs valid for this step: capture_graph_overlap first (same optimizer step index and "
                              "hyper-parameters)")
        gr["x"].copy_(x_nchw, non_blocking=True)
        gr["y"].copy_(target, non_blocking=True)
        if self.on_inputs_staged is not None:
            self.on_inputs_staged()
        gr["A"].replay()
        after_a()
        gr["B"].replay()
        self.step_count += 1
        after_b()
        return self.loss

    def profile_conv_time(self, x_nchw, target, steps=2):
        """Device time of the convolution kernels per local step (CUDA events on the launching stream around each of the
        60 conv launches, eager mode) and the number of kernel launches per step."""
        from .. import _lib

        with torch.cuda.device(self.device):
            snap = (self.flat.clone(), self.adam_m.clone(), self.adam_v.clone(), self.step_count, self._mv_zero)
            self._train_step_eager(x_nchw, target)
            torch.cuda.synchronize(self.device)
            self._prof = []
            n0 = _lib.launch_counter
            for _ in range(steps):
                self._train_step_eager(x_nchw, target)
            torch.cuda.synchronize(self.device)
            launches = (_lib.launch_counter - n0) // steps + 1
            conv_tags = ("fwd", "dgrad", "wgrad")
            ms = sum(a.elapsed_time(b) for a, b, t in self._prof if t in conv_tags) / steps
            self.conv_ms_by_kind, self.step_ms_by_family = {}, {}
            for a, b, tag in self._prof:
                d = self.conv_ms_by_kind if tag in conv_tags else self.step_ms_by_family
                d[tag] = d.get(tag, 0.0) + a.elapsed_time(b) / steps
            self.step_ms_by_family["conv (fwd + dgrad + wgrad)"] = ms
            self._prof = None
            self.flat.copy_(snap[0]); self.adam_m.copy_(snap[1]); self.adam_v.copy_(snap[2]); self.step_count = snap[3]
            self._mv_zero = snap[4]
        return ms, launches

    def init_random(self, seed=42):
        """Random-init weights of the reference architecture (models.py:408-413: Kaiming-normal fan_out convs, BN 1/0;
        nn.Linear default init for fc), generated on the device."""
        with torch.cuda.device(self.device):
            g = torch.Generator(device=self.device).manual_seed(seed)
            self.flat.zero_()
            for name, c in self.convs.items():
                w = self.p[name + ".weight"]
                std = (2.0 / (c.K * c.R * c.R)) ** 0.5
                w.copy_(torch.randn(w.shape, device=self.device, generator=g) * std)
            for bn in self.bns:
                self.p[bn + ".weight"].fill_(1.0)
                self.p[bn + ".running_var"].fill_(1.0)
            bound = 1.0 / (512 ** 0.5)
            self.p["fc.weight"].copy_((torch.rand((self.ncls, 512), device=self.device, generator=g) * 2 - 1) * bound)
            self.p["fc.bias"].copy_((torch.rand((self.ncls,), device=self.device, generator=g) * 2 - 1) * bound)
            self.reset_optimizer()
