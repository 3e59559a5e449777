"""DP-SGD local step (train.py:304-334: ``torchdp.PrivacyEngine(model, batch_size, sample_size, alphas, noise_multiplier,
max_grad_norm).attach(optimizer)``) on the GPU engine.

Algorithm (pytorch-dp 0.1b1 -- third-party, source NOT in the reference tree, so parity is unpinned; restated from its
published behaviour in ``oracle/dp_oracle.py``): per-sample gradients g_b of each sample's own loss; clip factor
``c_b = min(1, C / (|g_b| + 1e-6))`` over the whole parameter vector; ``grad = (sum_b c_b g_b + N(0, (sigma C)^2)) / B``; then the
attached optimizer's ordinary step.

The reference refuses this combination outright (train.py:306-310: "only implemented for local training and models without
BatchNorm") because batch statistics couple the samples of a batch and per-sample gradients stop existing.  Here the
BatchNorm layers act as frozen per-channel affine maps during a DP step (running statistics, as ``model.eval()`` would; gamma
and beta still train), which makes every sample's gradient well defined, and the step runs per hospital inside the federated
round (BASELINE.json configs[2]).

B200 mapping: the per-sample weight gradients are materialised layer by layer ([B][n] fp32; 5.7 GB at B = 128) by a tcgen05
weight-gradient kernel whose grid.z is the sample (``pm_conv_wgrad_persample_bf16``), then two HBM-bound passes (squared norms;
clipped sum) and one Philox/Box-Muller noise pass over the flat gradient."""
from __future__ import annotations

import ctypes

import torch

from .._lib import PrimiaError, call, ptr, stream
from .resnet18 import ResNet18Engine


class DPState:
    """buffers of one engine's DP steps (allocated once)"""

    def __init__(self, eng: ResNet18Engine):
        dev, B = eng.device, eng.B
        self.ps = {}  # per-sample gradient blocks, [B, n] fp32 per parameter tensor (except fc: never materialised)
        for name, shape in eng.param_order:
            if name.startswith("fc."):
                continue
            n = 1
            for d in shape:
                n *= d
            if name == "conv1.weight" and eng.mode == "bf16":
                n = 64 * eng.stem_kpad  # the stem runs as a 1x1 problem over the padded im2col tensor
            self.ps[name] = torch.empty((B, n), dtype=torch.float32, device=dev)
        self.norm2 = torch.zeros(B, dtype=torch.float64, device=dev)
        self.factors = torch.empty(B, dtype=torch.float32, device=dev)
        self.norms = torch.empty(B, dtype=torch.float32, device=dev)
        self.noise_flat = None
        self.x0 = None
        if eng.mode == "bf16":
            c1 = eng.convs["conv1"]
            self.x0 = torch.empty((B, c1.Ho, c1.Wo, eng.stem_kpad), dtype=torch.bfloat16, device=dev)
            self.w_stem = torch.empty((64, 1, 1, eng.stem_kpad), dtype=torch.bfloat16, device=dev)
            self.dw_stem = torch.empty((64, eng.stem_kpad), dtype=torch.float32, device=dev)
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)  # device-side noise counter (bumped by the noise launch)
        self.seed = None

    def nbytes(self):
        return sum(t.numel() * 4 for t in self.ps.values())


def _bn_eval_bwd(eng, st, bn, dy, y_out, x, dx, P, g_out=None):
    """BatchNorm (frozen statistics) backward + the per-sample gamma / beta gradients"""
    C = eng.bns[bn]
    g = g_out if g_out is not None else eng._gbuf(("dpg", tuple(dy.shape)), dy)
    call("pm_bn_eval_bwd" + eng.sfx, ptr(dy), ptr(y_out) if y_out is not None else None, ptr(eng.p[bn + ".weight"]),
         ptr(eng.bn_invstd[bn]), P, C, ptr(g), ptr(dx), stream())
    call("pm_bn_persample_param_grads" + eng.sfx, ptr(g), ptr(x), ptr(eng.bn_mean[bn]), ptr(eng.bn_invstd[bn]), eng.B, P // eng.B, C,
         ptr(st.ps[bn + ".weight"]), ptr(st.ps[bn + ".bias"]), stream())


def _wgrad_ps(eng, st, c, x, dy, name=None):
    out = st.ps[name or (c.name + ".weight")]
    if eng.mode == "f32":
        call("pm_conv_wgrad_persample_f32", ctypes.byref(c.desc), ptr(x), ptr(dy), ptr(out), ptr(eng.wgrad_ws), stream())
    else:  # the epilogue also accumulates |dw_b|^2 into the per-sample norms
        call("pm_conv_wgrad_persample_bf16", ctypes.byref(c.desc), ptr(x), ptr(dy), ptr(out), ptr(st.norm2), stream())


def _graph_key(eng, noise_multiplier, max_grad_norm, seed):
    return (eng.step_count + 1, float(noise_multiplier), float(max_grad_norm), seed, eng._hyper_key())


def dp_train_step(eng: ResNet18Engine, x_nchw, target, noise_multiplier=1.3, max_grad_norm=1.0, noise=None, seed=None):
    """one DP-SGD local step; returns the loss (device scalar).  ``noise``: explicit {name: N(0, (sigma C)^2) tensor in the
    reference layout} (tests); otherwise Philox noise from ``seed`` (default: a per-engine secret drawn from the OS) and a
    device-side counter that advances every step.  Replays the graph captured by ``capture_dp_graph`` when one matches."""
    gr = getattr(eng, "_dp_graph", None)
    if gr is not None and noise is None and gr["key"] == _graph_key(eng, noise_multiplier, max_grad_norm, seed):
        gr["x"].copy_(x_nchw, non_blocking=True)
        gr["y"].copy_(target, non_blocking=True)
        gr["graph"].replay()
        eng.step_count += 1
        return eng.loss
    return _dp_step_eager(eng, x_nchw, target, noise_multiplier, max_grad_norm, noise, seed)


def capture_dp_graph(eng: ResNet18Engine, x_nchw, target, noise_multiplier=1.3, max_grad_norm=1.0, seed=None):
    """capture one DP step (~450 launches) in a CUDA graph; the Philox counter lives in device memory, so replays draw fresh
    noise.  Valid at the captured optimizer step index / hyper-parameters (see ResNet18Engine.capture_graph)."""
    with torch.cuda.device(eng.device):
        snap = (eng.flat.clone(), eng.adam_m.clone(), eng.adam_v.clone(), eng.step_count, eng._mv_zero)
        gx, gy = x_nchw.clone(), target.clone()
        side = torch.cuda.Stream(eng.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            _dp_step_eager(eng, gx, gy, noise_multiplier, max_grad_norm, None, seed)  # warm-up: allocates every buffer
        torch.cuda.current_stream().wait_stream(side)
        eng.flat.copy_(snap[0]); eng.adam_m.copy_(snap[1]); eng.adam_v.copy_(snap[2]); eng.step_count = snap[3]
        eng._mv_zero = snap[4]
        torch.cuda.synchronize(eng.device)
        key = _graph_key(eng, noise_multiplier, max_grad_norm, seed)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            _dp_step_eager(eng, gx, gy, noise_multiplier, max_grad_norm, None, seed)
        eng.step_count = snap[3]
        eng._mv_zero = snap[4]
        eng._dp_graph = {"graph": graph, "x": gx, "y": gy, "key": key}


def _dp_step_eager(eng: ResNet18Engine, x_nchw, target, noise_multiplier, max_grad_norm, noise, seed):
    if eng.class_weights is not None or target.dtype != torch.int64:
        raise PrimiaError("DP-SGD step: hard labels without class weights (the per-sample loss must be a plain cross-entropy)")
    st = getattr(eng, "_dp", None)
    if st is None:
        st = eng._dp = DPState(eng)
    B = eng.B
    with torch.cuda.device(eng.device):
        was_training = eng.training
        eng.training = False           # BatchNorm as a frozen affine map: running statistics, no update
        try:
            eng.forward(x_nchw)
        finally:
            eng.training = was_training
        st.norm2.zero_()
        target = target.contiguous()
        call("pm_linear_ce_f32", ptr(eng.feat), ptr(eng.p["fc.weight"]), ptr(eng.p["fc.bias"]), ptr(target), None, None, B, 512,
             eng.ncls, ptr(eng.logits), ptr(eng.loss), ptr(eng.dfeat), ptr(eng.g["fc.weight"]), ptr(eng.g["fc.bias"]),
             ptr(eng.head_ws), stream())
        bn_ids = {bn: i for i, bn in enumerate(eng.bns)}
        last = eng.act[eng.blocks[-1][0] + ".out"]
        d_out = eng._gbuf(("d", last.shape, 0), last)
        hw = eng.final_hw * eng.final_hw
        call("pm_gap_bwd" + eng.sfx, ptr(eng.dfeat), B, hw, 512, ptr(d_out), stream())
        for bi in range(len(eng.blocks) - 1, -1, -1):
            pre, ca, cb, ds = eng.blocks[bi]
            xin = eng.act[eng.blocks[bi - 1][0] + ".out"] if bi > 0 else eng.act["p1"]
            out = eng.act[pre + ".out"]
            g = eng._gbuf(("g", out.shape), out)
            dcb = eng._gbuf(("dc", cb.name), out)
            _bn_eval_bwd(eng, st, pre + ".bn2", d_out, out, eng.act[cb.name], dcb, cb.P, g_out=g)
            if ds is not None:
                d_xin = eng._gbuf(("dx", xin.shape), xin)
                dcd = eng._gbuf(("dc", ds.name), out)
                _bn_eval_bwd(eng, st, pre + ".downsample.1", g, None, eng.act[ds.name], dcd, ds.P)
                _wgrad_ps(eng, st, ds, xin, dcd)
                eng._conv_dgrad(ds, dcd, d_xin, False)
            else:
                d_xin = g
            _wgrad_ps(eng, st, cb, eng.act[pre + ".a"], dcb)
            d_a = eng._gbuf(("da", out.shape), out)
            eng._conv_dgrad(cb, dcb, d_a, False)
            dca = eng._gbuf(("dc", ca.name), out)
            _bn_eval_bwd(eng, st, pre + ".bn1", d_a, eng.act[pre + ".a"], eng.act[ca.name], dca, ca.P)
            _wgrad_ps(eng, st, ca, xin, dca)
            eng._conv_dgrad(ca, dca, d_xin, True)
            d_out = d_xin
        c1 = eng.convs["conv1"]
        dc1 = eng._gbuf(("dc1",), eng.act["conv1"])
        d_a1 = eng._gbuf(("da1",), eng.act["a1"])
        pool = "pm_avgpool3s2_bwd" if eng.pooling == "avg" else "pm_maxpool3s2_bwd"
        if eng.pooling == "avg":
            call(pool + eng.sfx, ptr(d_out), B, c1.Ho, c1.Wo, 64, ptr(d_a1), stream())
        else:
            call(pool + eng.sfx, ptr(d_out), ptr(eng.pool_idx), B, c1.Ho, c1.Wo, 64, ptr(d_a1), stream())
        _bn_eval_bwd(eng, st, "bn1", d_a1, eng.act["a1"], eng.act["conv1"], dc1, c1.P)
        if eng.mode == "f32":
            _wgrad_ps(eng, st, c1, eng.x0, dc1)
        else:  # the stem as a dense 1x1 problem over the bf16 im2col tensor (k = (r*7+s)*Cin + c, zero padded to stem_kpad)
            call("pm_im2col_stem_bf16", ptr(eng._x_in if eng._x_in is not None else x_nchw.contiguous()), B, eng.cin, eng.size, eng.size,
                 7, 2, 3, eng.stem_kpad, ptr(st.x0), stream())
            _wgrad_ps(eng, st, eng.c1_gemm, st.x0, dc1, "conv1.weight")
        # ---- per-sample norms over the WHOLE parameter vector, clip factors, clipped sum
        for name, blk in st.ps.items():
            if eng.mode == "bf16" and len(eng.offsets[name][2]) == 4:
                continue   # conv weights: already accumulated by the per-sample weight-gradient epilogue
            call("pm_dp_sqnorm_f32", ptr(blk), B, blk.shape[1], ptr(st.norm2), stream())
        ld, inv = eng.ncls + 1, 1.0 / B   # pm_linear_ce_f32 left the unnormalised per-sample dlogits in head_ws
        call("pm_dp_fc_sqnorm_f32", ptr(eng.head_ws), ld, ctypes.c_float(inv), ptr(eng.feat), B, 512, eng.ncls, ptr(st.norm2), stream())
        # everything above carries the 1/B of the mean loss: |g_b| = B * |piece_b|
        call("pm_dp_clip_factors", ptr(st.norm2), B, ctypes.c_double(float(B)), ctypes.c_double(float(max_grad_norm)), ptr(st.factors),
             ptr(st.norms), stream())
        for name, blk in st.ps.items():
            dst = eng.g[name]
            if name == "conv1.weight" and eng.mode == "bf16":
                call("pm_dp_weighted_sum_f32", ptr(blk), ptr(st.factors), B, blk.shape[1], ptr(st.dw_stem), 0, stream())
                dst.view(64, -1).copy_(st.dw_stem[:, : 7 * 7 * eng.cin])
            else:
                call("pm_dp_weighted_sum_f32", ptr(blk), ptr(st.factors), B, blk.shape[1], ptr(dst), 0, stream())
        call("pm_dp_fc_weighted_f32", ptr(eng.head_ws), ld, ctypes.c_float(inv), ptr(eng.feat), ptr(st.factors), B, 512, eng.ncls,
             ptr(eng.g["fc.weight"]), ptr(eng.g["fc.bias"]), stream())
        # ---- grad = (sum_b c_b g_b + N(0, (sigma C)^2)) / B ; the sum above already carries the 1/B
        n = eng.n_param_flat
        if noise is not None:
            if st.noise_flat is None:
                st.noise_flat = torch.zeros(n, dtype=torch.float32, device=eng.device)
            eng.torch_layout_to_flat(noise, st.noise_flat)
            call("pm_dp_axpy_scale_f32", ptr(eng.grads), ptr(st.noise_flat), ctypes.c_float(inv), ctypes.c_float(1.0), n, stream())
        elif noise_multiplier > 0:
            if seed is None:
                if st.seed is None:
                    import secrets

                    st.seed = secrets.randbits(63)
                seed = st.seed
            call("pm_dp_add_noise_f32", ptr(eng.grads), n, ctypes.c_float(noise_multiplier * max_grad_norm * inv), seed, 1,
                 ptr(st.counter), stream())
        eng.optimizer_step()
    return eng.loss
