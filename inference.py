#!/usr/bin/env python
"""inference.py -- entry point with the reference's CLI (inference.py:47-76).

    python inference.py --model_weights model_weights/x.pt                       plain GPU inference
    python inference.py --model_weights ... --encrypted_inference                SPDZ fixed-precision forward on shares

Encrypted mode follows inference.py:154-158,279-321: VirtualWorkers data_owner / model_owner / crypto_provider (each a
GPU when 3 are visible), ``.fix_precision(precision_fractional=16, dtype="long").share(*workers, crypto_provider=...,
protocol="fss")`` for the weights and the image, then the forward pass on shares.  Built so far: every linear layer
(conv/fc Beaver matmul + truncation), BatchNorm on shares, average pooling, reconstruction/decoding.  The comparison-
based layers (ReLU, max-pool: function secret sharing) are the next row, so the encrypted mode
currently runs and times the linear-layer protocol of one image and says so.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv=None):
    import primia_b200.sy as sy
    from primia_b200.train import ResNet18Engine

    ap = argparse.ArgumentParser()
    ap.add_argument("--model_weights", default=None)
    ap.add_argument("--data_dir", default=None, help="accepted for CLI compatibility; a synthetic image is used")
    ap.add_argument("--encrypted_inference", action="store_true")
    ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--precision_fractional", type=int, default=16)
    ap.add_argument("--batch", type=int, default=1)
    cmd = ap.parse_args(argv)
    tick = time.time()
    hook = sy.TorchHook(torch)
    state = None
    if cmd.model_weights:
        ck = torch.load(cmd.model_weights, map_location="cpu", weights_only=False)
        state = ck["model_state_dict"]
    g = torch.Generator().manual_seed(42)
    img = torch.randn(cmd.batch, 3, 224, 224, generator=g)
    if not cmd.encrypted_inference:
        eng = ResNet18Engine(cmd.batch, 3, 3, 224, "max", "cuda:0", "f32")
        eng.init_random(42) if state is None else eng.load_state_dict(state)
        eng.training = False
        eng.forward(img.cuda())
        logits = eng.logits_only()
        pred = logits.argmax(1)
        print("prediction:", pred.tolist())
    else:
        from primia_b200.ring.resnet import SharedLinearLayers

        data_owner = sy.VirtualWorker(hook, id="data_owner")
        crypto_provider = sy.VirtualWorker(hook, id="crypto_provider")
        model_owner = sy.VirtualWorker(hook, id="model_owner")
        workers = [model_owner, data_owner]
        sy.local_worker.clients = workers
        prov = sy.make_crypto_provider(crypto_provider)
        net = SharedLinearLayers(workers, prov, 10, cmd.precision_fractional)
        xs = net.make_inputs(cmd.batch)
        net.preprocess(cmd.batch, 1)
        torch.cuda.synchronize()
        t0 = time.time()
        out = net.forward(xs)
        logits = out["fc"].get().float_precision()
        torch.cuda.synchronize()
        print(f"encrypted linear layers (20 convs + fc, Beaver protocol, pf={cmd.precision_fractional}): "
              f"{(time.time() - t0) * 1e3:.2f} ms/image; fc output {logits.flatten().tolist()}")
        print("ReLU / max-pool on shares (FSS) are not built yet: this is the linear-layer protocol only.")
    print(f"Took {time.time() - tick:.2f} seconds.")  # inference.py:326-328


if __name__ == "__main__":
    main()
