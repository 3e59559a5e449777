#!/usr/bin/env python
"""inference.py -- entry point with the reference's CLI (inference.py:47-76) and flow (inference.py:77-328).

    python inference.py --model_weights model_weights/x.pt [--data_dir ...]                  plain GPU inference
    python inference.py --model_weights model_weights/x.pt --encrypted_inference             SPDZ / FSS inference on shares

Encrypted mode is the reference's, verb for verb: VirtualWorkers ``data_owner`` / ``model_owner`` / ``crypto_provider`` (one
GPU each when three are visible, inference.py:154-158), the images tagged and loaded onto the data owner and found through
``PrivateGridNetwork.search`` (:213-229), ``model.fix_precision(precision_fractional=16, dtype="long").share(*workers,
crypto_provider=..., protocol="fss", requires_grad=False)`` (:279-286), ``model.pool`` and ``model.relu`` swapped (:289), and per
image ``data.fix_precision(..).share(..).get()`` -> ``model(data)`` -> ``.get().float_prec()`` -> argmax (:292-317).  Every
arithmetic step is the primia_b200 C ABI (Beaver matmuls on the int8 tensor cores, 80-step Newton BatchNorm, FSS ReLU /
max-pool).  ``--cuda_graph`` runs each image as two captured CUDA graphs -- offline (primitives, Newton, the model-only halves of
every Beaver product) and online (554 launches in one replay instead of ~1.5 k host round trips); same shares either way; ``--precision_fractional`` defaults to the reference's 16.

Image files / albumentations are out of scope: the data set is ``--num_images`` synthetic normalised
224 x 224 x 3 tensors (``--data_dir`` is accepted and ignored).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from collections import Counter

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main(argv=None):
    import primia_b200.sy as sy
    from torchlib.dataloader import RemoteTensorDataset
    from torchlib.models import resnet18
    from torchlib.utils import load_checkpoint

    tick = time.time()
    ap = argparse.ArgumentParser()
    ap.add_argument("--data_dir", default=None, help="accepted for CLI compatibility; synthetic images are classified")
    ap.add_argument("--model_weights", type=str, required=True, default=None, help="model weights to use")
    ap.add_argument("--encrypted_inference", action="store_true", help="Perform encrypted inference")
    ap.add_argument("--websockets_config", default=None, help="network workers are out of scope: refused")
    ap.add_argument("--cuda", action="store_true", help="accepted; the GPU is the only device")
    ap.add_argument("--http_protocol", action="store_true", help="accepted and ignored (no network workers)")
    ap.add_argument("--num_images", type=int, default=2)
    ap.add_argument("--precision_fractional", type=int, default=16)
    ap.add_argument("--cuda_graph", action="store_true", help="encrypted mode: replay each image's online phase as one CUDA graph")
    cmd_args = ap.parse_args(argv)
    if cmd_args.websockets_config:
        raise SystemExit("websocket / PyGrid workers are out of scope (SURVEY.md section 2): use VirtualWorkers")
    if not torch.cuda.is_available():
        raise SystemExit("primia_b200 needs a CUDA device (there is no CPU path)")
    device = torch.device("cuda")
    state = load_checkpoint(cmd_args.model_weights, map_location="cpu")
    args = state["args"]
    args.from_previous_checkpoint(cmd_args)
    torch.manual_seed(getattr(args, "seed", 1))

    if cmd_args.encrypted_inference:
        hook = sy.TorchHook(torch)
        data_owner = sy.VirtualWorker(hook, id="data_owner")
        crypto_provider = sy.VirtualWorker(hook, id="crypto_provider")
        model_owner = sy.VirtualWorker(hook, id="model_owner")
        if torch.cuda.device_count() >= 3:  # parties on GPU 0 / 1, the provider on GPU 2
            model_owner.device, data_owner.device, crypto_provider.device = (torch.device(f"cuda:{i}") for i in range(3))
        workers = [model_owner, data_owner]
        sy.local_worker.clients = [model_owner, data_owner]

    resolution = getattr(args, "inference_resolution", getattr(args, "train_resolution", 224))
    num_classes = 3
    val_mean_std = state["val_mean_std"] if "val_mean_std" in state else (torch.zeros(3), torch.ones(3))
    mean, std = val_mean_std
    g = torch.Generator().manual_seed(getattr(args, "seed", 1))
    raw = torch.randn(cmd_args.num_images, 3, resolution, resolution, generator=g)
    data = (raw - mean.view(1, -1, 1, 1)) / std.view(1, -1, 1, 1)  # a.Normalize(mean, std) of the loader (inference.py:193-199)
    if cmd_args.encrypted_inference:
        data.tag("#inference_data")
        data_owner.load_data([data.send(data_owner).get()])
        grid = sy.PrivateGridNetwork(data_owner, crypto_provider, model_owner)
        data_tensor = grid.search("#inference_data")["data_owner"][0]
        dataset = RemoteTensorDataset(data_tensor)
    else:
        dataset = [d for d in data]

    if getattr(args, "model", "resnet-18") != "resnet-18":
        raise ValueError("Model name not recognised / out of scope: only 'resnet-18' is built.")
    model = resnet18(pretrained=False, num_classes=num_classes, in_channels=3, adptpool=False, input_size=resolution,
                     pooling=args.pooling_type if hasattr(args, "pooling_type") else "avg")
    model.load_state_dict(state["model_state_dict"])
    model.to(device)
    fix_prec_kwargs = {"precision_fractional": cmd_args.precision_fractional, "dtype": "long"}
    graph = None
    if cmd_args.encrypted_inference:
        share_kwargs = {"crypto_provider": crypto_provider, "protocol": "fss", "requires_grad": False}
        model.fix_precision(**fix_prec_kwargs).share(*workers, **share_kwargs)
    model.eval()
    model.pool, model.relu = model.relu, model.pool
    if cmd_args.encrypted_inference and cmd_args.cuda_graph:
        from primia_b200.ring import EncryptedResNet18
        from primia_b200.ring.resnet import EncryptedInferenceGraph

        net = EncryptedResNet18.from_state_dict(state["model_state_dict"], workers, sy.make_crypto_provider(crypto_provider), 10,
                                                cmd_args.precision_fractional, input_size=resolution)
        graph = EncryptedInferenceGraph(net, dataset[0].get().unsqueeze(0))
    total_pred = []
    with torch.no_grad():
        for i, data in enumerate(dataset):
            while len(data.shape) < 4:
                data = data.unsqueeze(0)
            data = data.to(device)
            if graph is not None:
                graph.offline()
                output, _ = graph.online(data.get())
            else:
                if cmd_args.encrypted_inference:
                    data = data.fix_precision(**fix_prec_kwargs).share(*workers, **share_kwargs).get()
                output = model(data)
                if cmd_args.encrypted_inference:
                    output = output.get().float_prec()
            pred = output.argmax(dim=1)
            total_pred.append(pred.detach().cpu().item())
            main.last_logits = output.detach().float().cpu().clone()
    pred_dict = {"Inference Results": dict(enumerate(total_pred))}
    sys.stdout.write(json.dumps(pred_dict))
    print("\n{:s}".format(str(Counter(total_pred))))
    tock = time.time()
    print()
    print(f"Took {tock-tick} seconds.")  # inference.py:326-328
    return total_pred


if __name__ == "__main__":
    main()
