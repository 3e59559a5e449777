#!/usr/bin/env python
"""train.py -- entry point with the reference's CLI (train.py:555-631): ``python train.py --config <ini> --train_federated``.

Flow mirrors train.py:54-552 + torchlib/utils.py:516-856 (setup_pysyft), :936 (train_federated), :1108
(secure_aggregation_epoch), :1354 (test), :1470 (save_model), with every hospital (VirtualWorker) pinned to one GPU and
all arithmetic in the primia_b200 C ABI.  Data are synthetic 224x224x3 tensors (``--data_dir`` is accepted for CLI
compatibility; image I/O / albumentations are out of scope).
Multi-GPU: launch with ``python -m torch.distributed.run --nproc-per-node N train.py ...``: rank i hosts hospital i and
FedAvg is an NCCL all-reduce; in a single process the hospitals time-share the visible GPU(s) round-robin.
"""
from __future__ import annotations

import argparse
import configparser
import csv
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class Arguments:
    """subset of torchlib/utils.py:92-302 that the federated path reads (ini -> attributes)"""

    def __init__(self, cmd_args, config):
        c, f = config["config"], config["federated"] if "federated" in config else {}
        self.batch_size = c.getint("batch_size", 64)
        self.train_resolution = c.getint("train_resolution", 224)
        self.test_batch_size = c.getint("test_batch_size", 64)
        self.test_interval = c.getint("test_interval", 1)
        self.epochs = c.getint("epochs", 1)
        self.lr = c.getfloat("lr", 1e-4)
        self.end_lr = c.getfloat("end_lr", self.lr)
        self.beta1, self.beta2 = c.getfloat("beta1", 0.5), c.getfloat("beta2", 0.99)
        self.weight_decay = c.getfloat("weight_decay", 5e-4)
        self.seed = c.getint("seed", 42)
        self.optimizer = c.get("optimizer", "Adam")
        self.model = c.get("model", "resnet-18")
        self.pooling_type = c.get("pooling_type", "max")
        self.sync_every_n_batch = int(f.get("sync_every_n_batch", 1))
        self.keep_optim_dict = str(f.get("keep_optim_dict", "no")).lower() in ("yes", "true", "1")
        self.weighted_averaging = str(f.get("weighted_averaging", "no")).lower() in ("yes", "true", "1")
        self.precision_fractional = int(f.get("precision_fractional", 16))
        self.unencrypted_aggregation = cmd_args.unencrypted_aggregation or str(f.get("unencrypted_aggregation", "yes")).lower() in ("yes", "true", "1")
        self.train_federated = cmd_args.train_federated
        self.mode = cmd_args.mode
        self.batches_per_worker = cmd_args.batches_per_worker
        self.num_classes = 3
        self.in_channels = 3


def read_websocket_config(path):
    """torchlib/run_websocket_server.py:6-8: worker roster = rows of the CSV"""
    with open(path) as fh:
        return [row["id"] for row in csv.DictReader(fh)]


def lr_at(args, epoch):
    """log-linear schedule lr -> end_lr over the epochs (torchlib/utils.py:37-89, no restarts)"""
    if args.epochs <= 1:
        return args.lr
    import math

    t = (epoch - 1) / (args.epochs - 1)
    return math.exp(math.log(args.lr) + t * (math.log(args.end_lr) - math.log(args.lr)))


def main(argv=None):
    import primia_b200.sy as sy
    from primia_b200.train import HospitalWorker, ResNet18Engine, aggregation, federated_round

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(ROOT, "configs/torch/pneumonia-resnet-synthetic.ini"))
    ap.add_argument("--train_federated", action="store_true")
    ap.add_argument("--unencrypted_aggregation", action="store_true")
    ap.add_argument("--data_dir", default=None, help="accepted for CLI compatibility; synthetic tensors are used")
    ap.add_argument("--websockets_config", default=os.path.join(ROOT, "configs/websetting/config.csv"))
    ap.add_argument("--cuda", action="store_true", help="accepted; the GPU is the only device")
    ap.add_argument("--mode", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--batches_per_worker", type=int, default=4)
    ap.add_argument("--save_dir", default="model_weights")
    cmd = ap.parse_args(argv)
    config = configparser.ConfigParser()
    assert os.path.isfile(cmd.config), "config file not found"
    config.read(cmd.config)
    args = Arguments(cmd, config)
    if not args.train_federated:
        raise SystemExit("only the federated path (--train_federated) is built: it is the hot path (BASELINE.json north_star)")
    if args.model != "resnet-18":
        raise NotImplementedError("model unknown / out of scope: " + args.model)

    import torch.distributed as dist

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    group = None
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        group = dist.group.WORLD
    hook = sy.TorchHook(torch)
    names = read_websocket_config(cmd.websockets_config)
    if world > 1:
        names = [names[rank % len(names)] + (str(rank) if rank >= len(names) else "")]
    workers = {n: sy.VirtualWorker(hook, id=n, device=f"cuda:{torch.cuda.current_device()}" if world > 1 else None) for n in names}
    crypto_provider = sy.VirtualWorker(hook, id="crypto_provider")  # noqa: F841 (utils.py:603)

    # synthetic per-hospital data, tagged and "sent" to the owner like utils.py:643-742
    B, S = args.batch_size, args.train_resolution
    loaders = {}
    for n, w in workers.items():
        g = torch.Generator().manual_seed(args.seed + hash(n) % 1000)
        n_img = B * args.batches_per_worker
        data = torch.randn(n_img, 3, S, S, generator=g).tag("#traindata")
        target = torch.randint(0, 3, (n_img,), generator=g).tag("#traintargets")
        w.load_data([data.send(w).get(), target.send(w).get()])
    grid = sy.PrivateGridNetwork(*workers.values())
    data_ptrs, target_ptrs = grid.search("#traindata"), grid.search("#traintargets")
    for n, w in workers.items():
        ds = sy.BaseDataset(data_ptrs[n][0], target_ptrs[n][0])
        loaders[n] = sy.FederatedDataLoader(sy.FederatedDataset([ds]), batch_size=B, shuffle=True, seed=args.seed)

    # one model + optimizer per hospital (train.py:271-303), all starting from the same local_model
    hospitals = []
    for n, w in workers.items():
        eng = ResNet18Engine(B, args.num_classes, args.in_channels, S, args.pooling_type, str(w.device), args.mode,
                             optimizer=args.optimizer, lr=args.lr, betas=(args.beta1, args.beta2), weight_decay=args.weight_decay)
        eng.init_random(seed=args.seed)
        hospitals.append(HospitalWorker(n, eng))

    total_batches = sum(len(l) for l in loaders.values())
    weights = {n: len(l) / total_batches for n, l in loaders.items()} if args.weighted_averaging else None
    for epoch in range(1, args.epochs + 1):
        lr = lr_at(args, epoch)
        for h in hospitals:
            h.engine.lr = lr  # scheduler.adjust_learning_rate per worker, train.py:433-440
            h.batches = [(d.get(), t.get()) for d, t in loaders[h.id]]
        t0 = time.time()
        loss = federated_round(hospitals, args.sync_every_n_batch, weights, args.keep_optim_dict, group,
                               secure=not args.unencrypted_aggregation, precision_fractional=args.precision_fractional)
        torch.cuda.synchronize()
        dt = time.time() - t0
        n_img = sum(len(h.batches) for h in hospitals) * B * world
        if rank == 0:
            print("Train Epoch: {} \tLoss: {:.6f}\t({:.0f} images/s)".format(epoch, loss.item(), n_img / dt))
    if rank == 0:
        save_model(hospitals[0].engine, args, cmd.save_dir, args.epochs)
    if world > 1:
        dist.destroy_process_group()
    return hospitals


def save_model(engine, args, save_dir, epoch):
    """torchlib/utils.py:1470-1493 checkpoint contract (consumed by inference.py:82-93,277)"""
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, f"federated_resnet18_epoch_{epoch:03d}.pt")
    sd = {k: v.cpu() for k, v in engine.state_dict().items()}
    torch.save({"epoch": epoch, "model_state_dict": sd, "optim_state_dict": {}, "args": vars(args),
                "val_mean_std": torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])}, path)
    print("saved", path)
    return path


if __name__ == "__main__":
    main()
