#!/usr/bin/env python
"""train.py -- entry point with the reference's CLI (train.py:555-631): ``python train.py --config <ini> --train_federated``.

Flow mirrors train.py:54-552 + torchlib/utils.py:516-856 (setup_pysyft), :936 (train_federated), :1108
(secure_aggregation_epoch), :1470 (save_model): roster from ``configs/websetting/config.csv`` (the reference's file works as
is), one ``sy.VirtualWorker`` per hospital pinned to one GPU, ``model.copy().send(worker)`` per hospital, per-worker
optimizers, log-linear learning-rate schedule, FedAvg (plain or secure) every ``sync_every_n_batch`` batches, checkpoints in
the reference's format (``--resume_checkpoint`` continues from one).  All arithmetic is the primia_b200 C ABI.  Data are
synthetic 224x224x3 tensors (``--data_dir`` is accepted for CLI compatibility; image I/O / albumentations are out of scope,
SURVEY.md section 2 #4).
Multi-GPU: launch with ``python -m torch.distributed.run --nproc-per-node N train.py ...``: rank i hosts hospital i and
FedAvg is an NCCL all-reduce; in a single process the hospitals time-share the visible GPU(s) round-robin.
"""
from __future__ import annotations

import argparse
import configparser
import math
import os
import sys
import time
import zlib

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from torchlib.run_websocket_server import read_websocket_config  # noqa: E402
from torchlib.utils import (Arguments, LearningRateScheduler, MixUp, load_checkpoint, load_optimizer_state_dict,  # noqa: E402
                            save_model)


def stable_seed(base: int, name: str) -> int:
    """per-hospital data seed: reproducible across processes (python's hash() is salted per interpreter)"""
    return base + zlib.crc32(name.encode()) % 1000


def global_batch_counts(local_counts: dict, group):
    """{hospital: number of batches} over ALL ranks (utils.py:953-957 needs the global denominator for weighted averaging)"""
    if group is None:
        return dict(local_counts)
    import torch.distributed as dist

    gathered = [None] * dist.get_world_size(group)
    dist.all_gather_object(gathered, local_counts, group=group)
    out = {}
    for d in gathered:
        out.update(d)
    return out


def main(argv=None):
    import primia_b200.sy as sy
    from primia_b200.models import resnet18
    from primia_b200.train import HospitalWorker, federated_round

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=os.path.join(ROOT, "configs/torch/pneumonia-resnet-synthetic.ini"))
    ap.add_argument("--train_federated", action="store_true")
    ap.add_argument("--unencrypted_aggregation", action="store_true")
    ap.add_argument("--data_dir", default=None, help="accepted for CLI compatibility; synthetic tensors are used")
    ap.add_argument("--websockets_config", default=os.path.join(ROOT, "configs/websetting/config.csv"))
    ap.add_argument("--websockets", action="store_true", help="network workers are out of scope: refused")
    ap.add_argument("--cuda", action="store_true", help="accepted; the GPU is the only device")
    ap.add_argument("--resume_checkpoint", default=None, help="continue from a checkpoint in the reference's format")
    ap.add_argument("--training_name", default=None)
    ap.add_argument("--mode", default="bf16", choices=["bf16", "f32", "f32x3"])
    ap.add_argument("--batches_per_worker", type=int, default=4)
    ap.add_argument("--save_dir", default="model_weights")
    ap.add_argument("--dp_noise_multiplier", type=float, default=1.3, help="train.py:330 hard-codes 1.3")
    ap.add_argument("--dp_max_grad_norm", type=float, default=1.0, help="train.py:331 hard-codes 1.0")
    cmd = ap.parse_args(argv)
    config = configparser.ConfigParser()
    assert os.path.isfile(cmd.config), "config file not found"
    config.read(cmd.config)
    args = Arguments(cmd, config, mode="train")
    if cmd.websockets:
        raise SystemExit("websocket / PyGrid workers are out of scope (SURVEY.md section 2): use VirtualWorkers")
    if not args.train_federated:
        raise SystemExit("only the federated path (--train_federated) is built: it is the hot path (BASELINE.json north_star)")
    if args.model != "resnet-18":
        raise NotImplementedError("model unknown / out of scope: " + args.model)
    if cmd.unencrypted_aggregation is False and config.has_option("federated", "unencrypted_aggregation"):
        args.unencrypted_aggregation = config.getboolean("federated", "unencrypted_aggregation")

    import torch.distributed as dist

    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    group = None
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        group = dist.group.WORLD
    if args.deterministic:
        torch.manual_seed(args.seed)
    hook = sy.TorchHook(torch)
    # ---- setup_pysyft (utils.py:516-603): roster, crypto provider, VirtualWorkers
    worker_dict = read_websocket_config(cmd.websockets_config)
    worker_names = [d["id"] for _, d in worker_dict.items()]
    crypto_in_config = "crypto_provider" in worker_names
    assert args.unencrypted_aggregation or crypto_in_config, "No crypto provider in configuration"
    if crypto_in_config:
        worker_names.remove("crypto_provider")
    if world > 1:  # one hospital per rank; extra ranks get numbered names
        worker_names = [worker_names[rank % len(worker_names)] + (str(rank) if rank >= len(worker_names) else "")]
    dev_of = (lambda i: f"cuda:{torch.cuda.current_device()}") if world > 1 else (lambda i: None)
    workers = {n: sy.VirtualWorker(hook, id=n, verbose=False, device=dev_of(i)) for i, n in enumerate(worker_names)}
    for w in workers.values():
        w.object_store.clear_objects()
    crypto_provider = None
    if not args.unencrypted_aggregation:
        crypto_provider = sy.VirtualWorker(hook, id="crypto_provider", verbose=False)  # noqa: F841 (utils.py:598-601)

    # ---- synthetic per-hospital data, tagged and sent to the owner like utils.py:643-742
    B, S = args.batch_size, args.train_resolution
    num_classes = 3
    one_hot = args.mixup or args.weight_classes          # utils.py:671 / train.py:336: soft targets -> Cross_entropy_one_hot
    for n, w in workers.items():
        g = torch.Generator().manual_seed(stable_seed(args.seed, n))
        n_img = B * cmd.batches_per_worker
        data = torch.randn(n_img, 3, S, S, generator=g).tag("#traindata")
        target = torch.randint(0, num_classes, (n_img,), generator=g)
        if one_hot:
            target = torch.nn.functional.one_hot(target, num_classes).float()
        target = target.tag("#traintargets")
        w.load_data([data.send(w).get(), target.send(w).get()])
    grid = sy.PrivateGridNetwork(*workers.values())
    data_ptrs, target_ptrs = grid.search("#traindata"), grid.search("#traintargets")
    loaders = {}
    for n in workers:
        ds = sy.BaseDataset(data_ptrs[n][0], target_ptrs[n][0])
        loaders[n] = sy.FederatedDataLoader(sy.FederatedDataset([ds]), batch_size=B, shuffle=True, seed=args.seed)

    # ---- one model + optimizer per hospital (train.py:262-303), all starting from the same local_model
    model = resnet18(pretrained=False, num_classes=num_classes, in_channels=3, adptpool=False, input_size=S,
                     pooling=args.pooling_type)
    model = {w: model.copy().send(workers[w]) for w in worker_names} | {"local_model": model}
    opt_kw = dict(optimizer=args.optimizer, lr=args.lr, weight_decay=args.weight_decay)
    if args.optimizer == "Adam":
        opt_kw["betas"] = (args.beta1, args.beta2)
    class_weights = None
    if args.weight_classes:  # calc_class_weights (utils.py:470-513): inverse class frequency over all hospitals' targets
        counts = sum(t.get().sum(0) for ts in target_ptrs.values() for t in ts)
        if group is not None:
            dist.all_reduce(counts, group=group)
        class_weights = (1.0 / counts.clamp_min(1)).float()
        class_weights = class_weights / class_weights.sum()
    hospitals = []
    for n in worker_names:
        eng = model[n].engine_for(B if not (args.mixup and args.mixup_prob == 1.0) else B // 2, workers[n].device, cmd.mode,
                                  class_weights=class_weights, **opt_kw)
        # train.py:304-334: PrivacyEngine(noise_multiplier=1.3, max_grad_norm=1.0) attached to every hospital's optimizer.  The
        # reference exits here for federated training ("only ... local training and models without BatchNorm"); the build
        # runs the DP step with BatchNorm frozen (primia_b200/train/dp.py)
        dp = {"noise_multiplier": cmd.dp_noise_multiplier, "max_grad_norm": cmd.dp_max_grad_norm} if args.differentially_private else None
        hospitals.append(HospitalWorker(n, eng, dp=dp))
    optimizer = {h.id: h.engine for h in hospitals}  # the engine owns the optimizer state (Adam moments, step count)

    start_at_epoch = 1
    if cmd.resume_checkpoint:  # train.py:344-389
        print("Resume training from a given checkpoint.")
        state = load_checkpoint(cmd.resume_checkpoint)
        start_at_epoch = state["epoch"]
        ck_args = state["args"]
        opt_sd = state["optim_state_dict"]
        if getattr(ck_args, "train_federated", False):
            for w in worker_names:
                if w not in opt_sd:
                    raise SystemExit("The worker names of the checkpoint and the current configuration cannot be matched.")
                load_optimizer_state_dict(optimizer[w], opt_sd[w])
        else:
            assert len(opt_sd) == 2 and "param_groups" in opt_sd and "state" in opt_sd  # non-federated checkpoint
            for w in worker_names:
                load_optimizer_state_dict(optimizer[w], opt_sd)
        for key in model:
            model[key].load_state_dict(state["model_state_dict"])
        for h in hospitals:
            h.engine.load_state_dict(state["model_state_dict"])

    scheduler = LearningRateScheduler(args.epochs, math.log10(args.lr), math.log10(args.end_lr), restarts=args.restarts)
    mixup = MixUp(λ=args.mixup_lambda, p=args.mixup_prob) if args.mixup else None
    counts = global_batch_counts({n: len(l) for n, l in loaders.items()}, group)
    total_batches = sum(counts.values())
    weights = {n: c / total_batches for n, c in counts.items()} if args.weighted_averaging else None  # utils.py:953-957
    last_path = None
    for epoch in range(start_at_epoch, args.epochs + 1):
        for h in hospitals:
            new_lr = scheduler.adjust_learning_rate(h.engine, epoch - 1)  # train.py:433-440
            h.batches = []
            for d, t in loaders[h.id]:
                d, t = d.get(), t.get()
                if mixup is not None:
                    d, t = mixup((d, t))
                if d.shape[0] == h.engine.B:  # the engine's buffers are sized for full batches; a ragged tail is dropped
                    h.batches.append((d.contiguous(), t.contiguous()))
        t0 = time.time()
        loss = federated_round(hospitals, args.sync_every_n_batch, weights, args.keep_optim_dict, group,
                               secure=not args.unencrypted_aggregation, precision_fractional=int(args.precision_fractional),
                               opt_lr=None if args.keep_optim_dict else args.lr)
        torch.cuda.synchronize()
        dt = time.time() - t0
        n_img = sum(len(h.batches) * h.engine.B for h in hospitals) * world
        if rank == 0:
            print("Train Epoch: {} \tLoss: {:.6f}\tlr {:.2e}\t({:.0f} images/s)".format(epoch, loss.item(), new_lr, n_img / dt))
        model["local_model"].load_state_dict(hospitals[0].engine.state_dict())  # every hospital holds the aggregate
        if rank == 0 and (epoch % args.test_interval == 0 or epoch == args.epochs):
            last_path = os.path.join(cmd.save_dir, f"federated_resnet18_epoch_{epoch:03d}.pt")
            save_model(model, optimizer, last_path, args, epoch, torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]))
            print("saved", last_path)
    if world > 1:
        dist.destroy_process_group()
    main.last_checkpoint = last_path
    return hospitals


if __name__ == "__main__":
    main()
