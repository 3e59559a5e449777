"""The names train.py / inference.py import from the reference's ``torchlib/utils.py`` -- configuration, learning-rate
schedule, MixUp and the checkpoint contract -- on top of primia_b200.  Citations are into the reference checkout.

``Arguments`` lives at this import path on purpose: the reference pickles the object itself into every checkpoint
(utils.py:1470-1493, ``"args": args``), so ``torch.load`` of a reference-written file looks up ``torchlib.utils.Arguments``.
"""
from __future__ import annotations

import math
import os
from random import random

import torch


class Arguments:
    """utils.py:92-254: ini sections [config] [augmentation] [albumentations] [federated] [system] -> attributes.
    Keys the two hot paths never read get the reference's commented-out fallbacks instead of raising, so the compact
    synthetic-data ini works as well as the reference's full files."""

    def __init__(self, cmd_args, config, mode: str = "train", verbose: bool = True):
        assert mode in ["train", "inference"], "no other mode known"
        g = lambda sec, key, conv, fb: (conv(config.get(sec, key)) if config.has_option(sec, key) else fb)
        yes = lambda v: str(v).strip().lower() in ("yes", "true", "1", "on")
        self.name = getattr(cmd_args, "training_name", None) or "default"
        self.save_file = getattr(cmd_args, "save_file", "model_weights/completed_trainings.csv")
        self.batch_size = g("config", "batch_size", int, 1)
        self.test_batch_size = g("config", "test_batch_size", int, 1)
        self.train_resolution = g("config", "train_resolution", int, 224)
        self.inference_resolution = g("config", "inference_resolution", int, self.train_resolution)
        self.validation_split = g("config", "validation_split", int, 10)
        self.epochs = g("config", "epochs", int, 1)
        self.lr = g("config", "lr", float, 1e-3)
        self.end_lr = g("config", "end_lr", float, self.lr)
        self.deterministic = g("config", "deterministic", yes, False)
        self.restarts = g("config", "restarts", int, 0)
        self.seed = g("config", "seed", int, 1)
        self.test_interval = g("config", "test_interval", int, 1)
        self.log_interval = g("config", "log_interval", int, 10)
        self.optimizer = g("config", "optimizer", str, "SGD")
        self.differentially_private = g("config", "differentially_private", yes, False)
        assert self.optimizer in ["SGD", "Adam"], "Unknown optimizer"
        if self.optimizer == "Adam":
            self.beta1 = g("config", "beta1", float, 0.9)
            self.beta2 = g("config", "beta2", float, 0.999)
        self.model = g("config", "model", str, "resnet-18")
        assert self.model in ["simpleconv", "resnet-18", "vgg16"]
        self.pooling_type = g("config", "pooling_type", str, "max")
        self.pretrained = g("config", "pretrained", yes, False)
        self.weight_decay = g("config", "weight_decay", float, 0.0)
        self.weight_classes = g("config", "weight_classes", yes, False)
        self.mixup = g("augmentation", "mixup", yes, False)
        self.mixup_prob = g("augmentation", "mixup_prob", float, None)
        self.mixup_lambda = g("augmentation", "mixup_lambda", float, None)
        if self.mixup and self.mixup_prob == 1.0:
            self.batch_size *= 2
            print("Doubled batch size because of mixup")
        self.clahe = g("albumentations", "clahe", yes, False)
        self.train_federated = cmd_args.train_federated if mode == "train" else False
        self.unencrypted_aggregation = cmd_args.unencrypted_aggregation if mode == "train" else False
        if self.train_federated:
            self.sync_every_n_batch = g("federated", "sync_every_n_batch", int, 10)
            self.wait_interval = g("federated", "wait_interval", float, 0.1)
            self.keep_optim_dict = g("federated", "keep_optim_dict", yes, False)
            self.repetitions_dataset = g("federated", "repetitions_dataset", int, 1)
            if self.repetitions_dataset > 1:
                self.epochs = int(self.epochs / self.repetitions_dataset)
                if verbose:
                    print("Number of epochs was decreased to {:d} because of {:d} repetitions of dataset".format(
                        self.epochs, self.repetitions_dataset))
            self.weighted_averaging = g("federated", "weighted_averaging", yes, False)
            self.precision_fractional = g("federated", "precision_fractional", float, 16)
        self.visdom = False
        self.encrypted_inference = getattr(cmd_args, "encrypted_inference", False) if mode == "inference" else False
        self.data_dir = getattr(cmd_args, "data_dir", None)
        self.cuda = getattr(cmd_args, "cuda", True)
        self.websockets = False
        self.num_threads = g("system", "num_threads", int, 0)

    @classmethod
    def from_namespace(cls, args):
        """utils.py:256-268 (a checkpoint may hold an argparse Namespace instead)"""
        obj = cls.__new__(cls)
        for attr in dir(args):
            if not callable(getattr(args, attr)) and not attr.startswith("__"):
                setattr(obj, attr, getattr(args, attr))
        return obj

    @classmethod
    def from_dict(cls, d):
        """round-1 checkpoints of this repo stored ``vars(args)``"""
        obj = cls.__new__(cls)
        obj.__dict__.update(d)
        return obj

    def from_previous_checkpoint(self, cmd_args):
        """utils.py:270-281"""
        self.visdom = False
        if hasattr(cmd_args, "encrypted_inference"):
            self.encrypted_inference = cmd_args.encrypted_inference
        self.cuda = cmd_args.cuda
        self.websockets = False
        if "mixup" not in dir(self):
            self.mixup = False

    def incorporate_cmd_args(self, cmd_args):
        """utils.py:283-292"""
        for attr in dir(self):
            if not callable(getattr(self, attr)) and not attr.startswith("__") and attr in dir(cmd_args):
                setattr(self, attr, getattr(cmd_args, attr))

    def __str__(self):
        members = [a for a in dir(self) if not callable(getattr(self, a)) and not a.startswith("__")]
        try:
            from tabulate import tabulate

            return tabulate([[str(x), str(getattr(self, x))] for x in members])
        except ImportError:
            return "\n".join(f"{x}\t{getattr(self, x)}" for x in members)


class LearningRateScheduler:
    """utils.py:37-89: ``lr(epoch) = 10 ** ((log_end - log_start) / total_epochs * epoch + log_start)`` (log_linear) or the
    log-cosine variant, ``epoch`` taken modulo the restart period.  NB the interpolation divides by ``total_epochs``, so
    ``end_lr`` itself is never reached (epoch runs 0 .. total_epochs - 1, train.py:433-440)."""

    def __init__(self, total_epochs, log_start_lr, log_end_lr, schedule_plan="log_linear", restarts=None):
        if restarts == 0:
            restarts = None
        self.total_epochs = total_epochs if not restarts else total_epochs / (restarts + 1)
        if schedule_plan == "log_linear":
            self.calc_lr = lambda epoch: math.pow(10, ((log_end_lr - log_start_lr) / self.total_epochs) * epoch + log_start_lr)
        elif schedule_plan == "log_cosine":
            self.calc_lr = lambda epoch: math.pow(
                10, (math.cos(math.pi * (epoch / self.total_epochs)) / 2.0 + 0.5) * abs(log_start_lr - log_end_lr) + log_end_lr)
        else:
            raise NotImplementedError("Requested learning rate schedule {} not implemented".format(schedule_plan))

    def get_lr(self, epoch):
        epoch = epoch % self.total_epochs
        return self.calc_lr(epoch)

    def adjust_learning_rate(self, optimizer, epoch):
        """``optimizer``: anything with an ``lr`` attribute (ResNet18Engine) or torch-style ``param_groups``"""
        new_lr = self.get_lr(epoch)
        if hasattr(optimizer, "param_groups"):
            for group in optimizer.param_groups:
                group["lr"] = new_lr
        else:
            optimizer.lr = new_lr
        return new_lr


class MixUp(torch.nn.Module):
    """utils.py:327-400 for batched tensors (the federated loader's use, utils.py:1264-1267): the two halves of a batch are
    blended, ``x = l * x[:h] + (1 - l) * x[h:]`` and likewise the one-hot targets; an odd batch keeps its last sample.  Runs on
    whatever device the batch lives on (the owner's GPU)."""

    def __init__(self, λ=None, p=None):
        super().__init__()
        assert p is None or 0.0 <= p <= 1.0, "probability needs to be in [0,1]"
        self.p = p
        if λ:
            assert 0.0 <= λ <= 1.0, "mix factor needs to be in [0,1]"
        self.λ = λ

    def forward(self, x):
        assert len(x) == 2, "need data and target"
        x, y = x
        if self.p:
            if random() > self.p:
                return x, y
        L = x.shape[0]
        if not (torch.is_tensor(y) and y.shape[0] == L):
            raise ValueError("targets need to be tuple of equally shaped one hot encoded tensors")
        if L == 1:
            return x, y
        lam = self.λ if self.λ else random()
        if L % 2 == 0:
            h = L // 2
            return lam * x[:h] + (1.0 - lam) * x[h:], lam * y[:h] + (1.0 - lam) * y[h:]
        h = (L - 1) // 2
        out_x = torch.zeros((h + 1, *x.shape[1:]), device=x.device, dtype=x.dtype)
        out_y = torch.zeros((h + 1, *y.shape[1:]), device=y.device, dtype=y.dtype)
        out_x[-1], out_y[-1] = x[-1], y[-1]
        out_x[:-1] = lam * x[:h] + (1.0 - lam) * x[h:-1]
        out_y[:-1] = lam * y[:h] + (1.0 - lam) * y[h:-1]
        return out_x, out_y


# --------------------------------------------------------------------------------------------- checkpoint contract
def optimizer_state_dict(engine):
    """torch.optim.{Adam,SGD}.state_dict() of the hospital's optimizer (what utils.py:1471-1472 stores per worker):
    ``{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [{lr, betas, eps, weight_decay, ..., "params": [..]}]}``
    with per-parameter tensors in the reference's torch layout (KCRS conv weights)."""
    names = [n for n, _ in engine.param_order]
    group = {"lr": engine.lr, "weight_decay": engine.wd, "params": list(range(len(names)))}
    state = {}
    if engine.opt_name == "Adam":
        group.update({"betas": tuple(engine.betas), "eps": engine.opt_eps, "amsgrad": False})
        if engine.step_count > 0:
            m, v = engine.flat_to_torch_layout(engine.adam_m), engine.flat_to_torch_layout(engine.adam_v)  # real after >= 1 step
            for i, n in enumerate(names):
                state[i] = {"step": engine.step_count, "exp_avg": m[n].cpu(), "exp_avg_sq": v[n].cpu()}
    else:
        group.update({"momentum": 0, "dampening": 0, "nesterov": False})
    return {"state": state, "param_groups": [group]}


def load_optimizer_state_dict(engine, sd):
    """inverse of ``optimizer_state_dict`` (train.py:344-389 ``optimizer[w].load_state_dict``)"""
    group = sd["param_groups"][0]
    engine.lr = float(group["lr"])
    engine.wd = float(group.get("weight_decay", engine.wd))
    if "betas" in group:
        engine.betas = tuple(float(b) for b in group["betas"])
        engine.opt_eps = float(group.get("eps", engine.opt_eps))
    names = [n for n, _ in engine.param_order]
    if sd["state"]:
        steps = {int(s["step"]) for s in sd["state"].values()}
        assert len(steps) == 1, "per-parameter step counts differ"
        engine.torch_layout_to_flat({n: sd["state"][i]["exp_avg"] for i, n in enumerate(names)}, engine.adam_m)
        engine.torch_layout_to_flat({n: sd["state"][i]["exp_avg_sq"] for i, n in enumerate(names)}, engine.adam_v)
        engine.step_count = steps.pop()
        engine._mv_zero = False
    else:
        engine.reset_optimizer()


def save_model(model, optim, path, args, epoch, val_mean_std):
    """utils.py:1470-1493, same keys and nesting.  ``model``: {"local_model": module-or-engine, ...} when federated, else one
    module/engine; ``optim``: {worker: engine} when federated (the engine owns the optimizer state), else one engine."""
    as_sd = lambda m: {k: v.detach().cpu() for k, v in m.state_dict().items()}
    if args.train_federated:
        opt_state_dict = {key: optimizer_state_dict(o) for key, o in optim.items()}
        model_sd = as_sd(model["local_model"])
    else:
        opt_state_dict = optimizer_state_dict(optim)
        model_sd = as_sd(model)
    dirpath = os.path.split(path)[0]
    if dirpath and not os.path.isdir(dirpath):
        os.makedirs(dirpath)
    torch.save({"epoch": epoch, "model_state_dict": model_sd, "optim_state_dict": opt_state_dict, "args": args,
                "val_mean_std": val_mean_std}, path)
    return path


def load_checkpoint(path, map_location="cpu"):
    """torch.load of a checkpoint in the reference's shape; ``args`` normalised to an ``Arguments`` (inference.py:90-93)"""
    from argparse import Namespace

    state = torch.load(path, map_location=map_location, weights_only=False)
    a = state.get("args")
    if isinstance(a, Namespace):
        state["args"] = Arguments.from_namespace(a)
    elif isinstance(a, dict):
        state["args"] = Arguments.from_dict(a)
    return state
