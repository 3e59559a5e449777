"""CPU oracle for path T (per-hospital ResNet-18 step + FedAvg) -- TEST INFRASTRUCTURE ONLY.

torch-CPU fp32 restatement of the reference's training hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.
Citations are relative to ``/root/reference``.

Parity pin: ``tests/golden/make_golden.py`` imports the *reference's own*
``torchlib/models.py`` (with a stub ``syft`` module: the file's only syft use is an unused
``from syft import Plan`` at models.py:4) in the build container and stores the state_dict
key/shape list, and logits/loss/gradient digests for a seeded input; ``tests/test_oracle_train.py``
checks this restatement against those fixtures.  All float arithmetic of the reference
is PyTorch ATen CPU (third-party, torch 1.4 pinned in environment_torch.yml:98); torch 2.11
CPU is the stand-in.
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- T3
class BasicBlock(nn.Module):
    """BasicBlock -- torchlib/models.py:238-284."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)  # conv3x3 models.py:219-230
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        out = out + identity
        return self.relu(out)


class ResNet18(nn.Module):
    """ResNet(BasicBlock,[2,2,2,2]) -- torchlib/models.py:345-485, resnet18 :499-516, _resnet :487-496.

    fixed ``AvgPool2d(int(input_size/32))`` when ``adptpool=False`` (models.py:400-404);
    first pool max|avg 3x3 s2 p1 (models.py:384-389); fc replaced by Linear(512,num_classes)
    *after* the init loop (models.py:495) -- so fc keeps nn.Linear's default init.
    """

    def __init__(self, num_classes=3, in_channels=3, adptpool=False, input_size=224, pooling="max"):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        if pooling == "max":
            self.pool = nn.MaxPool2d(3, 2, 1)
        elif pooling == "avg":
            self.pool = nn.AvgPool2d(3, 2, 1)
        else:
            raise NotImplementedError(pooling)
        self.layer1 = self._make_layer(64, 2, 1)
        self.layer2 = self._make_layer(128, 2, 2)
        self.layer3 = self._make_layer(256, 2, 2)
        self.layer4 = self._make_layer(512, 2, 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1)) if adptpool else nn.AvgPool2d(int(input_size / 32))
        self.fc = nn.Linear(512, 1000)
        for m in self.modules():  # models.py:408-413
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.fc = nn.Linear(512, num_classes)  # models.py:495

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(
                nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes)
            )
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(planes, planes))
        return nn.Sequential(*layers)

    def forward(self, x):  # models.py:466-482
        x = self.pool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.avgpool(x), 1)
        return self.fc(x)


def resnet18(seed=42, **kw):
    torch.manual_seed(seed)
    return ResNet18(**kw)


# --------------------------------------------------------------------------- T4
def cross_entropy_one_hot(output, target, weight=None):
    """Cross_entropy_one_hot(reduction="mean") -- torchlib/utils.py:404-441."""
    w = torch.sum(weight * target, dim=1) if weight is not None else 1.0
    return torch.mean(w * torch.sum(-target * F.log_softmax(output, dim=1), dim=1))


def make_loss(class_weights=None, soft=False):
    if soft:
        return lambda o, t: cross_entropy_one_hot(o, t, class_weights)
    return nn.CrossEntropyLoss(weight=class_weights, reduction="mean")  # train.py:335-340


# --------------------------------------------------------------------------- T5
def make_optimizer(model, name="Adam", lr=1e-4, weight_decay=5e-4, betas=(0.5, 0.99)):
    """train.py:280-303 / utils.py:1131-1145 (kwargs from configs/torch/pneumonia-resnet-pretrained.ini)."""
    if name == "Adam":
        return torch.optim.Adam(model.parameters(), lr=lr, weight_decay=weight_decay, betas=betas)
    if name == "SGD":
        return torch.optim.SGD(model.parameters(), lr=lr, weight_decay=weight_decay)
    raise NotImplementedError("only Adam or SGD supported.")


# --------------------------------------------------------------------------- T2
def local_step(model, optimizer, loss_fn, data, target):
    """utils.py:1168-1174: zero_grad; pred; loss; backward; step; loss.item()."""
    optimizer.zero_grad()
    pred = model(data)
    loss = loss_fn(pred, target)
    loss.backward()
    optimizer.step()
    return loss.detach().item()


# --------------------------------------------------------------------------- T6
def aggregation(local_model, models, worker_ids, weights=None):
    """aggregation(secure=False) -- utils.py:1000-1092: per key (skipping num_batches_tracked)
    sum_w (w_i *) theta_i, / n if unweighted; load into local_model."""
    fresh = OrderedDict()
    for key in local_model.state_dict().keys():
        if "num_batches_tracked" in key:
            continue
        plist = [
            models[w].state_dict()[key].data.clone() * (weights[w] if weights else 1) for w in worker_ids
        ]
        s = torch.sum(torch.stack(plist), dim=0)
        fresh[key] = s if weights else s / len(worker_ids)
    local_model.load_state_dict(fresh, strict=False)
    return local_model


def send_new_models(local_model, models, worker_ids):
    """send_new_models -- utils.py:1095-1105."""
    for w in worker_ids:
        models[w].load_state_dict(local_model.state_dict())
    return models


def secure_aggregation_value(tensors, weights, base, pf):
    """secure branch of aggregation for ONE key -- utils.py:1045-1060,1078-1085:
    each worker's (param * w).fix_prec(pf) is shared, shares are summed, reconstructed, decoded.
    Reconstruction of a sum of sharings == sum of the encoded values (mod 2**64), so the
    random shares cancel; only encode/sum/decode arithmetic matters."""
    from oracle.ring_oracle import encode, decode

    acc = None
    for t, w in zip(tensors, weights):
        q = encode((t * w).float() if w != 1 else t.float(), base, pf)
        acc = q if acc is None else acc + q
    return decode(acc, base, pf)


# --------------------------------------------------------------------------- T1
def federated_round(models, local_model, optimizers, loss_fn, batches, worker_ids, sync_every_n_batch=1,
                    weights=None, opt_kwargs=None, keep_optim_dict=False):
    """secure_aggregation_epoch -- utils.py:1108-1233 (unencrypted aggregation), restated.

    batches[w] is a list of (data, target); hospitals are visited sequentially (utils.py:1160).
    Returns mean loss.
    """
    opt_kwargs = opt_kwargs or {}
    if not keep_optim_dict:  # utils.py:1131-1145
        for w in worker_ids:
            optimizers[w] = make_optimizer(models[w], **opt_kwargs)
    losses = []
    nb = {w: len(batches[w]) for w in worker_ids}
    for batch_idx in range(max(nb.values())):
        for w in worker_ids:
            if batch_idx >= nb[w]:
                continue
            d, t = batches[w][batch_idx]
            losses.append(local_step(models[w], optimizers[w], loss_fn, d, t))
        if batch_idx > 0 and batch_idx % sync_every_n_batch == 0:  # utils.py:1175
            aggregation(local_model, models, worker_ids, weights)
            send_new_models(local_model, models, [w for w in worker_ids if nb[w] > batch_idx])
            if not keep_optim_dict:
                for w in worker_ids:
                    optimizers[w] = make_optimizer(models[w], **opt_kwargs)
    aggregation(local_model, models, worker_ids, weights)  # utils.py:1220-1230
    send_new_models(local_model, models, worker_ids)
    return sum(losses) / len(losses)


def clone_model(m):
    return copy.deepcopy(m)
