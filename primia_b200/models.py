"""``resnet18`` with the reference's constructor (torchlib/models.py:345-516) -- the model object train.py / inference.py hold.

The module is a plain ``torch.nn.Module`` whose ``state_dict()`` is key- and layout-compatible with the reference's
(torchvision names, KCRS conv weights), so checkpoints interchange (torchlib/utils.py:1470-1493).  It owns no arithmetic:
``forward`` dispatches on the input

  * a CUDA float tensor  -> ``ResNet18Engine`` (eval-mode forward through the C ABI, primia_b200/train/resnet18.py);
  * a ``FixedPrecisionTensor > AdditiveSharingTensor`` after ``model.fix_precision().share(...)`` (primia_b200.sy)
                         -> ``EncryptedResNet18`` (the SPDZ forward on shares, primia_b200/ring/resnet.py),

exactly the two ways inference.py:314 calls ``model(data)``.  Training goes through ``ResNet18Engine`` directly
(primia_b200/train/federated.py); ``engine_for`` builds one from this module's weights.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _Tag(nn.Module):
    """parameter-free placeholder for relu / pool / avgpool: inference.py:289 swaps ``model.pool`` and ``model.relu``, and the
    forward follows whatever order the two attributes are in"""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind

    def extra_repr(self):
        return self.kind


class BasicBlock(nn.Module):
    """torchlib/models.py:238-284 (parameters only)"""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class ResNet(nn.Module):
    """ResNet(BasicBlock, [2,2,2,2]) -- torchlib/models.py:345-485"""

    def __init__(self, num_classes=1000, in_channels=3, adptpool=True, input_size=224, pooling="avg"):
        super().__init__()
        if pooling not in ("max", "avg"):
            raise NotImplementedError("pooling type unknown: {:s}".format(str(pooling)))  # models.py:388-389
        self.inplanes = 64
        self.in_channels, self.input_size, self.pooling, self.adptpool = in_channels, input_size, pooling, adptpool
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = _Tag("relu")
        self.pool = _Tag(pooling + "pool")
        self.layer1 = self._make_layer(64, 2, 1)
        self.layer2 = self._make_layer(128, 2, 2)
        self.layer3 = self._make_layer(256, 2, 2)
        self.layer4 = self._make_layer(512, 2, 2)
        self.avgpool = _Tag("adaptiveavgpool" if adptpool else "avgpool%d" % int(input_size / 32))
        self.fc = nn.Linear(512, 1000)
        for m in self.modules():  # models.py:408-413
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._engines = {}
        self._encrypted = None

    def _make_layer(self, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride, bias=False), nn.BatchNorm2d(planes))
        layers = [BasicBlock(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes
        for _ in range(1, blocks):
            layers.append(BasicBlock(planes, planes))
        return nn.Sequential(*layers)

    # ------------------------------------------------------------------ dispatch
    def stem_order(self):
        """("relu","pool") for the training order conv-bn-relu-pool (models.py:468-471); ("pool","relu") once inference.py:289 has
        swapped the two attributes"""
        first = "pool" if self.relu.kind.endswith("pool") else "relu"
        return (first, "relu" if first == "pool" else "pool")

    def engine_for(self, batch, device=None, mode="f32", **kw):
        """a ResNet18Engine holding this module's weights (training or plain inference)"""
        from .train import ResNet18Engine

        device = str(device or next(self.parameters()).device)
        if not device.startswith("cuda"):
            device = "cuda:0"
        key = (batch, device, mode)
        eng = self._engines.get(key)
        if eng is None:
            eng = ResNet18Engine(batch, self.fc.out_features, self.in_channels, self.input_size, self.pooling, device, mode,
                                 adptpool=self.adptpool, **kw)
            self._engines[key] = eng
        eng.load_state_dict(self.state_dict())
        return eng

    def forward(self, x):
        from .ring.tensors import FixedPrecisionTensor

        if isinstance(x, FixedPrecisionTensor):
            shared = getattr(self, "_sy_shared", None)
            if shared is None or getattr(self, "_sy_state", None) != "shared":
                raise RuntimeError("model(data) on shares needs model.fix_precision(...).share(...) first (inference.py:280-286)")
            if self.stem_order() != ("pool", "relu"):
                raise NotImplementedError("the encrypted forward is inference.py's: swap model.pool and model.relu first "
                                          "(inference.py:289)")
            if self._encrypted is None:
                from .ring.resnet import EncryptedResNet18

                ast = x.child
                self._encrypted = EncryptedResNet18(shared, ast.parties, ast.provider, x.base, x.precision_fractional,
                                                    self.input_size, ast.rng)
            return self._encrypted(x)
        if not (torch.is_tensor(x) and x.is_cuda):
            raise RuntimeError("primia_b200 models run on CUDA tensors or on secret shares (there is no CPU path)")
        if self.training:
            raise RuntimeError("training steps go through primia_b200.train (ResNet18Engine / federated_round); "
                               "model(x) is the eval-mode forward of inference.py")
        if self.stem_order() != ("relu", "pool") and self.pooling != "max":
            raise NotImplementedError("pool before relu only commutes for max pooling")
        # max-pool and ReLU commute (both monotone), so the swapped order of inference.py:289 gives the same plain forward
        eng = self.engine_for(x.shape[0], x.device, "f32")
        eng.training = False
        eng.forward(x.float().contiguous())
        return eng.logits_only().clone()


def resnet18(pretrained=False, progress=True, in_channels=3, pooling="avg", num_classes=1000, **kwargs):
    """torchlib/models.py:487-516: ``_resnet`` builds the 1000-class net, (optionally loads ImageNet weights), then replaces
    ``fc`` by ``Linear(512, num_classes)``.  ``pretrained=True`` needs a download and is refused (no network)."""
    if pretrained:
        raise NotImplementedError("pretrained ImageNet weights need a download; load a checkpoint with load_state_dict instead")
    model = ResNet(in_channels=in_channels, pooling=pooling, **kwargs)
    model.fc = nn.Linear(512, num_classes)
    return model
