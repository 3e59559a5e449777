"""Pin the path-T oracle against fixtures produced by importing the reference's torchlib/models.py."""
import os

import numpy as np
import torch

from oracle import train_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_resnet18_restatement_matches_reference_models_py():
    g = np.load(os.path.join(GOLDEN, "train_ref.npz"))
    torch.manual_seed(42)
    m = O.ResNet18(num_classes=3, in_channels=3, adptpool=False, input_size=64, pooling="max")
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [",".join(map(str, v.shape)) for v in sd.values()] == list(g["shapes"])
    # identical RNG consumption order => identical init
    np.testing.assert_allclose(np.array([v.double().sum().item() for v in sd.values()]), g["weight_sums"], rtol=0, atol=0)
    gen = torch.Generator().manual_seed(42)
    x = torch.randn(4, 3, 64, 64, generator=gen)
    y = torch.randint(0, 3, (4,), generator=gen)
    m.train()
    out = m(x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-6, atol=1e-6)
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    gn = np.array([p.grad.double().norm().item() for p in m.parameters()])
    np.testing.assert_allclose(gn, g["grad_norms"], rtol=1e-5)
    np.testing.assert_allclose(m.bn1.running_mean.numpy(), g["rm"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(m.bn1.running_var.numpy(), g["rv"], rtol=1e-6, atol=1e-7)


def test_param_counts_match_survey():
    m = O.resnet18()
    assert sum(p.numel() for p in m.parameters()) == 11178051
    assert len(list(m.parameters())) == 62
    n_state = sum(v.numel() for k, v in m.state_dict().items() if "num_batches_tracked" not in k)
    assert n_state == 11187651


def test_aggregation_mean_and_weighted():
    torch.manual_seed(0)
    ms = {w: O.ResNet18(input_size=32) for w in "ab"}
    local = O.ResNet18(input_size=32)
    O.aggregation(local, ms, ["a", "b"])
    for k, v in local.state_dict().items():
        if "num_batches_tracked" in k:
            continue
        torch.testing.assert_close(v, (ms["a"].state_dict()[k] + ms["b"].state_dict()[k]) / 2)
    O.aggregation(local, ms, ["a", "b"], weights={"a": 0.25, "b": 0.75})
    k = "layer1.0.bn1.running_var"
    torch.testing.assert_close(local.state_dict()[k], ms["a"].state_dict()[k] * 0.25 + ms["b"].state_dict()[k] * 0.75)


def test_cross_entropy_one_hot_equals_ce_on_hard_labels():
    torch.manual_seed(0)
    o = torch.randn(5, 3)
    y = torch.randint(0, 3, (5,))
    oh = torch.nn.functional.one_hot(y, 3).float()
    torch.testing.assert_close(O.cross_entropy_one_hot(o, oh), torch.nn.functional.cross_entropy(o, y))
