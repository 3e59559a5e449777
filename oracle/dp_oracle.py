"""CPU oracle for the DP-SGD local step (path T, row T9) -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED.  The reference attaches ``torchdp.PrivacyEngine`` (train.py:304-334); its arithmetic lives in the third-party
package ``pytorch-dp == 0.1b1`` (environment_torch.yml:136), whose source is not in the reference tree, and no reference test
or fixture pins it.  This file restates that package's published algorithm in the most literal way -- one backward pass per
sample -- and is what the CUDA path (primia_b200/train/dp.py) is checked against:

  per-sample gradient g_b of the sample's OWN loss (torchdp scales the backprops by the batch size under reduction="mean");
  flat clipping: c_b = clamp(C / (|g_b| + 1e-6), max=1) over the concatenation of all parameters (ConstantFlatClipper);
  p.grad = sum_b c_b g_b / B                        (PerSampleGradientClipper.step)
  p.grad += N(0, (noise_multiplier * C)^2) / B      (PrivacyEngine.step)
  optimizer.step()

BatchNorm: pytorch-dp rejects BatchNorm models and the reference exits for the federated + DP combination
(train.py:306-310).  The build (and this oracle) evaluates a DP step with the BatchNorm layers as frozen per-channel affine
maps (running statistics, i.e. ``model.eval()`` semantics with trainable gamma / beta) -- the standard way to make per-sample
gradients exist for such a network.
"""
from __future__ import annotations

import torch


def per_sample_grads(model, loss_fn, data, target):
    """[{name: grad of sample b's own loss}] with BatchNorm frozen (eval statistics)"""
    was = model.training
    model.eval()
    out = []
    try:
        for b in range(data.shape[0]):
            model.zero_grad()
            loss = loss_fn(model(data[b:b + 1]), target[b:b + 1])
            loss.backward()
            out.append({n: p.grad.detach().clone() for n, p in model.named_parameters()})
    finally:
        model.train(was)
    return out


def dp_step(model, optimizer, loss_fn, data, target, noise, noise_multiplier=1.3, max_grad_norm=1.0):
    """one DP-SGD step; ``noise``: {name: standard-normal tensor} (explicit randomness), scaled here by sigma * C.
    Returns (mean loss with frozen BN, per-sample norms, clip factors)."""
    B = data.shape[0]
    gs = per_sample_grads(model, loss_fn, data, target)
    norms = torch.stack([torch.sqrt(sum((g.double() ** 2).sum() for g in gb.values())) for gb in gs])
    factors = (max_grad_norm / (norms + 1e-6)).clamp(max=1.0)
    was = model.training
    model.eval()
    with torch.no_grad():
        loss = loss_fn(model(data), target).item()
    model.train(was)
    for n, p in model.named_parameters():
        summed = sum(f.float() * gb[n] for f, gb in zip(factors, gs))
        z = noise[n] * (noise_multiplier * max_grad_norm) if noise is not None else 0.0
        p.grad = (summed + z) / B
    optimizer.step()
    return loss, norms, factors
