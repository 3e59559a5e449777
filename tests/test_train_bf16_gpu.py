"""Throughput mode (bf16 activations / tensor-core convs, fp32 accumulation, fp32 master weights).

bf16 has an 8-bit mantissa, so the 1e-5 gate cannot apply (that one runs in fp32 mode, tests/test_train_gpu.py).
The yardstick here is PyTorch's own CPU bf16 autocast of the oracle model: the GPU throughput mode must track the
fp32 oracle at least as well as that standard mixed-precision evaluation does (per-tensor gradient cosine within
0.06 of autocast's per tensor and 0.01 on average, loss within 2%)."""
import copy

import pytest
import torch

from oracle import train_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def test_bf16_step_tracks_fp32_oracle_like_autocast():
    from primia_b200.train import ResNet18Engine

    B, size = 16, 96
    torch.manual_seed(42)
    m = O.ResNet18(input_size=size)
    m_amp = copy.deepcopy(m)
    eng = ResNet18Engine(B, 3, 3, size, "max", DEV, "bf16")
    eng.load_state_dict(m.state_dict())
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, 3, size, size, generator=g)
    y = torch.randint(0, 3, (B,), generator=g)
    m.train()
    loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()
    m_amp.train()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_amp = m_amp(x)
    torch.nn.functional.cross_entropy(out_amp.float(), y).backward()
    eng.forward(x.to(DEV))
    l = eng.loss_and_backward(y.to(DEV))
    torch.cuda.synchronize()
    assert abs(l.item() - loss.item()) / abs(loss.item()) < 0.02
    gd = eng.grad_dict()
    worst = (1.0, None)
    deficits = []
    for (n, p), pa in zip(m.named_parameters(), m_amp.parameters()):
        c_gpu, c_amp = cos(gd[n], p.grad), cos(pa.grad, p.grad)
        worst = min(worst, (c_gpu - c_amp, n))
        deficits.append(c_amp - c_gpu)
        assert c_gpu > c_amp - 0.06, (n, c_gpu, c_amp)
    print("bf16 vs autocast: worst cosine deficit", worst)
    assert sum(deficits) / len(deficits) < 0.01, sum(deficits) / len(deficits)
    eng.optimizer_step()
    assert torch.isfinite(eng.flat).all()
