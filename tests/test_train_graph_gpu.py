"""Parity ON THE BENCHMARKED PATH (path T, bf16 throughput mode as bench.py runs it):

  * the captured CUDA graph of a local step (side-stream weight gradients, programmatic dependent launch, fused stem pool,
    x-recomputed ReLU masks) replays to the same loss / gradients / weights as the eager launch sequence;
  * ``HospitalWorker.local_step_host`` (pinned host batch, copy stream, double-buffered slots) == ``local_step``;
  * PRIMIA_PDL=0 and PRIMIA_PDL=1 give the same step (the attribute only changes launch overlap);
  * the bf16 step at the bench configuration C2 (B = 64 per hospital, 224 x 224) against the torch-CPU fp32 oracle, with the
    measured per-tensor errors printed and FROZEN as the gate (gates below = the errors first measured on B200, x1.5).

Sources of run-to-run difference between two launches of the same kernels: fp32 ``red.add`` order in the split-K weight
gradients and double-precision atomics in the fused BatchNorm statistics.  Nothing else."""
import json
import os
import subprocess
import sys

import pytest
import torch

from oracle import train_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


def _engine(B, size, sd, mode="bf16"):
    from primia_b200.train import ResNet18Engine

    eng = ResNet18Engine(B, 3, 3, size, "max", DEV, mode)
    eng.load_state_dict(sd)
    return eng


def _batches(B, size, n, seed=11):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(B, 3, size, size, generator=g), torch.randint(0, 3, (B,), generator=g)) for _ in range(n)]


def _assert_same_step(a, b, what, lr=1e-4):
    """a, b: (loss, grads flat, weights flat) of two executions of the same step"""
    assert abs(a[0] - b[0]) <= 1e-6 * abs(b[0]), (what, a[0], b[0])
    assert rel(a[1], b[1]) < 1e-4, (what, "grads", rel(a[1], b[1]))
    # Adam's first step moves every weight by ~lr * sign(g): a gradient that differs in its last bits can flip the sign of
    # a ~0 gradient, nothing more
    assert (a[2] - b[2]).abs().max().item() <= 2.0 * lr + 1e-7, (what, (a[2] - b[2]).abs().max().item())
    assert rel(a[2], b[2]) < 1e-5, (what, "weights", rel(a[2], b[2]))


@pytest.mark.parametrize("B,size", [(16, 96), (64, 224)])
def test_graph_replay_equals_eager_step(B, size):
    torch.manual_seed(42)
    sd = O.ResNet18(input_size=size).state_dict()
    batches = _batches(B, size, 3)
    eager, graph = _engine(B, size, sd), _engine(B, size, sd)
    graph.capture_graph(batches[0][0].to(DEV), batches[0][1].to(DEV))
    assert graph._graph is not None and graph._graph["launches"] > 100
    assert torch.equal(eager.flat, graph.flat), "capture must leave the model state untouched"
    for i, (x, y) in enumerate(batches):
        for eng in (eager, graph):
            eng.reset_optimizer()  # the reference re-creates the optimizer after every aggregation (utils.py:1209-1218)
        # every step is compared from IDENTICAL weights: the 1e-7 split-K atomics noise of the previous step's gradients would
        # otherwise flip a few bf16 weight roundings and show up as ~1e-5 in the next loss (measured), which is not a property
        # of graph replay
        graph.flat.copy_(eager.flat)
        le = eager._train_step_eager(x.to(DEV), y.to(DEV)).item()
        lg = graph.train_step(x.to(DEV), y.to(DEV)).item()   # replay: step index and target dtype match the capture
        torch.cuda.synchronize()
        _assert_same_step((le, eager.grads, eager.flat), (lg, graph.grads, graph.flat), f"B={B} size={size} step {i}")
    assert graph.step_count == 1 and eager.step_count == 1


def test_two_graph_overlap_step_equals_eager_step():
    """capture_graph_overlap: graph A (forward, layer4 backward, optimizer on flat[split:]) + graph B (rest of the backward,
    optimizer on flat[:split]) == one eager step; the callbacks fire after A and after B (where bench.py / local_step_and_fedavg
    start the two all-reduces)"""
    B, size = 16, 96
    torch.manual_seed(42)
    sd = O.ResNet18(input_size=size).state_dict()
    batches = _batches(B, size, 2)
    eager, two = _engine(B, size, sd), _engine(B, size, sd)
    two.capture_graph_overlap(batches[0][0].to(DEV), batches[0][1].to(DEV))
    assert torch.equal(eager.flat, two.flat)
    off = two.split_offset
    assert off == two.offsets["layer4.0.conv1.weight"][0] and two.n_flat - off > 3 * off   # 75 % of the state goes first
    for x, y in batches:
        for eng in (eager, two):
            eng.reset_optimizer()
        two.flat.copy_(eager.flat)
        le = eager._train_step_eager(x.to(DEV), y.to(DEV)).item()
        seen = []
        snap = {}

        def after_a():
            seen.append("A")
            snap["tail"] = two.flat[off:two.n_param_flat].clone()   # layer4 + fc already stepped when the first all-reduce starts

        lt = two.train_step_overlapped(x.to(DEV), y.to(DEV), after_a, lambda: seen.append("B")).item()
        torch.cuda.synchronize()
        assert seen == ["A", "B"]
        _assert_same_step((le, eager.grads, eager.flat), (lt, two.grads, two.flat), "two-graph step")
        assert torch.equal(snap["tail"], two.flat[off:two.n_param_flat]), "graph B must not touch the first bucket"


def test_graph_is_not_replayed_when_the_optimizer_step_or_hyperparameters_differ():
    """Adam's bias correction and lr are baked into the captured launches: a replay is only legal at the captured step index
    with the captured hyper-parameters; otherwise the engine must launch eagerly."""
    B, size = 8, 64
    torch.manual_seed(1)
    sd = O.ResNet18(input_size=size).state_dict()
    (x, y), = _batches(B, size, 1)
    x, y = x.to(DEV), y.to(DEV)
    a, b = _engine(B, size, sd), _engine(B, size, sd)
    b.capture_graph(x, y)
    # second step without optimizer reset (keep_optim_dict = yes): step index 2 != captured 1 -> eager
    for eng in (a, b):
        eng.train_step(x, y)
        eng.train_step(x, y)
    torch.cuda.synchronize()
    assert a.step_count == b.step_count == 2
    assert rel(a.flat, b.flat) < 1e-5
    # changed learning rate (train.py:433-440 adjusts it per epoch): the stale graph must not be used
    for eng in (a, b):
        eng.reset_optimizer()
        eng.lr = 3e-3
        eng.train_step(x, y)
    torch.cuda.synchronize()
    assert rel(a.flat, b.flat) < 1e-5, "a graph captured with another lr was replayed"


def test_local_step_host_equals_local_step():
    from primia_b200.train import HospitalWorker

    B, size = 16, 96
    torch.manual_seed(42)
    sd = O.ResNet18(input_size=size).state_dict()
    batches = _batches(B, size, 4)
    dev_w, host_w = HospitalWorker("a", _engine(B, size, sd)), HospitalWorker("b", _engine(B, size, sd))
    host_w.engine.capture_graph(batches[0][0].to(DEV), batches[0][1].to(DEV))   # as bench.py's e2e leg runs it
    dev_w.engine.capture_graph(batches[0][0].to(DEV), batches[0][1].to(DEV))
    pinned = [(x.pin_memory(), y.pin_memory()) for x, y in batches]
    for i, (x, y) in enumerate(batches):
        for w in (dev_w, host_w):
            w.engine.reset_optimizer()
        host_w.engine.flat.copy_(dev_w.engine.flat)   # compare each step from identical weights (see above)
        ld = dev_w.local_step(x.to(DEV), y.to(DEV)).item()
        lh = host_w.local_step_host(*pinned[i])
        if i + 1 < len(batches):
            host_w.prefetch_host(*pinned[i + 1])     # loader look-ahead, overlapping this step
        lh = lh.item()
        torch.cuda.synchronize()
        _assert_same_step((ld, dev_w.engine.grads, dev_w.engine.flat), (lh, host_w.engine.grads, host_w.engine.flat),
                          f"host-fed step {i}")


def test_bf16_staged_batch_gives_the_identical_step():
    """bench.py's e2e leg ships bf16 pixels over PCIe: in bf16 mode the stem rounds every pixel to bf16 before its MMA, so a batch
    staged as bf16 and the same batch shipped as fp32 (already bf16-representable or not) give the bit-identical forward"""
    B, size = 16, 96
    torch.manual_seed(42)
    sd = O.ResNet18(input_size=size).state_dict()
    (x, y), = _batches(B, size, 1)
    a, b = _engine(B, size, sd), _engine(B, size, sd)
    la = a.train_step(x.to(DEV), y.to(DEV)).item()                      # fp32 pixels, rounded inside the stem
    lb = b.train_step(x.bfloat16().to(DEV), y.to(DEV)).item()           # the loader already rounded them
    torch.cuda.synchronize()
    assert la == lb and torch.equal(a.act["conv1"], b.act["conv1"]) and torch.equal(a.logits, b.logits)
    assert rel(b.grads, a.grads) < 1e-5


_WORKER = r"""
import sys, torch
sys.path.insert(0, {root!r})
from oracle import train_oracle as O
from primia_b200.train import ResNet18Engine
B, size = 16, 96
torch.manual_seed(42)
sd = O.ResNet18(input_size=size).state_dict()
g = torch.Generator().manual_seed(11)
x, y = torch.randn(B, 3, size, size, generator=g).cuda(), torch.randint(0, 3, (B,), generator=g).cuda()
eng = ResNet18Engine(B, 3, 3, size, "max", "cuda:0", "bf16")
eng.load_state_dict(sd)
eng.capture_graph(x, y)
loss = eng.train_step(x, y).item()
torch.cuda.synchronize()
torch.save({{"loss": loss, "grads": eng.grads.cpu(), "flat": eng.flat.cpu()}}, sys.argv[1])
"""


def test_pdl_on_and_off_give_the_same_step(tmp_path):
    """PRIMIA_PDL is read once per process, so each setting runs in its own interpreter."""
    outs = {}
    for pdl in ("1", "0"):
        out = tmp_path / f"pdl{pdl}.pt"
        env = dict(os.environ, PRIMIA_PDL=pdl)
        r = subprocess.run([sys.executable, "-c", _WORKER.format(root=ROOT), str(out)], env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[pdl] = torch.load(out)
    a, b = outs["1"], outs["0"]
    _assert_same_step((a["loss"], a["grads"], a["flat"]), (b["loss"], b["grads"], b["flat"]), "PDL on vs off")


# ------------------------------------------------------------------------------------------------ C2 against the oracle
# Per-tensor gates for the bf16 throughput mode at the bench configuration, FROZEN from the first B200 measurement
# (gpurun_out/bf16_c2_errors.json, copied to profiles/r02_bf16_c2_errors.json): gate = measured x 1.5.  bench.py prints the
# same numbers in its JSON line ("bf16_parity").  bf16 carries 8 significand bits, so 1e-5 is out of reach by construction;
# the fp32-accurate tensor-core mode (mode="f32x3", tests/test_train_x3_gpu.py) is the one held to 1e-5.
# First measurement (round 2, B200): loss_rel 3.3e-4, logits_rel 1.7e-2, grad_cos_min 0.888 (layer1.0.bn2.weight),
# grad_rel_max 0.484, grad_rel_median 0.353 -- the gradient error is what bf16 STORAGE of dy costs at random init (BatchNorm
# backward subtracts a per-channel common mode 1e3-1e4 x larger than the fluctuation that carries the signal); PyTorch's own
# CPU bf16 autocast of the oracle shows the same figures, which the test also checks tensor by tensor.
C2_GATES = {"loss_rel": 1e-3, "logits_rel": 2.6e-2, "grad_cos_min": 0.83, "grad_rel_max": 0.72, "grad_rel_median": 0.53}


def test_bf16_step_at_bench_config_c2_vs_fp32_oracle():
    B, size = 64, 224
    torch.manual_seed(42)
    m = O.ResNet18(input_size=size)
    (x, y), = _batches(B, size, 1, seed=42)
    eng = _engine(B, size, m.state_dict())
    eng.capture_graph(x.to(DEV), y.to(DEV))
    loss_gpu = eng.train_step(x.to(DEV), y.to(DEV)).item()      # the graph replay bench.py times
    torch.cuda.synchronize()
    import copy

    m_amp = copy.deepcopy(m)
    m.train()
    out = m(x)
    loss = torch.nn.functional.cross_entropy(out, y)
    loss.backward()
    m_amp.train()
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out_amp = m_amp(x)
    torch.nn.functional.cross_entropy(out_amp.float(), y).backward()
    gd = eng.grad_dict()
    per = {n: {"rel": rel(gd[n], p.grad), "cos": cos(gd[n], p.grad), "autocast_rel": rel(pa.grad, p.grad),
               "autocast_cos": cos(pa.grad, p.grad)} for (n, p), pa in zip(m.named_parameters(), m_amp.parameters())}
    rels = sorted(v["rel"] for v in per.values())
    rec = {"config": "C2: B=64, 224x224, bf16, CUDA graph replay", "loss_gpu": loss_gpu, "loss_oracle": loss.item(),
           "loss_rel": abs(loss_gpu - loss.item()) / abs(loss.item()), "logits_rel": rel(eng.logits, out.detach()),
           "grad_cos_min": min(v["cos"] for v in per.values()), "grad_rel_max": rels[-1], "grad_rel_median": rels[len(rels) // 2],
           "worst_cos": min(per.items(), key=lambda kv: kv[1]["cos"])[0], "worst_rel": max(per.items(), key=lambda kv: kv[1]["rel"])[0],
           "per_tensor": per}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bf16_c2_errors.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print({k: v for k, v in rec.items() if k != "per_tensor"})
    deficit = {n: v["autocast_cos"] - v["cos"] for n, v in per.items()}
    rec["worst_cos_deficit_vs_autocast"] = max(deficit.items(), key=lambda kv: kv[1])
    rec["mean_cos_deficit_vs_autocast"] = sum(deficit.values()) / len(deficit)
    with open(os.path.join(ROOT, "gpurun_out", "bf16_c2_errors.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print("vs CPU bf16 autocast:", rec["worst_cos_deficit_vs_autocast"], rec["mean_cos_deficit_vs_autocast"])
    assert rec["worst_cos_deficit_vs_autocast"][1] < 0.06 and rec["mean_cos_deficit_vs_autocast"] < 0.01
    assert rec["loss_rel"] < C2_GATES["loss_rel"]
    assert rec["logits_rel"] < C2_GATES["logits_rel"]
    assert rec["grad_cos_min"] > C2_GATES["grad_cos_min"]
    assert rec["grad_rel_max"] < C2_GATES["grad_rel_max"]
    assert rec["grad_rel_median"] < C2_GATES["grad_rel_median"]
