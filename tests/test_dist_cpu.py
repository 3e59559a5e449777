"""N>1 host logic on CPU: world_size-2 gloo run of the FedAvg reduction plan (pre-scale, SUM all-reduce, post-scale)
against the oracle's aggregation; the reference-arm bench line under 2 ranks; the PySyft-shaped verbs on CPU tensors."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PM_ROOT"])
from oracle import train_oracle as O
from primia_b200.train.federated import fedavg_scales
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
ids = ["alice", "bob"]
torch.manual_seed(100)
models = {}
for i, w in enumerate(ids):
    torch.manual_seed(100 + i)
    models[w] = O.ResNet18(input_size=32)
keys = [k for k in models["alice"].state_dict() if "num_batches_tracked" not in k]
for weights in (None, {"alice": 0.25, "bob": 0.75}):
    local = O.ResNet18(input_size=32)
    O.aggregation(local, models, ids, weights)
    mine = models[ids[rank]].state_dict()
    flat = torch.cat([mine[k].flatten().float() for k in keys])
    pre, post = fedavg_scales(ids[rank], world, weights)
    flat = flat * pre
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat = flat * post
    ref = torch.cat([local.state_dict()[k].flatten().float() for k in keys])
    err = ((flat - ref).norm() / ref.norm()).item()
    assert err < 1e-6, err
dist.barrier()
if rank == 0:
    print("DIST_OK")
dist.destroy_process_group()
'''


def test_fedavg_plan_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, PM_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DIST_OK" in r.stdout


WEIGHTS_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PM_ROOT"])
from train import global_batch_counts
from primia_b200.train.federated import fedavg_scales
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
# one hospital per rank, loaders of different length (utils.py:953-957: w_i = len(tl_i) / sum_j len(tl_j) over ALL hospitals)
local = {["alice", "bob"][rank]: [3, 5][rank]}
counts = global_batch_counts(local, dist.group.WORLD)
assert counts == {"alice": 3, "bob": 5}, counts
total = sum(counts.values())
weights = {n: c / total for n, c in counts.items()}
x = torch.full((4,), float(rank + 1))
pre, post = fedavg_scales(["alice", "bob"][rank], world, weights)
y = x * pre
dist.all_reduce(y, op=dist.ReduceOp.SUM)
y = y * post
assert torch.allclose(y, torch.full((4,), 1 * 3 / 8 + 2 * 5 / 8)), y      # the weighted mean, not n times the parameters
# number of rounds of the epoch = the MAXIMUM batch count over the ranks (exhausted hospitals keep joining the collectives)
t = torch.tensor([[3, 5][rank]]); dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert int(t) == 5
dist.barrier()
if rank == 0:
    print("WEIGHTS_OK")
dist.destroy_process_group()
'''


def test_weighted_averaging_denominator_is_global_world_size_2_gloo(tmp_path):
    """under torchrun every rank holds one hospital, so the weights of torchlib/utils.py:953-957 need the batch counts
    of ALL ranks (train.global_batch_counts); with rank-local counts every weight was 1.0 and FedAvg returned n x the parameters"""
    script = tmp_path / "w.py"
    script.write_text(WEIGHTS_WORKER)
    env = dict(os.environ, PM_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29535", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "WEIGHTS_OK" in r.stdout


def test_reference_arm_prints_one_json_line_rank0_only():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0", "--ref-batch", "2"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]  # stdout carries the JSON line and nothing else (bench.claim_stdout)
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "federated_round_images_per_sec" and d["n_gpus"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0


def test_pysyft_verbs_on_cpu():
    import primia_b200.sy as sy

    hook = sy.TorchHook(torch)
    alice = sy.VirtualWorker(hook, id="alice", device="cpu")
    bob = sy.VirtualWorker(hook, id="bob", device="cpu")
    for w, n in ((alice, 10), (bob, 7)):
        d = torch.arange(n * 2, dtype=torch.float32).view(n, 2).tag("#traindata")
        t = torch.arange(n).tag("#traintargets")
        w.load_data([d.send(w).get(), t.send(w).get()])
    grid = sy.PrivateGridNetwork(alice, bob)
    found = grid.search("#traindata")
    assert set(found) == {"alice", "bob"} and found["bob"][0].location is bob and found["bob"][0].shape == (7, 2)
    ds = sy.BaseDataset(found["bob"][0], grid.search("#traintargets")["bob"][0])
    tl = sy.FederatedDataLoader(sy.FederatedDataset([ds]), batch_size=3, shuffle=True)
    assert len(tl) == 3 and tl.federated_dataset.workers == ["bob"]
    seen = []
    for data, target in tl:
        assert data.location is bob
        seen += target.get().tolist()
        assert torch.equal(data.get()[:, 0], target.get().float() * 2)
    assert sorted(seen) == list(range(7))
