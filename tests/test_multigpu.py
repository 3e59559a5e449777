"""2-GPU checks (skipped with fewer devices): SPDZ parties on separate GPUs opening delta/eps through peer-mapped
pointers stay bit-exact; NCCL FedAvg of two hospitals equals the oracle's aggregation."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import ring_oracle as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
need2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")


def rnd(g, shape):
    return torch.randint(-(2 ** 63), 2 ** 63 - 1, tuple(shape), dtype=torch.int64, generator=g)


@need2
def test_parties_on_two_gpus_bit_exact():
    import primia_b200.ring as ring

    g = torch.Generator().manual_seed(21)
    B, C, H, Co, k, s, p = 1, 64, 28, 128, 3, 2, 1
    x, w = rnd(g, (B, C, H, H)), rnd(g, (Co, C, k, k))
    xs, ws = R.share_from_random(x, rnd(g, x.shape)), R.share_from_random(w, rnd(g, w.shape))
    Ho = (H + 2 * p - k) // s + 1
    M, K, N = Ho * Ho, C * k * k, Co
    a, b = rnd(g, (B, M, K)), rnd(g, (K, N))
    c = R.build_triple_c(a, b, "matmul")
    a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
    tri = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
    ref = R.conv2d_shared(xs, ws, tri, s, p, 10, 16)
    parties = [ring.Party("model_owner", "cuda:0"), ring.Party("data_owner", "cuda:1")]
    for j, pty in enumerate(parties):
        pty.crypto_store.add_primitives("matmul", ((B, M, K), (K, N)), [tuple(t.to(pty.device) for t in tri[j])])
    X = ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([xs[j].to(parties[j].device) for j in range(2)], parties), 10, 16)
    W = ring.FixedPrecisionTensor(ring.AdditiveSharingTensor([ws[j].to(parties[j].device) for j in range(2)], parties), 10, 16)
    out = ring.functional.conv2d(X, W, None, s, p)
    for j in range(2):
        assert out.child.child[j].device == parties[j].device
        assert torch.equal(out.child.child[j].cpu(), ref[j])
    rec = out.child.get().cpu()
    assert torch.equal(rec, ref[0] + ref[1])


NCCL_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PM_ROOT"])
from oracle import train_oracle as O
from primia_b200.train import HospitalWorker, ResNet18Engine, aggregation
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{rank}"))
ids = ["alice", "bob"]
models = {}
for i, w in enumerate(ids):
    torch.manual_seed(100 + i); models[w] = O.ResNet18(input_size=64)
for weights in (None, {"alice": 0.25, "bob": 0.75}):
    local = O.ResNet18(input_size=64)
    O.aggregation(local, models, ids, weights)
    eng = ResNet18Engine(2, 3, 3, 64, "max", f"cuda:{rank}", "f32")
    eng.load_state_dict(models[ids[rank]].state_dict())
    aggregation([HospitalWorker(ids[rank], eng)], weights, dist.group.WORLD)
    sd = eng.state_dict()
    for k, v in local.state_dict().items():
        if "num_batches_tracked" in k: continue
        err = ((sd[k].cpu().double() - v.double()).norm() / v.double().norm().clamp_min(1e-30)).item()
        assert err < 1e-6, (k, err)
    # secure aggregation (reference default): shares travel by all-to-all, reconstruction by int64 all-reduce
    eng.load_state_dict(models[ids[rank]].state_dict())
    aggregation([HospitalWorker(ids[rank], eng)], weights, dist.group.WORLD, secure=True, precision_fractional=16)
    sd = eng.state_dict()
    for k in local.state_dict():
        if "num_batches_tracked" in k: continue
        ts = [models[w].state_dict()[k] for w in ids]
        ref = O.secure_aggregation_value(ts, [weights[w] for w in ids] if weights else [1, 1], 10, 16)
        if weights is None: ref = ref / 2
        assert torch.equal(sd[k].cpu(), ref.reshape(sd[k].shape)), k
dist.barrier()
if rank == 0: print("NCCL_FEDAVG_OK")
dist.destroy_process_group()
'''


@need2
def test_nccl_fedavg_two_hospitals(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(NCCL_WORKER)
    env = dict(os.environ, PM_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_FEDAVG_OK" in r.stdout
