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


need3 = pytest.mark.skipif(torch.cuda.device_count() < 3, reason="needs 3 GPUs")


@need2
def test_newton_p2p_across_two_gpus_equals_fused():
    """the cross-GPU BatchNorm Newton kernel pair (openings over NVLink mailboxes) == the single-GPU fused kernel"""
    from primia_b200.ring import ops

    g = torch.Generator().manual_seed(31)
    iters, scale, Cc = 80, 10 ** 4, 20
    jobs0, jobs2 = [], []
    for C in (64, 256, 512):
        vq = R.encode(torch.rand(C, generator=g) * 0.5 + 0.75, 10, 4)
        vs = R.share_from_random(vq, rnd(g, vq.shape))
        a, b = rnd(g, (3 * (iters - 1), C)), rnd(g, (3 * (iters - 1), C))
        c = a * b
        a0, b0, c0 = rnd(g, a.shape), rnd(g, b.shape), rnd(g, c.shape)
        tri = [[a0, b0, c0], [a - a0, b - b0, c - c0]]
        kq = torch.full((iters,), 21 * scale, dtype=torch.int64)
        k0 = rnd(g, kq.shape)
        ks = [k0, kq - k0]
        jobs0.append(([vs[0].cuda(0), vs[1].cuda(0)], [[t.cuda(0) for t in tri[j]] for j in range(2)], (ks[0].cuda(0), ks[1].cuda(0))))
        jobs2.append(([vs[0].cuda(0), vs[1].cuda(1)], [[t.cuda(j) for t in tri[j]] for j in range(2)], (ks[0].cuda(0), ks[1].cuda(1))))
    with torch.cuda.device(0):
        ref = ops.bn_newton_fused(jobs0, iters, scale, Cc)
    for rep in range(2):
        got = ops.bn_newton_p2p(jobs2, iters, scale, Cc)
        torch.cuda.synchronize(0); torch.cuda.synchronize(1)
        assert all(int(e.item()) == 0 for e in ops.bn_newton_p2p.last_err)
        for (r0, r1), (g0, g1) in zip(ref, got):
            assert g0.device.index == 0 and g1.device.index == 1
            assert torch.equal(r0.cpu(), g0.cpu()) and torch.equal(r1.cpu(), g1.cpu())


def _enc_forward(devs, graph, size=32, pf=4, seed=5):
    """the whole encrypted forward for 2 images with fixed seeds on the given (model_owner, data_owner, provider) devices"""
    import primia_b200.ring as ring
    from oracle import train_oracle as O
    from primia_b200.ring.resnet import EncryptedInferenceGraph
    from primia_b200.ring.tensors import ShareRNG

    torch.manual_seed(42)
    model = O.ResNet18(input_size=size).eval()
    parties = [ring.Party("model_owner", devs[0]), ring.Party("data_owner", devs[1])]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", devs[2]), seed=seed)
    net = ring.EncryptedResNet18.from_state_dict(model.state_dict(), parties, prov, 10, pf, input_size=size, rng=ShareRNG(seed=99))
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randn(1, 3, size, size, generator=g) for _ in range(3)]
    outs = []
    if graph:
        eg = EncryptedInferenceGraph(net, imgs[0])
        for img in imgs[1:]:
            eg.offline()
            logits, _ = eg.online(img)
            for d in set(devs):
                torch.cuda.synchronize(d)
            outs.append(([s.cpu().clone() for s in eg.out_shares.child.child], logits.cpu().clone()))
    else:
        for img in imgs[1:]:
            o = net.forward(net.share_input(img))
            for d in set(devs):
                torch.cuda.synchronize(d)
            outs.append(([s.cpu() for s in o.child.child], o.get().float_prec().cpu()))
    model.pool, model.relu = model.relu, model.pool
    with torch.no_grad():
        want = [model(img) for img in imgs[1:]]
    return outs, want


@need3
@pytest.mark.parametrize("graph", [False, True])
def test_encrypted_forward_three_gpu_placement_equals_single_gpu(graph):
    """SURVEY.md section 8e placement -- model_owner cuda:0, data_owner cuda:1, crypto provider cuda:2, every opening a peer
    read over NVLink, the BatchNorm Newton iterations exchanged through P2P mailboxes -- gives exactly the shares of the
    one-GPU placement (same Philox seeds => same primitives and sharings), eagerly and as one multi-device CUDA graph."""
    one, want = _enc_forward(["cuda:0", "cuda:0", "cuda:0"], graph)
    three, _ = _enc_forward(["cuda:0", "cuda:1", "cuda:2"], graph)
    for (s1, l1), (s3, l3), w in zip(one, three, want):
        assert torch.equal(s1[0], s3[0]) and torch.equal(s1[1], s3[1]), "shares differ between placements"
        assert torch.equal(l1, l3) and (l3 - w).abs().max() < 0.15


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
# the overlapped schedule (two-graph step, layer4 bucket all-reduced during the rest of the backward) == step, then FedAvg
B, size = 8, 64
torch.manual_seed(5)
base = O.ResNet18(input_size=size).state_dict()
g = torch.Generator().manual_seed(10 + rank)
x, y = torch.randn(B, 3, size, size, generator=g).cuda(), torch.randint(0, 3, (B,), generator=g).cuda()
a = ResNet18Engine(B, 3, 3, size, "max", f"cuda:{rank}", "bf16"); a.load_state_dict(base)
b = ResNet18Engine(B, 3, 3, size, "max", f"cuda:{rank}", "bf16"); b.load_state_dict(base)
a.train_step(x, y); aggregation([HospitalWorker(ids[rank], a)], None, dist.group.WORLD)
b.capture_graph_overlap(x, y)
HospitalWorker(ids[rank], b).local_step_and_fedavg(x, y, dist.group.WORLD)
torch.cuda.synchronize()
err = ((a.flat - b.flat).norm() / a.flat.norm()).item()
assert err < 1e-5, err
chk = b.flat.clone(); dist.all_reduce(chk, op=dist.ReduceOp.AVG)
assert torch.allclose(chk, b.flat, rtol=0, atol=1e-7), "ranks must hold the same averaged state"
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
