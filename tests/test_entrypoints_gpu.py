"""The two entry points end to end on the GPU, through the PySyft-shaped verbs:

  train.py --train_federated  ->  checkpoint in the reference's format (torchlib/utils.py:1470-1493)
      ->  inference.py (plain)                    == the oracle's eval forward of the same weights
      ->  inference.py --encrypted_inference      == the plain logits up to fixed-point error, same argmax;
  a checkpoint in the reference's shape written by the ORACLE loads into both; --resume_checkpoint restores model + Adam
  state; module-level fix_precision / share / get / float_precision iterate parameters AND buffers (hook.py:626-632)."""
import os
import sys

import pytest
import torch

from oracle import train_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

INI = """[config]
batch_size = 4
train_resolution = {res}
inference_resolution = {res}
test_batch_size = 4
test_interval = 1
validation_split = 10
epochs = {epochs}
lr = 1e-4
end_lr = 1e-5
restarts = 0
beta1 = 0.5
beta2 = 0.99
weight_decay = 5e-4
deterministic = yes
seed = 42
optimizer = Adam
model = resnet-18
pretrained = no
weight_classes = {wc}
pooling_type = {pool}
[augmentation]
mixup = {mixup}
mixup_prob = 0.5
mixup_lambda = 0.4
[federated]
sync_every_n_batch = 1
keep_optim_dict = {keep}
repetitions_dataset = 1
weighted_averaging = {weighted}
precision_fractional = 16
"""


def write_ini(tmp_path, res=64, epochs=1, keep="no", weighted="no", pool="max", wc="no", mixup="no", name="c.ini"):
    p = tmp_path / name
    p.write_text(INI.format(res=res, epochs=epochs, keep=keep, weighted=weighted, pool=pool, wc=wc, mixup=mixup))
    return str(p)


def reference_shaped_checkpoint(path, model, res, pooling="max"):
    """a file with exactly the keys and nesting the reference writes (utils.py:1481-1492), ``args`` a pickled Arguments"""
    import argparse
    import configparser

    from torchlib.utils import Arguments

    cfg = configparser.ConfigParser()
    cfg.read_string(INI.format(res=res, epochs=1, keep="no", weighted="no", pool=pooling, wc="no", mixup="no"))
    args = Arguments(argparse.Namespace(train_federated=True, unencrypted_aggregation=True, data_dir=None, cuda=True), cfg,
                     verbose=False)
    opt = O.make_optimizer(model)
    torch.save({"epoch": 1, "model_state_dict": model.state_dict(), "optim_state_dict": {"alice": opt.state_dict()},
                "args": args, "val_mean_std": torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])}, path)


def test_train_writes_reference_format_checkpoint_and_inference_reads_it(tmp_path, capsys):
    import inference
    import train

    ini = write_ini(tmp_path, res=64, epochs=2, weighted="yes")
    hospitals = train.main(["--config", ini, "--train_federated", "--unencrypted_aggregation", "--mode", "f32",
                            "--batches_per_worker", "2", "--save_dir", str(tmp_path / "w")])
    assert [h.id for h in hospitals] == ["alice", "bob", "charlie"]           # crypto_provider is not a hospital
    ck = train.main.last_checkpoint
    state = torch.load(ck, weights_only=False)
    assert set(state) == {"epoch", "model_state_dict", "optim_state_dict", "args", "val_mean_std"}
    assert type(state["args"]).__module__ == "torchlib.utils" and state["args"].train_federated
    assert set(state["optim_state_dict"]) == {"alice", "bob", "charlie"}
    osd = state["optim_state_dict"]["alice"]
    assert set(osd) == {"state", "param_groups"} and osd["param_groups"][0]["betas"] == (0.5, 0.99)
    ref = O.ResNet18(input_size=64)
    assert list(state["model_state_dict"].keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(state["model_state_dict"])                            # loads into the reference architecture
    # all hospitals hold the aggregate after the final FedAvg
    assert torch.equal(hospitals[0].engine.flat.cpu(), hospitals[2].engine.flat.cpu())   # (hospitals round-robin over the visible GPUs)
    # plain inference on the checkpoint == the oracle's eval forward on the same synthetic images
    preds = inference.main(["--model_weights", ck, "--num_images", "3"])
    g = torch.Generator().manual_seed(42)
    imgs = torch.randn(3, 3, 64, 64, generator=g)
    ref.eval()
    with torch.no_grad():
        want = ref(imgs)
    assert preds == want.argmax(1).tolist()
    assert (inference.main.last_logits - want[-1:]).abs().max() < 1e-4 * want.abs().max().clamp_min(1)
    out = capsys.readouterr().out
    assert '"Inference Results"' in out and "Took" in out


def test_encrypted_inference_entry_point_matches_plain_model(tmp_path):
    """inference.py --encrypted_inference on a reference-shaped checkpoint written by the oracle: the decoded logits track
    the plaintext model (pf = 4 keeps the 32-bit FSS comparison meaningful; the reference's pf = 16 is exercised share for
    share in tests/test_fss_gpu.py) -- both through the literal verb flow and through the captured CUDA graph."""
    import inference

    res = 32
    torch.manual_seed(42)
    model = O.ResNet18(input_size=res)
    with torch.no_grad():
        gg = torch.Generator().manual_seed(7)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=gg) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=gg) * 0.5 + 0.75)
    ck = str(tmp_path / "ref.pt")
    reference_shaped_checkpoint(ck, model, res)
    g = torch.Generator().manual_seed(42)
    imgs = torch.randn(2, 3, res, res, generator=g)
    model.eval()
    with torch.no_grad():
        want = model(imgs)
    for extra in ([], ["--cuda_graph"]):
        preds = inference.main(["--model_weights", ck, "--encrypted_inference", "--num_images", "2", "--precision_fractional", "4"]
                               + extra)
        assert (inference.main.last_logits - want[-1:]).abs().max() < 0.15, (extra, inference.main.last_logits, want[-1:])
        assert len(preds) == 2
    plain = inference.main(["--model_weights", ck, "--num_images", "2"])
    assert plain == want.argmax(1).tolist()


def test_resume_checkpoint_restores_model_and_optimizer(tmp_path):
    import train

    ini = write_ini(tmp_path, res=64, epochs=1, keep="yes")
    h1 = train.main(["--config", ini, "--train_federated", "--unencrypted_aggregation", "--mode", "f32", "--batches_per_worker", "2",
                     "--save_dir", str(tmp_path / "a")])
    ck = train.main.last_checkpoint
    flat, m, steps = h1[0].engine.flat.cpu().clone(), h1[0].engine.adam_m.cpu().clone(), h1[0].engine.step_count
    assert steps == 2 and m.abs().sum() > 0
    ini2 = write_ini(tmp_path, res=64, epochs=1, keep="yes", name="c2.ini")   # start_at_epoch == epochs: one more epoch runs
    import primia_b200.train.federated as F

    seen = {}
    orig = F.federated_round

    def spy(workers, *a, **k):
        seen["flat"] = workers[0].engine.flat.cpu().clone()
        seen["m"] = workers[0].engine.adam_m.cpu().clone()
        seen["steps"] = workers[0].engine.step_count
        return orig(workers, *a, **k)

    import unittest.mock as mock

    with mock.patch("primia_b200.train.federated_round", spy):
        train.main(["--config", ini2, "--train_federated", "--unencrypted_aggregation", "--mode", "f32", "--batches_per_worker", "2",
                    "--resume_checkpoint", ck, "--save_dir", str(tmp_path / "b")])
    assert torch.equal(seen["flat"], flat), "resumed model differs from the checkpointed one"
    assert seen["steps"] == steps and torch.allclose(seen["m"], m, rtol=0, atol=0)


def test_module_verbs_iterate_parameters_and_buffers():
    import primia_b200.sy as sy
    from primia_b200.models import resnet18

    hook = sy.TorchHook(torch)
    alice, bob, cp = (sy.VirtualWorker(hook, id=n, device="cuda:0") for n in ("alice", "bob", "cp"))
    model = resnet18(num_classes=3, pooling="max", adptpool=False, input_size=32)
    with torch.no_grad():
        model.bn1.running_mean.normal_()
    sent = model.copy().send(alice)
    assert sent.location is alice and model.location is None and next(sent.parameters()).is_cuda
    assert sent.get().location is None
    model.cuda()
    ref = {k: v.clone() for k, v in model.state_dict().items()}
    model.fix_precision(precision_fractional=4, dtype="long").share(alice, bob, crypto_provider=cp, protocol="fss",
                                                                    requires_grad=False)
    keys = set(model._sy_shared)
    assert "bn1.running_mean" in keys and "bn1.running_var" in keys and "conv1.weight" in keys and "fc.bias" in keys
    assert not any(k.endswith("num_batches_tracked") for k in keys)
    assert len(keys) == 122 - 20
    sh = model._sy_shared["bn1.running_mean"].child
    assert len(sh.child) == 2 and not torch.equal(sh.child[0] + sh.child[1], sh.child[0])
    model.get().float_precision()                                           # reconstruct + decode back into the module
    for k, v in model.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            assert (v - ref[k]).abs().max() <= 1e-4 + 1e-4 * ref[k].abs().max(), k
    with pytest.raises(TypeError):
        torch.randn(3).share(alice, bob, crypto_provider=cp)               # native.py:931-932
    loss_fn = torch.nn.CrossEntropyLoss().send(alice)
    assert loss_fn.location is alice


def test_federated_dataloader_on_gpu():
    """FederatedDataLoader (fl/dataloader.py:159-258) for PriMIA's single-worker datasets: batches are gathered on the owner's
    GPU, every sample appears exactly once per epoch, shuffling changes the order between epochs"""
    import primia_b200.sy as sy

    hook = sy.TorchHook(torch)
    w = sy.VirtualWorker(hook, id="alice", device="cuda:0")
    data = torch.arange(10, dtype=torch.float32).view(10, 1, 1, 1).repeat(1, 3, 2, 2).tag("#traindata")
    target = torch.arange(10).tag("#traintargets")
    w.load_data([data.send(w).get(), target.send(w).get()])
    grid = sy.PrivateGridNetwork(w)
    ds = sy.BaseDataset(grid.search("#traindata")["alice"][0], grid.search("#traintargets")["alice"][0])
    loader = sy.FederatedDataLoader(sy.FederatedDataset([ds]), batch_size=4, shuffle=True, seed=1)
    assert len(loader) == 3 and loader.federated_dataset.workers == ["alice"]
    orders = []
    for _ in range(2):
        seen = []
        for d, t in loader:
            assert d.location is w and d.get().is_cuda and d.get().shape[1:] == (3, 2, 2)
            assert torch.equal(d.get()[:, 0, 0, 0].long(), t.get())
            seen += t.get().tolist()
        assert sorted(seen) == list(range(10))
        orders.append(seen)
    assert orders[0] != orders[1]
    assert [len(b[1]) for b in sy.FederatedDataLoader(sy.FederatedDataset([ds]), batch_size=4, drop_last=True)] == [4, 4]


@pytest.mark.parametrize("mode", ["f32", "bf16"])
def test_avg_pooling_and_adaptive_pool_variants(mode):
    """pooling_type = avg (models.py:386-387) and adptpool (models.py:400-402) through the engine vs the oracle"""
    from primia_b200.train import ResNet18Engine

    B, size = 4, 64
    torch.manual_seed(3)
    m = O.ResNet18(input_size=size, pooling="avg", adptpool=True)
    eng = ResNet18Engine(B, 3, 3, size, "avg", "cuda:0", mode, adptpool=True)
    eng.load_state_dict(m.state_dict())
    g = torch.Generator().manual_seed(4)
    x, y = torch.randn(B, 3, size, size, generator=g), torch.randint(0, 3, (B,), generator=g)
    m.train()
    loss = torch.nn.functional.cross_entropy(m(x), y)
    loss.backward()
    eng.forward(x.cuda())
    l = eng.loss_and_backward(y.cuda()).item()
    torch.cuda.synchronize()
    tol = 1e-5 if mode == "f32" else 3e-2
    assert abs(l - loss.item()) / loss.item() < tol
    gd = eng.grad_dict()
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()
    if mode == "f32":
        assert rel(gd["conv1.weight"], m.conv1.weight.grad) < 5e-3 and rel(gd["fc.weight"], m.fc.weight.grad) < 1e-5
