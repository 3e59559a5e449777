"""Host-side logic that needs no GPU: worker roster formats, configuration object, learning-rate schedule, MixUp, the
checkpoint's pickled ``args`` (reference: torchlib/run_websocket_server.py:6-8, torchlib/utils.py:37-89,92-302,327-400)."""
import argparse
import configparser
import io
import math
import pickle

import torch

from torchlib.run_websocket_server import read_websocket_config, worker_names
from torchlib.utils import Arguments, LearningRateScheduler, MixUp

REFERENCE_ROSTER = "id,alice,bob,charlie,crypto_provider\nhost,127.0.0.1,127.0.0.1,127.0.0.1,127.0.0.1\nport,8777,8778,8779,8780\n"


def test_reference_column_major_roster(tmp_path):
    """the reference's configs/websetting/config.csv (one column per worker) through the reference's call chain
    ``[d["id"] for _, d in read_websocket_config(path).items()]`` (utils.py:521-522)"""
    p = tmp_path / "config.csv"
    p.write_text(REFERENCE_ROSTER)
    d = read_websocket_config(str(p))
    assert list(d.keys()) == [1, 2, 3, 4]                       # pandas: column numbers
    assert d[1] == {"id": "alice", "host": "127.0.0.1", "port": "8777"}
    assert worker_names(str(p)) == ["alice", "bob", "charlie", "crypto_provider"]
    try:                                                         # and identical to what pandas (the reference's reader) returns
        from pandas import read_csv

        ref = read_csv(str(p), header=None, index_col=0).to_dict()
        assert {k: {a: str(b) for a, b in v.items()} for k, v in ref.items()} == d
    except ImportError:
        pass


def test_row_major_roster_still_accepted(tmp_path):
    p = tmp_path / "rows.csv"
    p.write_text("id,host,port\nalice,127.0.0.1,8777\nbob,127.0.0.1,8778\n")
    assert worker_names(str(p)) == ["alice", "bob"]


def test_repo_roster_is_the_reference_layout():
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = worker_names(os.path.join(root, "configs/websetting/config.csv"))
    assert names[-1] == "crypto_provider" and "host" not in names and "port" not in names


def _args(extra=""):
    cfg = configparser.ConfigParser()
    cfg.read_string("[config]\nbatch_size = 8\ntrain_resolution = 64\nepochs = 10\nlr = 1e-4\nend_lr = 1e-5\nrestarts = 0\n"
                    "optimizer = Adam\nbeta1 = 0.5\nbeta2 = 0.99\nweight_decay = 5e-4\nmodel = resnet-18\nseed = 42\n"
                    "[federated]\nsync_every_n_batch = 2\nkeep_optim_dict = no\nweighted_averaging = yes\n" + extra)
    cmd = argparse.Namespace(train_federated=True, unencrypted_aggregation=True, data_dir=None, cuda=True)
    return Arguments(cmd, cfg, mode="train", verbose=False)


def test_arguments_and_pickle_roundtrip():
    a = _args()
    assert (a.batch_size, a.train_resolution, a.sync_every_n_batch, a.weighted_averaging, a.keep_optim_dict) == (8, 64, 2, True, False)
    assert a.beta1 == 0.5 and a.precision_fractional == 16 and a.inference_resolution == 64
    b = pickle.loads(pickle.dumps(a))                            # what torch.save does with the checkpoint's "args"
    assert b.__dict__ == a.__dict__
    buf = io.BytesIO()
    torch.save({"args": a}, buf)
    buf.seek(0)
    assert torch.load(buf, weights_only=False)["args"].lr == a.lr
    ns = argparse.Namespace(lr=3e-4, model="resnet-18")
    assert Arguments.from_namespace(ns).lr == 3e-4
    a.from_previous_checkpoint(argparse.Namespace(encrypted_inference=True, cuda=False))
    assert a.encrypted_inference is True and a.mixup is False


def test_mixup_doubles_batch_when_always_applied():
    a = _args("[augmentation]\nmixup = yes\nmixup_prob = 1.0\nmixup_lambda = 0.3\n")
    assert a.batch_size == 16


def test_learning_rate_schedule_is_the_references_formula():
    """utils.py:56-60: 10 ** ((log_end - log_start) / total_epochs * epoch + log_start); end_lr is never reached"""
    s = LearningRateScheduler(10, math.log10(1e-4), math.log10(1e-5), restarts=0)
    assert abs(s.get_lr(0) - 1e-4) < 1e-12
    for e in range(10):
        assert abs(s.get_lr(e) - 10 ** (-4 - e / 10)) < 1e-15
    assert s.get_lr(9) > 1e-5
    r = LearningRateScheduler(10, -4, -5, restarts=1)           # period 5
    assert abs(r.get_lr(5) - 1e-4) < 1e-12 and abs(r.get_lr(7) - r.get_lr(2)) < 1e-15
    c = LearningRateScheduler(8, -3, -5, schedule_plan="log_cosine")
    assert abs(c.get_lr(0) - 1e-3) < 1e-12 and abs(c.get_lr(4) - 1e-4) < 1e-12

    class Opt:
        lr = 0.0
    o = Opt()
    assert s.adjust_learning_rate(o, 3) == o.lr == s.get_lr(3)
    t = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1.0)
    s.adjust_learning_rate(t, 2)
    assert t.param_groups[0]["lr"] == s.get_lr(2)


def test_mixup_batched():
    """utils.py:327-400 on a batch: halves blended with lambda; odd batch keeps its last sample"""
    x = torch.arange(6 * 2, dtype=torch.float32).view(6, 2)
    y = torch.nn.functional.one_hot(torch.tensor([0, 1, 2, 0, 1, 2]), 3).float()
    mx, my = MixUp(λ=0.25, p=None)((x, y))
    assert torch.allclose(mx, 0.25 * x[:3] + 0.75 * x[3:]) and torch.allclose(my, 0.25 * y[:3] + 0.75 * y[3:])
    assert torch.allclose(my.sum(1), torch.ones(3))
    mx, my = MixUp(λ=0.25, p=None)((x[:5], y[:5]))
    assert mx.shape[0] == 3 and torch.equal(mx[-1], x[4]) and torch.allclose(mx[:2], 0.25 * x[:2] + 0.75 * x[2:4])
    one = MixUp(λ=0.5, p=None)((x[:1], y[:1]))
    assert torch.equal(one[0], x[:1])


def test_stable_seed_is_process_independent():
    import subprocess
    import sys
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"import sys; sys.path.insert(0, {root!r}); from train import stable_seed; print(stable_seed(42, 'alice'), stable_seed(42, 'bob'))"
    outs = {subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONHASHSEED=str(h))).stdout
            for h in (1, 2)}
    assert len(outs) == 1 and outs.pop().strip()
