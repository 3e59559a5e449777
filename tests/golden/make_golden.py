"""Generate the committed golden fixtures by EXECUTING THE REFERENCE'S OWN SOURCES.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Nothing is copied from the reference: plain-torch functions are located in the reference
files with ``ast`` and exec'd in a namespace holding minimal stubs for the PySyft runtime
(``allow_command`` -> identity; ``x.owner.crypto_store.get_keys`` -> a dict lookup).  Their
*outputs* on seeded inputs are stored as .npz fixtures next to this script.

Fixtures
  ring_preconv.npz   _pre_conv/_post_conv  (syft/frameworks/torch/nn/functional.py:79-201)
  ring_spdz.npz      spdz_mask/spdz_compute/triple_mat_mul (syft/frameworks/torch/mpc/spdz.py:22-122)
  ring_newton.npz    control flow of reciprocal(method="newton") run on exact python ints
                     (syft/frameworks/torch/tensors/interpreters/precision.py:507-518)
  ring_pool.npz      _pre_pool/_post_pool and the op trace of _pool2d's max branch (nn/functional.py:312-525)
  fss_dif.npz        DIF keygen / eval / H executed from syft/frameworks/torch/mpc/fss.py (shaloop -> hashlib)
  train_ref.npz      torchlib/models.py resnet18: state_dict keys/shapes, logits, loss, grad digests
"""
import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def extract(path, names):
    src = open(os.path.join(REF, path)).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []  # @allow_command only whitelists the dotted name (syft/generic/utils.py:27-31)
            out[node.name] = ast.get_source_segment(src, node).split("\n")
            # drop decorator lines kept by get_source_segment
            lines = [ln for ln in out[node.name] if not ln.startswith("@")]
            out[node.name] = "\n".join(lines)
    assert set(out) == set(names), (names, list(out))
    return out


def rand_i64(gen, shape):
    return torch.randint(-(2 ** 63), 2 ** 63 - 1, shape, dtype=torch.int64, generator=gen)


def gen_preconv():
    fns = extract("syft/frameworks/torch/nn/functional.py", ["_pre_conv", "_post_conv"])
    ns = {"torch": torch}
    for s in fns.values():
        exec(s, ns)
    g = torch.Generator().manual_seed(42)
    cases = [  # (B, C, H, W, Cout, k, stride, pad)
        (1, 3, 11, 11, 4, 7, 2, 3),
        (2, 4, 9, 9, 5, 3, 1, 1),
        (1, 4, 10, 10, 6, 3, 2, 1),
        (2, 8, 8, 8, 3, 1, 2, 0),
        (1, 2, 5, 7, 2, 3, 1, 0),
    ]
    out = {"cases": np.array(cases, dtype=np.int64)}
    for i, (B, C, H, W, Co, k, s, p) in enumerate(cases):
        x = rand_i64(g, (B, C, H, W))
        w = rand_i64(g, (Co, C, k, k))
        im, wr, b_, co_, ho_, wo_ = ns["_pre_conv"](x, w, None, s, p, 1, 1)
        res = torch.matmul(im, wr)

        class _T(torch.Tensor):
            pass

        post = ns["_post_conv"](None, res, b_, co_, ho_, wo_)
        out[f"x{i}"] = x.numpy()
        out[f"w{i}"] = w.numpy()
        out[f"im{i}"] = im.numpy()
        out[f"wr{i}"] = wr.numpy()
        out[f"post{i}"] = post.numpy()
    np.savez_compressed(os.path.join(HERE, "ring_preconv.npz"), **out)


def gen_spdz():
    fns = extract("syft/frameworks/torch/mpc/spdz.py", ["spdz_mask", "spdz_compute", "triple_mat_mul", "slice"])

    class Store:
        def __init__(self):
            self.t = None

        def get_keys(self, **kw):
            return self.t

    class Owner:
        def __init__(self):
            self.crypto_store = Store()

    class Pool:  # multiprocessing.Pool() stand-in: starmap in-process (spdz.py:108-109)
        def starmap(self, f, args):
            return [f(*a) for a in args]

        def close(self):
            pass

    mp = types.SimpleNamespace(Pool=Pool)
    import math

    ns = {"th": torch, "math": math, "multiprocessing": mp, "N_CORES": 4}
    # ``slice`` in the reference sets ``x_slice.owner``; plain tensors accept attributes.
    for s in fns.values():
        exec(s, ns)

    g = torch.Generator().manual_seed(43)
    out = {}
    shapes = [("matmul", (2, 6, 5), (5, 3)), ("matmul", (1, 7, 4), (4, 2)), ("mul", (4,), (6, 4)), ("mul", (6, 4), (4,)),
              ("mul", (3, 5), (3, 5))]
    out["n"] = np.array(len(shapes))
    for i, (op, sx, sy_) in enumerate(shapes):
        a = rand_i64(g, sx)
        b = rand_i64(g, sy_)
        c = torch.matmul(a, b) if op == "matmul" else a * b
        a0, b0 = rand_i64(g, sx), rand_i64(g, sy_)
        c0 = rand_i64(g, tuple(c.shape))
        tri = [(a0, b0, c0), (a - a0, b - b0, c - c0)]
        x, y = rand_i64(g, sx), rand_i64(g, sy_)
        x0, y0 = rand_i64(g, sx), rand_i64(g, sy_)
        xs, ys = [x0, x - x0], [y0, y - y0]
        owners = [Owner(), Owner()]
        ds, es = [], []
        for j in range(2):
            owners[j].crypto_store.t = tri[j]
            for t in tri[j]:
                t.owner = owners[j]  # hooked tensors always carry .owner in PySyft
            xs[j].owner = owners[j]
            d, e = ns["spdz_mask"](xs[j], ys[j], op, "long", torch.int64, 2 ** 64)
            ds.append(d)
            es.append(e)
        delta, eps = ds[0] + ds[1], es[0] + es[1]
        zs = []
        for j in range(2):
            delta.owner = owners[j]
            zs.append(ns["spdz_compute"](j, delta, eps, op, "long", torch.int64, 2 ** 64))
        assert torch.equal(zs[0] + zs[1], torch.matmul(x, y) if op == "matmul" else x * y)
        out[f"op{i}"] = np.array(op)
        for j in range(2):
            out[f"x{i}_{j}"] = xs[j].numpy()
            out[f"y{i}_{j}"] = ys[j].numpy()
            out[f"a{i}_{j}"], out[f"b{i}_{j}"], out[f"c{i}_{j}"] = (t.numpy() for t in tri[j])
            out[f"d{i}_{j}"] = ds[j].numpy()
            out[f"e{i}_{j}"] = es[j].numpy()
            out[f"z{i}_{j}"] = zs[j].numpy()
    np.savez_compressed(os.path.join(HERE, "ring_spdz.npz"), **out)


def gen_newton():
    """Run the reference's newton branch on a tiny exact-rational stand-in to pin its control
    flow (number of iterations, operand order, where / C happens)."""
    src = open(os.path.join(REF, "syft/frameworks/torch/tensors/interpreters/precision.py")).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "reciprocal":
            fn = ast.get_source_segment(src, node)
    assert fn is not None
    import textwrap

    ns = {}
    exec(textwrap.dedent(fn), ns)

    class Rec:
        """records the op trace"""

        trace = []

        def __init__(self, name):
            self.name = name

        def _new(self, op, other):
            n = f"t{len(Rec.trace)}"
            Rec.trace.append((n, op, self.name, getattr(other, "name", repr(other))))
            return Rec(n)

        def __mul__(self, o):
            return self._new("mul", o)

        def __rsub__(self, o):
            return self._new("rsub", o)

        def __truediv__(self, o):
            return self._new("div", o)

    r = ns["reciprocal"](Rec("v"), method="newton")
    tr = np.array(["|".join(t) for t in Rec.trace])
    np.savez_compressed(os.path.join(HERE, "ring_newton.npz"), trace=tr, result=np.array(r.name))


def gen_train():
    sys.modules.setdefault("syft", types.SimpleNamespace(Plan=object))
    sys.path.insert(0, os.path.join(REF, "torchlib"))
    import importlib

    models = importlib.import_module("models")
    torch.manual_seed(42)
    m = models.resnet18(pretrained=False, in_channels=3, num_classes=3, adptpool=False, input_size=64, pooling="max")
    sd = m.state_dict()
    keys = np.array(list(sd.keys()))
    shapes = np.array([",".join(map(str, v.shape)) for v in sd.values()])
    ws = np.array([v.double().sum().item() for v in sd.values()])  # before the forward updates running stats
    g = torch.Generator().manual_seed(42)
    x = torch.randn(4, 3, 64, 64, generator=g)
    y = torch.randint(0, 3, (4,), generator=g)
    m.train()
    out = m(x)
    loss = torch.nn.CrossEntropyLoss()(out, y)
    loss.backward()
    gn = np.array([p.grad.double().norm().item() for p in m.parameters()])
    gs = np.array([p.grad.double().sum().item() for p in m.parameters()])
    np.savez_compressed(
        os.path.join(HERE, "train_ref.npz"), keys=keys, shapes=shapes, logits=out.detach().numpy(),
        loss=np.array(loss.item()), grad_norms=gn, grad_sums=gs, weight_sums=ws,
        rm=m.bn1.running_mean.numpy(), rv=m.bn1.running_var.numpy(),
    )


def gen_fss():
    """Execute the reference's fss.py (DIF keygen / eval, the H PRG, compress / uncompress) with stand-ins for the modules
    that cannot be imported here: ``shaloop`` -> hashlib (SHA-256 / SHA-512 of each 16-byte row), PySyft runtime -> stubs."""
    import hashlib
    import importlib.util

    def sha_loop(name):
        def f(x, out):
            for i in range(x.shape[0]):
                out[i] = np.frombuffer(getattr(hashlib, name)(x[i].tobytes()).digest(), dtype=np.uint8)
        return f

    sys.modules["shaloop"] = types.SimpleNamespace(sha256_loop_func=sha_loop("sha256"), sha512_loop_func=sha_loop("sha512"))
    sy = types.ModuleType("syft")
    sy.exceptions = types.ModuleType("syft.exceptions")
    sy.exceptions.EmptyCryptoPrimitiveStoreError = type("EmptyCryptoPrimitiveStoreError", (Exception,), {})
    sy.workers = types.ModuleType("syft.workers")
    sy.workers.websocket_client = types.ModuleType("syft.workers.websocket_client")
    sy.workers.websocket_client.WebsocketClientWorker = type("WebsocketClientWorker", (), {})
    sy.generic = types.ModuleType("syft.generic")
    sy.generic.utils = types.ModuleType("syft.generic.utils")
    sy.generic.utils.allow_command = lambda f: f
    sy.generic.utils.remote = lambda f, location=None: f
    for name, mod in (("syft", sy), ("syft.exceptions", sy.exceptions), ("syft.workers", sy.workers),
                      ("syft.workers.websocket_client", sy.workers.websocket_client), ("syft.generic", sy.generic),
                      ("syft.generic.utils", sy.generic.utils)):
        sys.modules[name] = mod
    path = os.path.join(REF, "syft/frameworks/torch/mpc/fss.py")
    src = open(path).read()
    # numpy >= 2 (NEP 50) refuses ``(-1) ** uint64_array``; under the reference's numpy 1.x the python int was value-cast and
    # the power promoted to float64.  Executing the same expression with an explicit float64 base reproduces that result.
    src = src.replace("(-1) **", "np.float64(-1) **")
    fss = types.ModuleType("ref_fss")
    fss.__file__ = path
    exec(compile(src, path, "exec"), fss.__dict__)
    np.random.seed(1234)
    n_values = 24
    with np.errstate(all="ignore"):
        alpha, s00, s01, *rest = fss.DIF.keygen(n_values)
    cw, leaf = rest[:-1], rest[-1]
    out = {"alpha": alpha, "s00": s00, "s01": s01, "leaf": leaf}
    for i, c in enumerate(cw):
        out[f"tauL{i}"], out[f"tL{i}"], out[f"tauR{i}"], out[f"tR{i}"] = (np.asarray(v) for v in c[:4])
        out[f"sig{i}"], out[f"s{i}"] = c[4], c[5]
    # inputs around alpha: equal, +-1, extremes, random
    x = np.concatenate([alpha[:6], alpha[6:12] + 1, alpha[12:18] - 1, np.array([0, 2 ** 32 - 1, 1, 2 ** 31, 2 ** 31 - 1, 12345], dtype=np.uint64)])
    x = x % (2 ** 32)
    with np.errstate(all="ignore"):
        e0 = fss.DIF.eval(0, x.copy(), s00, *cw, leaf)
        e1 = fss.DIF.eval(1, x.copy(), s01, *cw, leaf)
    assert np.array_equal((e0 + e1), (x <= alpha).astype(np.int64)), (e0 + e1, x <= alpha)
    out["x"], out["e0"], out["e1"] = x, e0, e1
    seed = np.stack([s00[:, :4], s01[:, :4]])[0]
    out["H_in"], out["H_out"] = seed.copy(), fss.H(seed.copy()).copy()
    np.savez_compressed(os.path.join(HERE, "fss_dif.npz"), **out)


def gen_pool():
    """_pre_pool / _post_pool executed on int64 tensors, and the control flow of _pool2d's max branch (which slices are
    compared, in which order, with which shapes) recorded by running the reference's own ``_pool2d`` over a stand-in
    AdditiveSharingTensor that computes on the reconstructed value."""
    fns = extract("syft/frameworks/torch/nn/functional.py", ["_pre_pool", "_post_pool", "_pool2d"])
    trace = []

    class Loc:
        def __init__(self, id):
            self.id = id

    locs = [Loc("alice"), Loc("bob")]

    class AST:
        """two additive shares; comparisons and products are evaluated on the reconstructed value and re-shared"""

        def __init__(self, child=None, **kw):
            self.child = child
            self.locations = locs

        def get_class_attributes(self):
            return {}

        @property
        def shape(self):
            return self.child["alice"].shape

        def _plain(self):
            return self.child["alice"] + self.child["bob"]

        def _share(self, v):
            r = torch.full_like(v, 12345)
            return AST({"alice": r, "bob": v - r})

        def __getitem__(self, idx):
            trace.append("slice|" + repr(idx).replace("slice(None, None, None)", ":").replace(" ", ""))
            return AST({k: v[idx] for k, v in self.child.items()})

        def __add__(self, o):
            trace.append(f"add|{tuple(self.shape)}")
            return AST({k: self.child[k] + o.child[k] for k in self.child})

        def __sub__(self, o):
            trace.append(f"sub|{tuple(self.shape)}")
            return AST({k: self.child[k] - o.child[k] for k in self.child})

        def __ge__(self, o):
            trace.append(f"ge|{tuple(self.shape)}")
            return self._share((self._plain() >= o._plain()).long())

        def __mul__(self, o):
            trace.append(f"mul|{tuple(self.shape)}")
            return self._share(self._plain() * o._plain())

    class FPT:
        def __init__(self, **kw):
            self.child = None

        def get_class_attributes(self):
            return {}

        def on(self, t, wrap=False):
            self.child = t
            return self

    def remote(f, location=None):
        def g(*a, return_value=False, return_arity=1, **kw):
            return f(*a, **kw)
        return g

    sy = types.SimpleNamespace(AdditiveSharingTensor=AST, FixedPrecisionTensor=FPT)
    ns = {"torch": torch, "sy": sy, "remote": remote}
    for s_ in fns.values():
        exec(s_, ns)
    g = torch.Generator().manual_seed(44)
    out = {}
    cases = [(1, 3, 8, 8, 3, 2, 1), (2, 2, 7, 7, 3, 2, 1), (1, 2, 6, 6, 2, 2, 0)]  # (B, C, H, W, k, stride, pad)
    out["cases"] = np.array(cases, dtype=np.int64)
    for i, (B, C, H, W, k, st, pd) in enumerate(cases):
        x = torch.randint(-1000, 1000, (B, C, H, W), dtype=torch.int64, generator=g)
        im, *params = ns["_pre_pool"](x, k, st, pd, 1)
        out[f"x{i}"], out[f"im{i}"] = x.numpy(), im.numpy()
        out[f"post{i}"] = ns["_post_pool"](im.sum(-1), *params).numpy()
        x0 = rand_i64(g, (B, C, H, W))
        fp = FPT().on(AST({"alice": x0, "bob": x - x0}))
        del trace[:]
        res = ns["_pool2d"](fp, kernel_size=k, stride=st, padding=pd, dilation=1, mode="max")
        got = res.child._plain()
        want = torch.nn.functional.max_pool2d(torch.nn.functional.pad(x.double(), (pd, pd, pd, pd)), k, st).long()
        assert torch.equal(got, want)
        out[f"max{i}"] = got.numpy()
        out[f"trace{i}"] = np.array(list(trace))
    np.savez_compressed(os.path.join(HERE, "ring_pool.npz"), **out)


if __name__ == "__main__":
    gen_pool()
    gen_preconv()
    gen_spdz()
    gen_newton()
    gen_train()
    gen_fss()
    print("golden fixtures written to", HERE)
