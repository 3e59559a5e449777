"""The C-ABI library loads and exports every symbol include/primia_b200.h declares (no compute calls)."""
import os

import pytest


def test_library_exports_every_declared_symbol():
    from primia_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        from primia_b200.build import build

        build()
    _lib.lib()
    assert len(_lib.PROTOS) >= 40
    assert _lib.MISSING == [], f"declared but not exported: {_lib.MISSING}"
    assert _lib.lib().pm_version() >= 100


def test_no_cpu_fallback():
    """CPU tensors are rejected loudly: the product path never computes on the host."""
    import torch

    from primia_b200._lib import PrimiaError
    from primia_b200.ring import ops

    with pytest.raises(PrimiaError):
        ops.trunc_div(torch.zeros(4, dtype=torch.int64), 10)


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(__file__), "..", "primia_b200")
    for dp, _dn, fn in os.walk(root):
        for f in fn:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "").replace("CPU oracle", ""), f
