"""The C-ABI library loads and exports every symbol include/primia_b200.h declares (no compute calls)."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
    """the oracle is test infrastructure: no product module (package, torchlib shim, entry points) imports it"""
    import re

    here = os.path.join(os.path.dirname(__file__), "..")
    files = [os.path.join(here, f) for f in ("train.py", "inference.py", "train_federated.py")]
    for pkg in ("primia_b200", "torchlib"):
        for dp, _dn, fn in os.walk(os.path.join(here, pkg)):
            files += [os.path.join(dp, f) for f in fn if f.endswith(".py")]
    pat = re.compile(r"^\s*(from\s+\.*oracle[\s.]|import\s+oracle|from\s+\S*\s+import\s+.*\boracle\b|__import__\(.oracle)", re.M)
    for f in files:
        src = open(f).read()
        assert not pat.search(src), f
        assert "importlib" not in src or "oracle" not in src, f


def test_every_kernel_launched_with_the_pdl_attribute_waits_on_its_predecessor():
    """common.cuh: a kernel started through pm_launch() may be scheduled while its predecessor drains, so its body must run
    pm_pdl_sync() before touching global memory.  (A fused-head kernel without it read the last activation early: loss off by
    6 % in one graph replay on B200.)  Static check over the sources: every kernel named at a pm_launch() site has the wait."""
    import glob
    import re

    src = "\n".join(open(f).read() for f in sorted(glob.glob(os.path.join(ROOT, "primia_b200", "csrc", "*.cu*"))))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    names = set(re.findall(r"pm_launch\(\s*([A-Za-z_0-9]+)", src))
    for var, kernel in re.findall(r"auto\s+(\w+)\s*=\s*([A-Za-z_0-9]+)\s*<", src):   # `auto kern = conv_halo_kernel<...>`
        if var in names:
            names.discard(var)
            names.add(kernel)
    names -= {"kernel", "void"}  # pm_launch itself
    assert len(names) >= 12, names
    for n in sorted(names):
        m = re.search(r"__global__[^;{]*?\b" + n + r"\s*\([^{;]*\)\s*\{", src, re.S)
        assert m, f"no definition found for {n}"
        i, depth = m.end(), 1
        while depth and i < len(src):
            depth += (src[i] == "{") - (src[i] == "}")
            i += 1
        assert "pm_pdl_sync" in src[m.end():i], f"{n} is launched with the PDL attribute but never waits"


def test_ctypes_mirrors_have_the_layout_the_c_compiler_gives_the_header_structs(tmp_path):
    """include/primia_b200.h is plain C: compile a probe with gcc that prints sizeof / offsetof of every struct a binding has to
    mirror, and hold primia_b200/_lib.py's ctypes.Structure classes to it (a drifted field would silently corrupt descriptors)."""
    import ctypes
    import shutil
    import subprocess

    from primia_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    pairs = {"pm_conv_t": _lib.ConvDesc, "pm_wcvt_t": _lib.WCvt, "pm_newton_job_t": _lib.NewtonJob,
             "pm_newton_p2p_job_t": _lib.NewtonP2PJob, "pm_aug_sample_t": _lib.AugSample}
    lines = []
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _t in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "primia_b200.h"\nint main(void) {\n' + "\n".join(lines) + "\nreturn 0; }\n")
    exe = tmp_path / "probe"
    subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)   # the header is valid C
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, value = line.split()
        cls = pairs[cname]
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == want, (cname, field, value, want)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
