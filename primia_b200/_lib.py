"""ctypes binding of ``libprimia_b200.so`` (the C ABI declared in ``include/primia_b200.h``).

The header is the single source of truth: prototypes are parsed from it, so every declared
symbol must be exported by the library (checked at import and by ``tests/test_abi.py``).
There is NO fallback: if the library is missing or a call fails, a Python exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "primia_b200.h")
LIB_PATH = os.path.join(HERE, "libprimia_b200.so")


class PrimiaError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """pm_conv_t"""

    _fields_ = [(n, ctypes.c_int) for n in ("B", "H", "W", "C", "K", "R", "S", "stride", "pad", "Ho", "Wo")]


class WCvt(ctypes.Structure):
    """pm_wcvt_t"""

    _fields_ = [("w", ctypes.c_void_p), ("w_fwd", ctypes.c_void_p), ("w_dgrad", ctypes.c_void_p), ("K", ctypes.c_int),
                ("C", ctypes.c_int), ("RS", ctypes.c_int), ("Cpad", ctypes.c_int)]


class NewtonJob(ctypes.Structure):
    """pm_newton_job_t"""

    _fields_ = [(n, ctypes.c_void_p) for n in ("v0", "v1", "a0", "b0", "c0", "a1", "b1", "c1", "k0", "k1", "x0", "x1")] + [
        ("C", ctypes.c_int)]


class NewtonP2PJob(ctypes.Structure):
    """pm_newton_p2p_job_t"""

    _fields_ = [(n, ctypes.c_void_p) for n in ("v", "a", "b", "c", "k", "x")] + [("C", ctypes.c_int)]


class AugSample(ctypes.Structure):
    """pm_aug_sample_t"""

    _fields_ = [("src_off", ctypes.c_int64), ("Hs", ctypes.c_int32), ("Ws", ctypes.c_int32), ("C", ctypes.c_int32),
                ("tab_off", ctypes.c_int32), ("fix", ctypes.c_int32 * 6), ("cy", ctypes.c_int32), ("cx", ctypes.c_int32),
                ("flip", ctypes.c_int32), ("area2", ctypes.c_int32), ("noise_sigma", ctypes.c_float), ("reserved", ctypes.c_uint32),
                ("noise_seed", ctypes.c_uint64)]


_SCALARS = {
    "int": ctypes.c_int,
    "size_t": ctypes.c_size_t,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "pm_stream_t": ctypes.c_void_p,
}


def _ctype(decl: str):
    decl = decl.strip()
    decl = re.sub(r"\s+", " ", decl)
    if "*" in decl:
        return ctypes.c_void_p
    base = decl.replace("const ", "").split(" ")[0]
    return _SCALARS[base]


def parse_header(path: str = HEADER):
    """Return {name: (restype, [argtypes])} for every ``pm_*`` prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"(const char\*|int|size_t)\s+(pm_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.groups()
        args = args.strip()
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        restype = {"int": ctypes.c_int, "size_t": ctypes.c_size_t, "const char*": ctypes.c_char_p}[ret]
        protos[name] = (restype, argtypes)
    return protos


_lib = None
MISSING = []
PROTOS = parse_header()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PrimiaError(
                f"{LIB_PATH} is missing: build it with `python -m primia_b200.build` "
                "(there is no CPU / PyTorch fallback for the hot paths)"
            )
        _lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOS.items():
            try:
                fn = getattr(_lib, name)
            except AttributeError:
                MISSING.append(name)  # tests/test_abi.py requires this list to be empty
                continue
            fn.restype = restype
            fn.argtypes = argtypes
    return _lib


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise PrimiaError("primia_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise PrimiaError("tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# kernels launched per entry point when it is not exactly one (used for the gpu_launches claim in bench.py)
KERNELS_PER_CALL = {
    "pm_conv_wgrad_f32": 2, "pm_linear_ce_f32": 2, "pm_stem_pool_bn_bwd_bf16": 2,
    "pm_spdz_combine_matmul_i64": 1, "pm_ring_gemm2_tc_i64": 5, "pm_ring_gemm_planes_i64": 2, "pm_bn_newton_p2p_i64": 2, "pm_dp_add_noise_f32": 2,
}
launch_counter = 0


def call(name: str, *args):
    """Invoke ``pm_<name>``; non-zero status raises PrimiaError with the library's message."""
    try:
        fn = getattr(lib(), name)
    except AttributeError as e:
        raise PrimiaError(f"symbol {name} declared in primia_b200.h is not exported by {LIB_PATH}") from e
    global launch_counter
    launch_counter += KERNELS_PER_CALL.get(name, 1)
    rc = fn(*args)
    if rc != 0:
        msg = lib().pm_last_error().decode()
        raise PrimiaError(f"{name} failed (status {rc}): {msg}")
    return rc
