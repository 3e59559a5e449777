"""ctypes front of ``oracle/fss_oracle_c.c`` with the numpy oracle's signatures (``oracle/fss_oracle.py``) -- TEST
INFRASTRUCTURE ONLY.  ``build()`` compiles it with gcc into ``oracle/_build/`` (git-ignored; travels to the GPU box)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import fss_oracle as _py

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fss_oracle_c.c")
LIB = os.path.join(HERE, "_build", "libfss_oracle.so")
N_BITS = 32
_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def H(seed):
    seed = _c(seed, np.uint64)
    n = seed.shape[1]
    out = np.empty((2, 6, n), np.uint64)
    lib().fss_H(_p(seed), ctypes.c_size_t(n), _p(out))
    return out


def dif_keygen(alpha, seeds):
    alpha, seeds = _c(alpha, np.uint64), _c(seeds, np.uint64)
    n = alpha.shape[0]
    bits, sigma_cw, s_cw = np.empty((N_BITS, 4, n), np.uint8), np.empty((N_BITS, 2, n), np.uint64), np.empty((N_BITS, 2, n), np.uint64)
    leaf = np.empty((N_BITS + 1, n), np.int32)
    lib().fss_dif_keygen(_p(alpha), _p(seeds), ctypes.c_size_t(n), _p(bits), _p(sigma_cw), _p(s_cw), _p(leaf))
    return {"alpha": alpha.copy(), "s0": seeds.copy(), "bits": bits, "sigma_cw": sigma_cw, "s_cw": s_cw, "leaf": leaf}


def dif_eval(b, x_masked, key):
    x = _c(np.asarray(x_masked).reshape(-1), np.int64)
    n = x.shape[0]
    out = np.empty(n, np.int64)
    lib().fss_dif_eval(ctypes.c_int(b), _p(x), _p(_c(key["s0"][b], np.uint64)), _p(_c(key["bits"], np.uint8)),
                       _p(_c(key["sigma_cw"], np.uint64)), _p(_c(key["s_cw"], np.uint64)), _p(_c(key["leaf"], np.int32)),
                       ctypes.c_size_t(n), _p(out))
    return out.reshape(np.asarray(x_masked).shape)


split_alpha = _py.split_alpha


def fss_le(x1_sh, x2_sh, key, alpha_sh):
    """fss_oracle.fss_le with the C evaluation (mask / open in numpy: trivial)"""
    r = [x1_sh[j] - x2_sh[j] + alpha_sh[j].reshape(x1_sh[j].shape) for j in range(2)]
    with np.errstate(over="ignore"):
        masked = (r[0] + r[1]) % (1 << N_BITS)
    return [dif_eval(j, masked, key) for j in range(2)]
