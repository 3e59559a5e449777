"""Pin the FSS (DIF) oracle against keys and outputs produced by executing the reference's own fss.py
(tests/golden/make_golden.py::gen_fss; shaloop replaced by hashlib SHA-512)."""
import os

import numpy as np

from oracle import fss_oracle as F

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fss_dif.npz"))


def golden_key():
    n = G["alpha"].shape[0]
    bits = np.stack([np.stack([G[f"tauL{i}"], G[f"tL{i}"], G[f"tauR{i}"], G[f"tR{i}"]]).astype(np.uint8).reshape(4, n) for i in range(32)])
    return {"alpha": G["alpha"], "s0": np.stack([G["s00"], G["s01"]]), "bits": bits,
            "sigma_cw": np.stack([G[f"sig{i}"] for i in range(32)]), "s_cw": np.stack([G[f"s{i}"] for i in range(32)]),
            "leaf": G["leaf"]}


def test_prg_H_matches_reference():
    assert np.array_equal(F.H(G["H_in"]), G["H_out"])


def test_keygen_reproduces_reference_keys_from_the_same_randomness():
    ref = golden_key()
    key = F.dif_keygen(G["alpha"], ref["s0"])
    assert np.array_equal(key["bits"], ref["bits"])
    assert np.array_equal(key["sigma_cw"], ref["sigma_cw"])
    assert np.array_equal(key["s_cw"], ref["s_cw"])
    assert np.array_equal(key["leaf"], ref["leaf"])


def test_eval_matches_reference_shares():
    key = golden_key()
    assert np.array_equal(F.dif_eval(0, G["x"], key), G["e0"])
    assert np.array_equal(F.dif_eval(1, G["x"], key), G["e1"])
    assert np.array_equal(G["e0"] + G["e1"], (G["x"] <= G["alpha"]).astype(np.int64))


def test_le_protocol_on_shares():
    rng = np.random.default_rng(7)
    n = 64
    alpha = rng.integers(0, 2 ** 32, n, dtype=np.uint64)
    seeds = rng.integers(0, 2 ** 63, (2, 2, n), dtype=np.uint64)
    key = F.dif_keygen(alpha, seeds)
    a_sh = F.split_alpha(alpha, rng.integers(0, 2 ** 32, n, dtype=np.uint64))
    x1 = rng.integers(-1000, 1000, n).astype(np.int64)
    x2 = rng.integers(-1000, 1000, n).astype(np.int64)
    x2[:8] = x1[:8]
    sh = lambda v: (lambda r: [r, v - r])(rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64))
    out = F.fss_le(sh(x1), sh(x2), key, a_sh)
    assert np.array_equal(out[0] + out[1], (x1 <= x2).astype(np.int64))


# ------------------------------------------------------------------------------------------------ the C twin of the oracle
def test_c_oracle_matches_reference_fixtures_and_numpy_oracle():
    """oracle/fss_oracle_c.c (used for the full-size 224 x 224 encrypted forward) against the SAME reference-generated keys and
    shares, and against the numpy oracle on fresh randomness"""
    from oracle import fss_oracle_c as C

    assert np.array_equal(C.H(G["H_in"]), G["H_out"])
    ref = golden_key()
    key = C.dif_keygen(G["alpha"], ref["s0"])
    for k in ("bits", "sigma_cw", "s_cw", "leaf"):
        assert np.array_equal(key[k], ref[k]), k
    assert np.array_equal(C.dif_eval(0, G["x"], ref), G["e0"]) and np.array_equal(C.dif_eval(1, G["x"], ref), G["e1"])
    rng = np.random.default_rng(11)
    n = 777
    alpha = rng.integers(0, 2 ** 32, n, dtype=np.uint64)
    seeds = rng.integers(0, 2 ** 63, (2, 2, n), dtype=np.uint64)
    kp, kc = F.dif_keygen(alpha, seeds), C.dif_keygen(alpha, seeds)
    for k in ("bits", "sigma_cw", "s_cw", "leaf"):
        assert np.array_equal(kp[k], kc[k]), k
    x = rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)
    for b in range(2):
        assert np.array_equal(F.dif_eval(b, x, kp), C.dif_eval(b, x, kc))
    a_sh = F.split_alpha(alpha, rng.integers(0, 2 ** 32, n, dtype=np.uint64))
    x1, x2 = rng.integers(-1000, 1000, n).astype(np.int64), rng.integers(-1000, 1000, n).astype(np.int64)
    sh = lambda v: (lambda r: [r, v - r])(rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64))
    s1, s2 = sh(x1), sh(x2)
    op, oc = F.fss_le(s1, s2, kp, a_sh), C.fss_le(s1, s2, kc, a_sh)
    assert all(np.array_equal(a, b) for a, b in zip(op, oc))
    assert np.array_equal(oc[0] + oc[1], (x1 <= x2).astype(np.int64))
