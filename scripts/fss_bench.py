"""Time the FSS kernels (DIF keygen / eval) at the stem-activation size; prints hashes/s.  GPU only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from primia_b200 import ring

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 112 * 112
keys = ring.fss.build_fss_keys(n, dev, 1, 1)
x = ring.ops.random_i64((n,), 2, 2, dev)


def timeit(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


alpha = ring.ops.random_i64((n,), 3, 1, dev) & 0xFFFFFFFF
seeds = ring.ops.random_i64((2, 2, n), 3, 2, dev) & 0x7FFFFFFFFFFFFFFF
t_kg = timeit(lambda: ring.fss.dif_keygen(alpha, seeds))
win = keys[0].window(n)
t_ev = timeit(lambda: ring.fss.dif_eval(0, x, win))
seed = seeds[0].contiguous()
t_prg = timeit(lambda: ring.fss.prg_sha512(seed))
print(json.dumps({"n": n, "keygen_ms": t_kg, "eval_ms": t_ev, "prg_ms": t_prg,
                  "keygen_Ghash_s": 64 * n / t_kg / 1e6, "eval_Ghash_s": 32 * n / t_ev / 1e6, "prg_Ghash_s": n / t_prg / 1e6}))
