"""isolated bandwidth of the BN kernels at the stem / layer1 / layer3 sizes (B=64), CUDA-event timed"""
import ctypes, sys, torch
sys.path.insert(0, '.')
from primia_b200._lib import call, ptr, stream
dev = "cuda:0"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us
for (P, C) in [(802816, 64), (200704, 64), (50176, 128), (12544, 256), (3136, 512)]:
    x = torch.randn(P, C, device=dev).bfloat16(); dy = torch.randn(P, C, device=dev).bfloat16()
    y = torch.relu(x); dx = torch.empty_like(x); g = torch.empty_like(x)
    mean = torch.zeros(C, device=dev); invstd = torch.ones(C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev); dg = torch.empty(C, device=dev); db = torch.empty(C, device=dev)
    stats = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    mb = P * C * 2 / 1e6
    t_copy = timeit(lambda: dx.copy_(x))
    t_apply = timeit(lambda: call("pm_bn_bwd_apply_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(sums), P, C, ptr(dx), ptr(dg), ptr(db), stream()))
    t_red = timeit(lambda: call("pm_bn_bwd_reduce_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), P, C, ptr(sums), ptr(g), stream()))
    t_fwd = timeit(lambda: call("pm_bn_fwd_fused_bf16", ptr(x), ptr(stats), P, C, ctypes.c_float(1e-5), ctypes.c_float(0.1), ptr(gamma), ptr(beta), None, 1, ptr(dx), ptr(mean), ptr(invstd), None, None, stream()))
    t_stat = timeit(lambda: call("pm_bn_stats_bf16", ptr(x), P, C, ptr(stats), stream()))
    print(f"P={P:7d} C={C:3d} tensor={mb:7.1f}MB | copy {t_copy:7.1f}us {2*mb/t_copy:6.2f}TB/s | bwd_apply {t_apply:7.1f}us {4*mb/t_apply:5.2f} | bwd_reduce {t_red:7.1f}us {4*mb/t_red:5.2f} | fwd_apply {t_fwd:7.1f}us {2*mb/t_fwd:5.2f} | stats {t_stat:7.1f}us {mb/t_stat:5.2f}")
