"""isolated timing of the BN kernels at every ResNet-18 BN size (B=64), 20 launches captured in a CUDA graph so that
host launch overhead does not pollute the small sizes"""
import ctypes, sys, torch
sys.path.insert(0, '.')
from primia_b200._lib import call, lib, ptr, stream
dev = "cuda:0"
lib().pm_bn_bwd_fused_ws_doubles.restype = ctypes.c_size_t
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3  # us
for (P, C) in [(802816, 64), (200704, 64), (50176, 128), (12544, 256), (3136, 512)]:
    x = torch.randn(P, C, device=dev).bfloat16(); dy = torch.randn(P, C, device=dev).bfloat16()
    y = torch.relu(x); dx = torch.empty_like(x); g = torch.empty_like(x)
    mean = torch.zeros(C, device=dev); invstd = torch.ones(C, device=dev); gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    dg = torch.empty(C, device=dev); db = torch.empty(C, device=dev)
    stats = torch.ones(2 * C, dtype=torch.float64, device=dev) * P
    ws = torch.zeros(int(lib().pm_bn_bwd_fused_ws_doubles(C)), dtype=torch.float64, device=dev)
    mb = P * C * 2 / 1e6
    t_copy = timeit(lambda: dx.copy_(x))
    t_bwd = timeit(lambda: call("pm_bn_bwd_fused_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), P, C, ptr(ws), None, ptr(dx), ptr(dg), ptr(db), stream()))
    t_bwdg = timeit(lambda: call("pm_bn_bwd_fused_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), P, C, ptr(ws), ptr(g), ptr(dx), ptr(dg), ptr(db), stream()))
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    def two_kernel():
        call("pm_bn_bwd_reduce_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), P, C, ptr(sums), None, stream())
        call("pm_bn_bwd_apply_bf16", ptr(dy), ptr(y), ptr(x), ptr(mean), ptr(invstd), ptr(gamma), ptr(sums), P, C, ptr(dx), ptr(dg), ptr(db), stream())
    t_two = timeit(two_kernel)
    t_fwd = timeit(lambda: call("pm_bn_fwd_fused_bf16", ptr(x), ptr(stats), P, C, ctypes.c_float(1e-5), ctypes.c_float(0.1), ptr(gamma), ptr(beta), None, 1, ptr(dx), ptr(mean), ptr(invstd), None, None, stream()))
    t_fwdr = timeit(lambda: call("pm_bn_fwd_fused_bf16", ptr(x), ptr(stats), P, C, ctypes.c_float(1e-5), ctypes.c_float(0.1), ptr(gamma), ptr(beta), ptr(dy), 1, ptr(dx), ptr(mean), ptr(invstd), None, None, stream()))
    print(f"P={P:7d} C={C:3d} tensor={mb:6.1f}MB | copy {t_copy:6.1f}us | bwd_fused {t_bwd:6.1f}us ({7*mb/t_bwd:4.2f}TB/s) +g_out {t_bwdg:6.1f}us two-kernel {t_two:6.1f}us | fwd {t_fwd:6.1f}us ({2*mb/t_fwd:4.2f}) +res {t_fwdr:6.1f}us")
