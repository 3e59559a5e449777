"""Run-to-run determinism of the bf16 forward: same weights, same batch, twice; with the BatchNorm statistics fused into the
conv epilogue (float shared-memory atomics + double global atomics) and with the separate reduce kernel."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import train_oracle as O
from primia_b200.train import ResNet18Engine

B, size = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 96
torch.manual_seed(42)
sd = O.ResNet18(input_size=size).state_dict()
g = torch.Generator().manual_seed(11)
x, y = torch.randn(B, 3, size, size, generator=g).cuda(), torch.randint(0, 3, (B,), generator=g).cuda()
for fuse in (True, False):
    runs = []
    for r in range(3):
        eng = ResNet18Engine(B, 3, 3, size, "max", "cuda:0", "bf16")
        eng.fuse_stats = fuse
        eng.load_state_dict(sd)
        eng.forward(x)
        loss = eng.loss_and_backward(y).item()
        torch.cuda.synchronize()
        runs.append((loss, {k: v.clone() for k, v in eng.act.items()}, {k: v.clone() for k, v in eng.bn_mean.items()},
                     {k: v.clone() for k, v in eng.bn_invstd.items()}, eng.grads.clone()))
    print(f"fuse_stats={fuse}: losses", [r[0] for r in runs])
    a, b = runs[0], runs[1]
    for k in a[1]:
        d = (a[1][k] != b[1][k]).float().mean().item()
        if d > 0:
            print(f"   act {k}: {d:.2e} of elements differ")
    for k in a[2]:
        dm = (a[2][k] != b[2][k]).sum().item(); di = (a[3][k] != b[3][k]).sum().item()
        if dm or di:
            print(f"   bn {k}: mean differs in {dm} channels, invstd in {di}")
    print("   grads rel diff", ((a[4] - b[4]).norm() / b[4].norm()).item())
