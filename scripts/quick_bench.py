import sys, torch
sys.path.insert(0, '.')
from primia_b200.train import ResNet18Engine
from oracle import train_oracle as O
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
e = ResNet18Engine(B, 3, 3, 224, "max", "cuda:0", mode)
e.load_state_dict(O.resnet18().state_dict())
x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.randint(0, 3, (B,), device="cuda")
for _ in range(3): e.train_step(x, y)
torch.cuda.synchronize()
s = torch.cuda.Event(enable_timing=True); t = torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): e.train_step(x, y)
t.record(); torch.cuda.synchronize()
ms = s.elapsed_time(t) / 10
print(f"{mode} step B={B}: {ms:.2f} ms -> {B/ms*1000:.0f} img/s loss={e.loss.item():.4f}")
if len(sys.argv) > 3:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): e.train_step(x, y)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
