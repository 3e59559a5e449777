import sys, copy, torch
sys.path.insert(0, '.')
from oracle import train_oracle as O
from primia_b200.train import ResNet18Engine
DEV = "cuda:0"
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
def run(mode, B, size):
    torch.manual_seed(42)
    m = O.ResNet18(input_size=size)
    eng = ResNet18Engine(B, 3, 3, size, "max", DEV, mode)
    eng.load_state_dict(m.state_dict())
    g = torch.Generator().manual_seed(42)
    x = torch.randn(B, 3, size, size, generator=g); y = torch.randint(0, 3, (B,), generator=g)
    m.train(); out = m(x); loss = torch.nn.functional.cross_entropy(out, y); loss.backward()
    eng.forward(x.to(DEV)); l = eng.loss_and_backward(y.to(DEV)); torch.cuda.synchronize()
    print(f"== {mode} B={B} size={size}: loss gpu {l.item():.6f} cpu {loss.item():.6f} logits rel {rel(eng.logits, out.detach()):.2e}")
    gd = eng.grad_dict()
    for n, p in m.named_parameters():
        a, b = gd[n].double().cpu().flatten(), p.grad.double().flatten()
        cos = (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()
        print(f"  {n:34s} rel {rel(gd[n], p.grad):.2e} cos {cos:.5f} |g| {b.norm().item():.2e}")
for spec in sys.argv[1:]:
    mode, B, size = spec.split(",")
    run(mode, int(B), int(size))
