"""one eager bf16 training step (B=64) and one encrypted linear-layer forward: target for ncu kernel filters"""
import sys, torch
sys.path.insert(0, '.')
from primia_b200.train import ResNet18Engine
what = sys.argv[1] if len(sys.argv) > 1 else "train"
if what == "train":
    B = 64
    e = ResNet18Engine(B, 3, 3, 224, "max", "cuda:0", "bf16")
    e.init_random(42)
    x = torch.randn(B, 3, 224, 224, device="cuda"); y = torch.randint(0, 3, (B,), device="cuda")
    for _ in range(2): e.train_step(x, y)
    torch.cuda.synchronize()
else:
    from primia_b200 import ring
    from primia_b200.ring.resnet import SharedLinearLayers
    parties = [ring.Party("model_owner", "cuda:0"), ring.Party("data_owner", "cuda:0")]
    prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", "cuda:0"), seed=42)
    net = SharedLinearLayers(parties, prov, 10, 16)
    xs = net.make_inputs(1)
    for _ in range(2):
        net.preprocess(1, 1); net.forward(xs)
    torch.cuda.synchronize()
