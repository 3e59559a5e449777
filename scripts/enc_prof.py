"""One traced + one preprocessed encrypted forward at 224x224 for ncu (launch list / --set full).  GPU only.
The profiled region (cudaProfilerStart/Stop) is the ONLINE forward of the second image."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from primia_b200 import ring

size = int(sys.argv[1]) if len(sys.argv) > 1 else 224
dev = "cuda:0"
torch.manual_seed(42)
parties = [ring.Party("model_owner", dev), ring.Party("data_owner", dev)]
prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", dev), seed=42)
net = ring.EncryptedResNet18.from_state_dict(torchvision.models.resnet18(num_classes=3).state_dict(), parties, prov, 10, 16, input_size=size)
img = torch.randn(1, 3, size, size) * 0.1
net.trace(net.share_input(img))
net.preprocess(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out, pred = net.predict(img.to(dev))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pred", pred.item())
