"""End-to-end encrypted inference timing (path E): offline (triples + FSS keys) and online ms/image.  GPU only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from primia_b200 import ring, _lib

size = int(sys.argv[1]) if len(sys.argv) > 1 else 224
pf = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda:0"
torch.manual_seed(42)
import torchvision
m = torchvision.models.resnet18(num_classes=3)
parties = [ring.Party("model_owner", dev), ring.Party("data_owner", dev)]
prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", dev))
net = ring.EncryptedResNet18.from_state_dict(m.state_dict(), parties, prov, 10, pf, input_size=size)
img = torch.randn(1, 3, size, size)
x = net.share_input(img)
t0 = time.perf_counter(); net.trace(x); torch.cuda.synchronize(); t_trace = time.perf_counter() - t0
print("trace (on-demand) s:", t_trace, "schedule entries:", len(net.schedule), flush=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
off, on, launches = [], [], []
for r in range(reps):
    torch.cuda.synchronize()
    e0, e1, e2 = ev(), ev(), ev()
    w0 = time.perf_counter()
    e0.record(); net.preprocess(1); e1.record()
    torch.cuda.synchronize(); w1 = time.perf_counter()
    l0 = _lib.launch_counter
    out = net.forward(net.share_input(img)); e2.record()
    torch.cuda.synchronize(); w2 = time.perf_counter()
    off.append((e0.elapsed_time(e1), (w1 - w0) * 1e3)); on.append((e1.elapsed_time(e2), (w2 - w1) * 1e3))
    launches.append(_lib.launch_counter - l0)
print(json.dumps({"size": size, "pf": pf, "offline_ms(dev,wall)": off, "online_ms(dev,wall)": on, "launches": launches,
                  "store_bytes_left": [p.crypto_store.nbytes() for p in parties], "generated_GB_per_image": prov.generated_bytes / (reps + 1) / 1e9,
                  "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
