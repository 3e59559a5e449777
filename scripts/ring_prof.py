import sys, time, torch
sys.path.insert(0, '.')
from primia_b200 import ring
from primia_b200.ring.resnet import SharedLinearLayers
from torch.profiler import profile, ProfilerActivity
parties = [ring.Party("model_owner", "cuda:0"), ring.Party("data_owner", "cuda:0")]
prov = ring.spdz.TripleProvider(ring.Party("crypto_provider", "cuda:0"), seed=42)
net = SharedLinearLayers(parties, prov, 10, 16)
xs = net.make_inputs(1)
for _ in range(2):
    net.preprocess(1, 1); net.forward(xs)
torch.cuda.synchronize()
net.preprocess(1, 1); torch.cuda.synchronize()
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    net.forward(xs); torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - t0) * 1e3)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
