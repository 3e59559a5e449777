#!/bin/bash
# usage: scripts/bench_overlap.sh N  -- bench.py at N GPUs with the FedAvg overlap on and off
N=${1:-2}
for ov in 1 0; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$ov bench.py --gpus $N --steps 40 --warmup 5 --overlap $ov 2>gpurun_out/bench_ov$ov.err > gpurun_out/bench_n${N}_ov$ov.json
  python - "$ov" "gpurun_out/bench_n${N}_ov$ov.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[2]))
print("overlap", sys.argv[1], "value", round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 4))
PY
done
