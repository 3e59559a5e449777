#!/bin/bash
# 4-GPU box: multi-GPU tests (P2P Newton, 3-GPU placement eager + graph, NCCL FedAvg), then path E in the 3-GPU placement
OUT=gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/r02_multigpu_tests.log 2>&1; tail -12 $OUT/r02_multigpu_tests.log
timeout 400 python bench.py --path E --steps 5 > $OUT/r02_bench_pathE_3gpu.json 2> $OUT/r02_bench_pathE.err; tail -3 $OUT/r02_bench_pathE.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_pathE_3gpu.json"))
for p in d["placements"]:
    print(p.get("placement", "")[:40], "online", p.get("online_ms"), "offline", p.get("offline_ms"), "launches", p.get("gpu_launches"), p.get("error"))
print("value", d["value"], "cpu", d.get("cpu_baseline", {}).get("value"), "roofline frac", d["roofline"]["frac"], "linear", d["linear_layers"]["online_ms"])
PY
