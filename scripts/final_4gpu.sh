#!/bin/bash
# round-end measurement set on a 4-GPU box: multi-GPU tests, path E in the 3-GPU placement, path T at N = 4 with / without overlap
OUT=gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -m gpu -q > $OUT/r02_multigpu_tests.log 2>&1; tail -3 $OUT/r02_multigpu_tests.log
timeout 300 python bench.py --path E --steps 5 > $OUT/r02_bench_pathE_3gpu.json 2> $OUT/r02_bench_pathE.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_pathE_3gpu.json"))
for p in d["placements"]:
    print(p.get("placement", "")[:40], "online", p.get("online_ms"), "offline", p.get("offline_ms"), p.get("error"))
print("value", d["value"], "cpu", d.get("cpu_baseline", {}).get("value"), "roofline frac", d["roofline"]["frac"])
PY
scripts/bench_overlap.sh 4
