#!/bin/bash
# round-end measurement set on ONE GPU: whole GPU test suite, both bench arms, the f32 parity mode and C3
OUT=gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q --durations=10 ) > $OUT/r02_gpu_tests.log 2>&1
tail -15 $OUT/r02_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_n1.json 2> $OUT/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r02_bench_reference_arm.json 2> $OUT/r02_bench_ref.err
timeout 300 python bench.py --mode f32 --steps 10 --warmup 3 --no-cpu > $OUT/r02_bench_f32_mode.json 2> $OUT/r02_bench_f32.err
timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu > $OUT/r02_bench_c3_n1.json 2> $OUT/r02_bench_c3.err
timeout 120 python scripts/bn_bench.py > $OUT/r02_bn_bench.txt 2>&1
python - <<'PY'
import json
for f in ("r02_bench_n1", "r02_bench_reference_arm", "r02_bench_f32_mode", "r02_bench_c3_n1"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        e = d.get("encrypted_inference") or {}
        print(f, round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1),
              "| enc online", e.get("online_ms"), "offline", e.get("offline_ms"), "cpu", (e.get("cpu_baseline") or {}).get("value"),
              "| cpu_baseline", (d.get("cpu_baseline") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
    except Exception as exc:
        print(f, "FAILED", exc)
PY
tail -6 $OUT/r02_bn_bench.txt
