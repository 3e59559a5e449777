#!/bin/bash
# path T on one B200: the training-side GPU tests, then the default bench line (and the fused head off, for the A/B)
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_train_bf16_gpu.py tests/test_train_graph_gpu.py tests/test_dp_gpu.py tests/test_entrypoints_gpu.py -m gpu -q 2>&1 | tail -15
for fh in 1 0; do
  echo "== PRIMIA_FUSE_HEAD=$fh"
  PRIMIA_FUSE_HEAD=$fh timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
done
