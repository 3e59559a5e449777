#!/bin/bash
# one B200: ring / FSS parity incl. the hoisted weight side, then the path E bench line hoisted vs not
timeout 900 python -m pytest tests/test_ring_gpu.py tests/test_fss_gpu.py tests/test_entrypoints_gpu.py -m gpu -q -x 2>&1 | tail -8
for h in 1 0; do
  echo "== PRIMIA_HOIST_WEIGHT_SIDE=$h"
  PRIMIA_HOIST_WEIGHT_SIDE=$h timeout 400 python bench.py --path E --steps 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); e=d.get('encrypted_inference',d)
print('online', e['online_ms'], 'offline', e['offline_ms'], 'launches', e['gpu_launches'])
l=e['linear_layers']; print('linear online', l['online_ms'], 'offline', l['offline_triple_gen_ms'], 'launches', l.get('online_launches'))"
done
