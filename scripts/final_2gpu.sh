#!/bin/bash
# 2-GPU box, end of round: the whole GPU suite (incl. the 2-GPU tests), then path T at N = 2 with the overlapped FedAvg
OUT=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/r02_gpu_tests_2gpu.log 2>&1; tail -6 $OUT/r02_gpu_tests_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 40 --warmup 5 > $OUT/r02_bench_n2_overlap.json 2> $OUT/r02_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n2_overlap.json')); print('N=2', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config'].get('fedavg'))"
