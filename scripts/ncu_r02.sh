#!/bin/bash
# ncu passes of one bf16 training step (B = 64): launch list with DRAM bytes, and --set full of the top kernel families.
# Run on the GPU box (gpurun); outputs CSV under gpurun_out/ (the .ncu-rep stays on the box: too large to bring back).
set -x
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file $OUT/r02_launches_raw.csv python scripts/prof_targets.py train > $OUT/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"stem_pool_bn_bwd|stem_pool_reduce|head_fused|bn_relu_maxpool|bn_bwd_bf16|bn_fwd_bf16|conv_halo|wgrad_halo|wgrad_tma|stem_kernel|conv_tma|dgrad_s2" \
    --launch-skip 98 -c 100 -o /tmp/r02_full python scripts/prof_targets.py train > $OUT/r02_ncu2.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > $OUT/r02_full_raw.csv 2>>$OUT/r02_ncu2.log
ls -la /tmp/r02_full.ncu-rep $OUT/r02_full_raw.csv
