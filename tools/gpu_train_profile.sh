#!/bin/bash
# ncu evidence for ONE warmed-up training step (forward x3 passes, backward, clip + Adam, operand repack):
#   (1) the launch list (gpu__time_duration per launch), (2) `--set full` of the tcgen05 attention backward and of the
#   weight-gradient / data-gradient GEMM instantiations.  Results land in gpurun_out/ (scratch).
#   gpurun --timeout 900 -- 'bash tools/gpu_train_profile.sh [tag]'        (about 2 GPU-minutes)
TAG=${1:-t}
OUT=gpurun_out
mkdir -p $OUT
W="python bench.py --workload train --ncu-window --no-cpu --warmup 3"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_train_launches.csv $W > $OUT/${TAG}_train_launches.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:attn_bwd_tc_kernel|attn_bwd_stats_kernel|attn_bwd_dq_store_kernel' -c 12 \
    -f -o $OUT/${TAG}_attn_bwd_full $W > $OUT/${TAG}_attn_bwd_full.log 2>&1
ncu -i $OUT/${TAG}_attn_bwd_full.ncu-rep --page raw --csv > $OUT/${TAG}_attn_bwd_full_raw.csv 2>/dev/null
ls -la $OUT | tail -8
