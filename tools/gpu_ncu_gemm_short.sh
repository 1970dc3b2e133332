#!/bin/bash
# Short ncu pass after a GEMM change: launch list of one warmed-up forward + `--set full` of the first 14 `<256>` GEMM launches
# (every distinct shape of the forward appears among them).   gpurun --timeout 900 -- 'bash tools/gpu_ncu_gemm_short.sh [tag]'
TAG=${1:-k}
OUT=gpurun_out
mkdir -p $OUT
W="python bench.py --ncu-window --no-cpu --train-steps 0"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv $W > $OUT/${TAG}_launches.log 2>&1
timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled \
    -k 'regex:gemm_bf16_tcgen05_kernel<\(int\)256' -c 14 -f -o $OUT/${TAG}_gemm_full $W > $OUT/${TAG}_gemm_full.log 2>&1
ncu -i $OUT/${TAG}_gemm_full.ncu-rep --page raw --csv > $OUT/${TAG}_gemm_full_raw.csv 2>/dev/null
rm -f $OUT/${TAG}_gemm_full.ncu-rep
ls -la $OUT | tail -5
