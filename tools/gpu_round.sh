#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list of the same command, ncu --set full captures of
# the two dominant kernels.  Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.json
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --pipeline 0 > $OUT/${TAG}_launches_bench.log 2>&1
  # <256> instantiations only, in launch order per forward: OCR projection, QTV 2 x (qkv, out, up, down) [x3], then the
  # answer transformer's shared qkv, out, up, down ... [bf16]; 47 per forward -> skip one forward + 5, take QTV layer 1
  # and the first bf16 layer
  timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_tcgen05_kernel<\(int\)256' --launch-skip 52 -c 8 \
      -f -o $OUT/${TAG}_gemm_full python bench.py --steps 1 --warmup 1 --no-cpu --pipeline 0 > $OUT/${TAG}_gemm_full.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc --launch-skip 18 -c 2 \
      -f -o $OUT/${TAG}_attn_full python bench.py --steps 1 --warmup 1 --no-cpu --pipeline 0 > $OUT/${TAG}_attn_full.log 2>&1
  for r in gemm attn; do
    ncu -i $OUT/${TAG}_${r}_full.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_full_raw.csv 2>/dev/null
  done
  ls -la $OUT/
  # gpurun copies back at most 64 MiB: the CSV pages matter more than the reports
  if [ $(du -sm $OUT | cut -f1) -gt 56 ]; then rm -f $OUT/${TAG}_gemm_full.ncu-rep; fi
fi
