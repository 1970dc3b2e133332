#!/bin/bash
# Round-2 GPU pass: parity tests, bench line, the ncu launch list of ONE warmed-up forward, and `ncu --set full`
# captures of (a) the throughput GEMM instantiations, (b) the tcgen05 attention, (c) the full-size memory-bound kernels.
# Everything lands in gpurun_out/ (scratch); the summaries made from it are committed under profiles/.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round2.sh [tag]'      (SKIP_TESTS=1 / SKIP_NCU=1 to skip parts;
#   ONLY_ATTN=1: of the `--set full` captures only the attention one; SKIP_MEM=1: GEMM and attention captures only -- what was
#   re-taken after the cta_group::2 GEMM and the P-in-TMEM attention went in)
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -4 $OUT/${TAG}_pytest.log
fi
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 300 $OUT/${TAG}_bench.json
if [ -z "$SKIP_NCU" ]; then
  W="python bench.py --ncu-window --no-cpu --train-steps 0"
  timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file $OUT/${TAG}_launches.csv $W > $OUT/${TAG}_launches.log 2>&1
  if [ -z "$ONLY_ATTN" ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled \
      -k 'regex:gemm_bf16_tcgen05_kernel<\(int\)256' -c 40 -f -o $OUT/${TAG}_gemm_full $W > $OUT/${TAG}_gemm_full.log 2>&1
  fi
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_tc -c 8 \
      -f -o $OUT/${TAG}_attn_full $W > $OUT/${TAG}_attn_full.log 2>&1
  if [ -z "$ONLY_ATTN$SKIP_MEM" ]; then
  timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled \
      -k 'regex:add_ln_kernel|feat_concat_kernel|ocr_finish_kernel|phoc_build_kernel|split_bf16_kernel|sim_scores_kernel' -c 60 \
      -f -o $OUT/${TAG}_mem_full $W > $OUT/${TAG}_mem_full.log 2>&1
  timeout 600 ncu --profile-from-start off --set full --clock-control none --kernel-name-base demangled \
      -k 'regex:ptr_score_kernel<\(int\)16|attn_dec_kernel<\(int\)4' -c 6 \
      -f -o $OUT/${TAG}_tail_full $W > $OUT/${TAG}_tail_full.log 2>&1
  fi
  for r in gemm attn mem tail; do
    ncu -i $OUT/${TAG}_${r}_full.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_full_raw.csv 2>/dev/null
  done
  ls -la $OUT/ | tail -20
  # gpurun copies back at most 64 MiB: the CSV pages matter more than the reports
  for r in mem gemm tail attn; do
    if [ $(du -sm $OUT | cut -f1) -gt 56 ]; then rm -f $OUT/${TAG}_${r}_full.ncu-rep; fi
  done
fi
