#!/bin/bash
# Round-end evidence pass on one B200: the GPU test suite, the bench lines of every workload (eval with cpu_baseline, the
# reference arm, train, m4c, the six stress shapes) -> gpurun_out/<tag>_*.json.     gpurun --timeout 1500 -- 'bash tools/gpu_final.sh [tag]'
TAG=${1:-f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench_eval.json 2> $OUT/${TAG}_bench_eval.err; tail -c 400 $OUT/${TAG}_bench_eval.json; echo
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 300 $OUT/${TAG}_bench_reference.json; echo
timeout 300 python bench.py --workload train --no-cpu > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err
timeout 300 python bench.py --workload m4c --no-cpu --train-steps 0 > $OUT/${TAG}_bench_m4c.json 2> $OUT/${TAG}_bench_m4c.err
for s in "64 15 64" "128 15 32" "64 30 32" "128 30 16" "256 30 8" "256 60 4"; do
  set -- $s
  timeout 300 python bench.py --workload stress --frames $1 --ocr-per-frame $2 --batch $3 --steps 5 --warmup 3 --no-cpu --train-steps 0 > $OUT/${TAG}_stress_$1x$2.json 2> $OUT/${TAG}_stress_$1x$2.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/${TAG}_bench_*.json")+glob.glob("$OUT/${TAG}_stress_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j.get("roofline") or {}
        print(f.split("/")[-1], round(j["value"],2), j["unit"], round(j["ms_per_step"],3), "e2e", round(j["e2e"]["value"],2), r.get("kernel"), r.get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
