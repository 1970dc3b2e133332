#!/bin/bash
# The six shapes of the stress sweep (BASELINE configs[4]) back to back; JSON lines land in gpurun_out/f_stress_FxOf.json.
#   gpurun --timeout 900 -- "bash tools/gpu_stress_sweep.sh"
mkdir -p gpurun_out
for s in "64 15 64" "128 15 32" "64 30 32" "128 30 16" "256 30 8" "256 60 4"; do
  set -- $s
  timeout 300 python bench.py --workload stress --frames $1 --ocr-per-frame $2 --batch $3 --steps 5 --warmup 3 --no-cpu --train-steps 0 > gpurun_out/f_stress_$1x$2.json 2> gpurun_out/f_stress_$1x$2.err
  tail -c 200 gpurun_out/f_stress_$1x$2.json | head -c 10 >/dev/null
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f_stress_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); r=j["roofline"]
        print(f, j["value"], j["ms_per_step"], r.get("kernel"), r.get("share"), r.get("frac"))
    except Exception as e: print(f, "ERR", e)
PY

