timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload train --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('train', j['value'], j['ms_per_step'], j['e2e']['value'])"
timeout 120 python tools/gemm_bench.py --iters 20 > gpurun_out/g_gemm_shapes_pair2.jsonl 2>&1
T2S_GEMM_PAIR=1 timeout 120 python tools/gemm_bench.py --iters 20 > gpurun_out/g_gemm_shapes_pair1.jsonl 2>&1
