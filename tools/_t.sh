timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn" 2>&1 | tail -5
timeout 300 python -m pytest tests/test_dropout_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --train-steps 0 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eval', j['value'], j['ms_per_step'], j['e2e']['value']); print({k:(round(v['ms_per_step'],3) if isinstance(v,dict) and 'ms_per_step' in v else v) for k,v in j.get('kernels',{}).items()})"
