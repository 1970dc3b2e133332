timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k gemm 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_kernels_gpu.py 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('eval', round(j['value'],1), round(j['ms_per_step'],3), round(j['e2e']['value'],1), round(j['roofline']['frac'],3), j['clocks']['sm_mhz'], j['train_step']['value'])"
