for rep in 1 2; do
for s in 24 48 64; do
  T2S_B200_OVERLAP_SMS=$s python bench.py --steps 20 --warmup 5 --no-cpu --train-steps 0 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sms $s eval', round(j['value'],1), round(j['ms_per_step'],3), round(j['e2e']['value'],1), j['clocks']['sm_mhz'])"
done
done
