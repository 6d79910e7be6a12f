cd /root/repo
for w in cfg1 cfg2; do
 for c in 65536 98304 131072; do
  NNAM_CHUNK_ROWS=$c timeout 300 python bench.py --workload $w --no-cpu-baseline --no-cli --no-strong --extra "" --steps 5 --warmup 3 2>/dev/null | python -c "
import sys, json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w chunk=$c', round(r['value']/1e6,2), 'M/s', round(r['ms_per_step'],2), 'ms e2e', round(r['e2e']['value']/1e6,2), round(r['roofline']['frac'],3), r['clocks']['sm_mhz'])"
 done
done
