cd /root/repo
timeout 300 python bench.py --workload cfg2 --no-cpu-baseline --no-cli --no-strong --extra "" --steps 3 --warmup 3 2>/dev/null | python -c "
import sys, json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 (box indicator) e2e', round(r['e2e']['value']/1e6,2), r['e2e']['transfer'][-60:])"
for w in cfg3 cfg4 cfg3t; do
 for t in f16 f32; do
  NNAM_TRANSFER=$t timeout 300 python bench.py --workload $w --no-cpu-baseline --no-cli --no-strong --extra "" --steps 4 --warmup 3 2>/dev/null | python -c "
import sys, json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $t', round(r['value']/1e6,2), 'M/s  e2e', round(r['e2e']['value']/1e6,2), round(r['e2e']['ms_per_step'],1),'ms')"
 done
done
