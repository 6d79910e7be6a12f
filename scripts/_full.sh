cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^    \|^$" | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for w in cfg1 cfg3 cfg3t cfg4 cfg4g cfg5 cfg5b; do
  timeout 400 python bench.py --workload $w --no-cli --no-strong --extra "" --steps 5 --warmup 3 > gpurun_out/r02_bench_${w}_v3.json 2>/dev/null; python -c "
import sys, json
r=json.loads(open('gpurun_out/r02_bench_${w}_v3.json').read().strip().splitlines()[-1]); print('$w', round(r['value']/1e6,2), 'M/s', round(r['ms_per_step'],2), 'ms  e2e', round(r['e2e']['value']/1e6,2), round(r['roofline']['frac'],3), r['parity']['argmax_agreement_raw'], r['parity']['max_abs'])"
done
