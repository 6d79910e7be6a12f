cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^    \|^$" | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default_v5.json 2> gpurun_out/r02_bench_default_v5.err; python -c "
import json
r=json.loads(open('gpurun_out/r02_bench_default_v5.json').read().strip().splitlines()[-1])
print('cfg2', r['value'], r['ms_per_step'], 'e2e', r['e2e']['value'], r['e2e']['transfer'][:12], r['e2e']['probed_compact_fractions'], r['roofline']['frac'], r['parity']['argmax_agreement_raw'], r['cli']['seconds'], r['clocks']['sm_mhz'])
for x in r['extra']: print(x['workload'][:5], x['value'], x['e2e'])"
