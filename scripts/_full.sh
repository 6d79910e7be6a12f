cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^    \|^$" | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02_bench_default_v3.json 2> gpurun_out/r02_bench_default_v3.err; tail -c 3000 gpurun_out/r02_bench_default_v3.json
