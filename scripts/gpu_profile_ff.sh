#!/bin/bash
# ncu evidence for the feed-forward path (cfg2): launch list + full captures of K2 / K4 / K1.
# Run under gpurun from the repo root; outputs land in gpurun_out/ (copy the summaries into profiles/).
set -u
TAG=${1:-r02}
WL=${2:-cfg2}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-cli --extra none"
$CMD > gpurun_out/${TAG}_${WL}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_${WL}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 133 -c 7 -f -o gpurun_out/${TAG}_${WL}_gemm $CMD > gpurun_out/${TAG}_${WL}_ncu2.log 2>&1
# (the head of a single net runs inside gemm_logsoftmax_kernel; NNAM_FUSED_HEAD=0 brings head_fast_kernel back)
ncu --set full --clock-control none --import-source on -k regex:splice_transform -s 20 -c 2 -f -o gpurun_out/${TAG}_${WL}_splice $CMD > gpurun_out/${TAG}_${WL}_ncu4.log 2>&1
ls -la gpurun_out/
