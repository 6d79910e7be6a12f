#!/bin/bash
# ncu evidence for the recurrent path (cfg3): launch list + full capture of K3.  Run under gpurun from the repo root.
set -u
TAG=${1:-r02}
WL=${2:-cfg3}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-cli --extra none"
$CMD > gpurun_out/${TAG}_${WL}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_${WL}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_${WL}_launches.csv $CMD > gpurun_out/${TAG}_${WL}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rnn_seq_mc_kernel -s 12 -c 1 -f -o gpurun_out/${TAG}_${WL}_rnn_mc $CMD > gpurun_out/${TAG}_${WL}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:seq_wide2_kernel -s 12 -c 1 -f -o gpurun_out/${TAG}_${WL}_rnn_wide $CMD > gpurun_out/${TAG}_${WL}_ncu3.log 2>&1
ls -la gpurun_out/ | grep ${TAG}_${WL}
