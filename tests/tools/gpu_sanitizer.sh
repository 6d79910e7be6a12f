#!/bin/bash
# compute-sanitizer over small instances of every kernel family (tests/tools/gpu_sanitize_cases.py); logs -> gpurun_out/.
# Usage (on the GPU box): bash scripts/gpu_sanitizer.sh [tools...]   default: racecheck synccheck memcheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tools=${@:-racecheck synccheck memcheck}
for tool in $tools; do
  for c in k1 k2 k4 k3_64 k3_128 k3_mixed; do
    log=gpurun_out/r02_sanitizer_${tool}_${c}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tests/tools/gpu_sanitize_cases.py $c > $log 2>&1
    echo "$tool $c rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|case .* ok' $log | tr '\n' ' ')"
  done
done
