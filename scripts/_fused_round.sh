cd /root/repo
for a in "65536 512 0 0" "65536 512 1 0" "65536 512 0 1" "65536 2048 0 0" "65536 2048 1 0"; do timeout 120 python scripts/gpu_fused_one.py $a; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_logsoftmax -c 1 -o gpurun_out/fused_k512 python scripts/gpu_fused_one.py 65536 512 0 0 > gpurun_out/fused_ncu.log 2>&1; tail -3 gpurun_out/fused_ncu.log
