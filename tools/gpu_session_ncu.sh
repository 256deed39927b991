#!/bin/bash
# ncu evidence only: launch list of the bench's pair path + one --set full capture per kernel (summarised here).
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== ncu launch list (bench, pair path only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-dense --no-n2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "== ncu full, every kernel of the library"
timeout 1200 ncu --set full --clock-control none -k regex:'k1_|k2_|k4_|k5_|k6_|ps_' -c 40 -o gpurun_out/prof_all_$TAG python tools/gpu_probe.py fec4096 --dense --no-k3 --once > gpurun_out/ncu_all_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_all_$TAG.ncu-rep gpurun_out/ncu_all_kernels_$TAG.txt > /dev/null 2>&1
if [ $(stat -c %s gpurun_out/prof_all_$TAG.ncu-rep 2>/dev/null || echo 0) -gt 45000000 ]; then rm -f gpurun_out/prof_all_$TAG.ncu-rep; fi
ls -la gpurun_out | tail; du -sh gpurun_out
