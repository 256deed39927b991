#!/bin/bash
# One GPU-box session: parity tests, bench, ncu evidence.  Usage: tools/gpu_session.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench_$TAG.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_$TAG.log
echo "== ncu launch list (bench, pair path only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-dense > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "== ncu full, one launch per kernel"
for K in k1_pair_build k5_derivative_shift k5_force_weights k4_mulliken_orbital k2_unpack k2_pack ps_keys ps_scan ps_scatter; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o gpurun_out/prof_${K}_$TAG python tools/gpu_probe.py fec4096 --dense --no-k3 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_dm_real' -c 2 -o gpurun_out/prof_k3_dm_real_$TAG python tools/gpu_probe.py fec1024 --dense > gpurun_out/ncu_dm_$TAG.log 2>&1
ls -la gpurun_out | tail -30
