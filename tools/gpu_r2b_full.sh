#!/bin/bash
# Full session of the second part of round 2: all GPU tests, smoke, bench (both arms), launch list, ncu of the new kernels.
TAG=${1:-r2be}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/pytest_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_ref_$TAG.log | cut -c1-300
echo "== ncu launch list (bench, pair path only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --pair-only > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "== ncu full: new kernels"
timeout 900 ncu --set full --clock-control none -k regex:'k3_dm_real_tma|n4_|geo_' -c 16 -o gpurun_out/prof_new_$TAG python tools/ncu_r2b_probe.py > gpurun_out/ncu_new_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_new_$TAG.ncu-rep gpurun_out/ncu_new_kernels_$TAG.txt > /dev/null 2>&1
if [ $(stat -c %s gpurun_out/prof_new_$TAG.ncu-rep 2>/dev/null || echo 0) -gt 45000000 ]; then rm -f gpurun_out/prof_new_$TAG.ncu-rep; fi
ls -la gpurun_out | tail -12
