#!/bin/bash
# One GPU-box session: parity tests, bench, ncu evidence.  Usage: tools/gpu_session.sh [tag]
# (gpurun copies back at most 64 MiB: the all-kernel ncu capture goes without --import-source)
TAG=${1:-r01}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== probe (all kernels, device resident)"
timeout 600 python tools/gpu_probe.py fec4096 --dense --no-k3 2>&1 | tail -1 | tee gpurun_out/probe_$TAG.log
echo "== A/B: every species pair through the looped evaluator (instruction-fetch experiment, DESIGN section 8)"
DFTBP_B200_FORCE_GENERIC=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k1_ or k5_ or derivative or stress" 2>&1 | tail -3 | tee gpurun_out/pytest_generic_$TAG.log
DFTBP_B200_FORCE_GENERIC=1 timeout 600 python tools/gpu_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/probe_generic_$TAG.log
echo "== neighbour list (N2) probe"; timeout 300 python tools/n2_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/n2_probe_$TAG.log
echo "== bench"; timeout 900 python bench.py 2>&1 | tail -3 | tee gpurun_out/bench_$TAG.log
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_$TAG.log
echo "== ncu launch list (bench, pair path only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-dense --no-n2 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "== ncu full, every kernel of the library (one probe run, one launch per kernel and call)"
timeout 1200 ncu --set full --clock-control none -k regex:'k1_|k2_|k4_|k5_|k6_|ps_|nl_|pop_' -c 60 -o gpurun_out/prof_all_$TAG python tools/gpu_probe.py fec4096 --dense --no-k3 --once > gpurun_out/ncu_all_$TAG.log 2>&1
# summarise on the box; keep the report itself only if it fits the 64 MiB copy-back limit
python tools/ncu_summary.py gpurun_out/prof_all_$TAG.ncu-rep gpurun_out/ncu_all_kernels_$TAG.txt > /dev/null 2>&1
if [ $(stat -c %s gpurun_out/prof_all_$TAG.ncu-rep 2>/dev/null || echo 0) -gt 45000000 ]; then rm -f gpurun_out/prof_all_$TAG.ncu-rep; fi
ls -la gpurun_out | tail -30
du -sh gpurun_out
