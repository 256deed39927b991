#!/bin/bash
# K1 A/B: parity tests on the default kernel, then probe timings per launch configuration.
TAG=${1:-r2b}
mkdir -p gpurun_out
echo "== pytest test_gpu_parity (default K1)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_k1_$TAG.log
for k in ${K1CFGS:-0 1 8}; do
  echo "== K1 kernel cfg $k"
  DFTBP_B200_K1_KERNEL=$k timeout 300 python tools/gpu_probe.py fec4096 --k1-only 2>&1 | tail -1 | tee gpurun_out/probe_k1_${k}_$TAG.log
done
