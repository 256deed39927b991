#!/bin/bash
# Short GPU session: parity tests, probe A/B of a tuning switch.
TAG=${1:-r01e}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
for b in 2 1 0; do
echo "== probe, K5 force-weights batch $b"
DFTBP_B200_K5W_BATCH=$b timeout 600 python tools/gpu_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/probe_k5w${b}_$TAG.log
done
