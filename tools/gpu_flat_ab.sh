#!/bin/bash
# Span-pipeline kernels (flat_stream.cuh): parity tests, then A/B timings against the warp-per-block kernels.
TAG=${1:-r2w}
mkdir -p gpurun_out
echo "== pytest k5 / k6 / scc"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "k5 or k6 or scc or stress" 2>&1 | tail -15 | tee gpurun_out/pytest_flat_$TAG.log
for f in 1 0; do
echo "== probe flat=$f"
DFTBP_B200_K5_FLAT=$f DFTBP_B200_K6_FLAT=$f timeout 300 python tools/gpu_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/probe_flat${f}_$TAG.log
done
