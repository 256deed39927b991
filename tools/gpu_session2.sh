#!/bin/bash
# Short GPU session: parity tests, K1 launch-shape A/B, ncu of the pair-path and streaming kernels.
TAG=${1:-r01e}
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_$TAG.log
echo "== probe (default: 8 warps x 2 CTAs)"
timeout 600 python tools/gpu_probe.py fec4096 --dense --no-k3 2>&1 | tail -1 | tee gpurun_out/probe_$TAG.log
for cfg in 63 44; do
echo "== probe, K1 cfg $cfg"
DFTBP_B200_K1_CFG=$cfg timeout 600 python tools/gpu_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/probe_k1cfg${cfg}_$TAG.log
done
echo "== ncu full, K1 / K4 / K2 pack"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_pair_build|k4_|k2_pack' -o gpurun_out/prof_$TAG python tools/gpu_probe.py fec4096 --dense --no-k3 --once > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out | tail -8
