#!/bin/bash
# Session of the TMA-fed K3 and the device-side adopt: focused parity tests, K3 variants at config size, N2 probe.
TAG=${1:-r2ba}
mkdir -p gpurun_out
echo "== pytest (K3 + adopt)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_neighbour_list_gpu.py tests/test_zy_config_size_gpu.py -m gpu -x -q -k "k3 or adopt or density or neighbour or md or chain" 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
echo "== K3 variants at n = 31744"
timeout 300 python tools/tune_k3.py --full 2>&1 | tail -12 | tee gpurun_out/tune_k3_$TAG.log
echo "== N2 probe"
timeout 200 python tools/n2_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/n2_probe_$TAG.log
DFTBP_B200_ADOPT_HOST=1 timeout 200 python tools/n2_probe.py fec4096 2>&1 | tail -1 | tee gpurun_out/n2_probe_hostadopt_$TAG.log
