#!/bin/bash
TAG=${1:-r2bb}
mkdir -p gpurun_out
echo "== pytest (K3 variants, adopt, coulomb)"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_neighbour_list_gpu.py tests/test_zw_coulomb_gpu.py -m gpu -x -q -k "k3 or adopt or density or coulomb or invr or madelung or device_resident" 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
echo "== K3 variants at n = 31744"
K3_VARIANTS=0,30,33,35,36,37 timeout 300 python tools/tune_k3.py --full 2>&1 | tail -12 | tee gpurun_out/tune_k3_$TAG.log
echo "== Coulomb probe"
timeout 300 python tools/coulomb_probe.py water2046 --oracle 2>&1 | tail -1 | tee gpurun_out/coulomb_water2046_$TAG.log
timeout 300 python tools/coulomb_probe.py fec4096 --oracle 2>&1 | tail -1 | tee gpurun_out/coulomb_fec4096_$TAG.log
