#!/bin/bash
TAG=${1:-r2bc}
mkdir -p gpurun_out
echo "== pytest (coulomb + short gamma + smoke)"
timeout 900 python -m pytest tests/test_zw_coulomb_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
echo "== K3 sweep of shapes (default kernel)"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/k3_shapes_$TAG.log
import os, sys, torch
sys.path.insert(0, os.getcwd())
from dftbplus_b200.api import HotPath, new_dense
hp = HotPath(0); hp.use_torch_stream()
f64 = dict(dtype=torch.float64, device="cuda")
for n, nocc in [(2004, 1336), (4092, 2728), (4096, 2048), (8192, 4096), (16384, 8192), (31744, 14336), (31744, 1000), (10000, 9999)]:
    C = new_dense(n); C.copy_(torch.randn(n, n, **f64) / n ** 0.5)
    fill = torch.zeros(n, **f64); fill[:nocc] = 2.0
    dm = new_dense(n)
    hp.getDensityMatrix(C, fill, dm=dm); torch.cuda.synchronize()
    hp.timer_enable(True)
    for _ in range(3): hp.getDensityMatrix(C, fill, dm=dm)
    ms = hp.timer_read()["densityMatrix"][0] / 3
    hp.timer_enable(False)
    Cc = C[:, :nocc].contiguous()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.matmul(Cc, Cc.t()); a.record()
    for _ in range(3): torch.matmul(Cc, Cc.t())
    b.record(); b.synchronize()
    g = a.elapsed_time(b) / 3
    print("n %6d nLevels %6d  k3 %9.3f ms %6.2f TFLOP/s   cuBLAS DGEMM (full square) %9.3f ms %6.2f TFLOP/s" % (n, nocc, ms, n*n*nocc/ms/1e9, g, 2.0*n*n*nocc/g/1e9), flush=True)
    del C, dm, Cc
PY
