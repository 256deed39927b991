#!/usr/bin/env python3
"""Time the K3 tile variants (DFTBP_B200_K3_VARIANT) against cuBLAS DGEMM.  Development aid."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftbplus_b200.api import HotPath, new_dense  # noqa: E402

dev = torch.device("cuda:0")
hp = HotPath(0)
hp.use_torch_stream()
f64 = dict(dtype=torch.float64, device=dev)
res = {}
SIZES = [(31744, 14336)] if "--full" in sys.argv else [(7936, 3584), (16384, 8192)]
# 0: default (TMA-fed 64x64, 4 stages, 3 CTAs / SM), 20: the cp.async kernel of round 1, 31-34: TMA-fed tile variants
VARIANTS = [int(v) for v in os.environ.get("K3_VARIANTS", "0,20,31,32,33,34").split(",")]
for n, nocc in SIZES:
    C = new_dense(n)
    C.copy_(torch.randn(n, n, **f64) / n ** 0.5)
    fill = torch.zeros(n, **f64)
    fill[:nocc] = 2.0
    dm = new_dense(n)
    ref = None
    for v in VARIANTS:
        os.environ["DFTBP_B200_K3_VARIANT"] = str(v)
        hp.getDensityMatrix(C, fill, dm=dm)
        torch.cuda.synchronize()
        if ref is None:
            ref = torch.tril(2.0 * (C[:, :nocc] @ C[:, :nocc].t()))
        err = (dm - ref).abs().max().item() / ref.abs().max().item()
        hp.timer_enable(True)
        for _ in range(3):
            hp.getDensityMatrix(C, fill, dm=dm)
        ms = hp.timer_read()["densityMatrix"][0] / 3
        res["n%d_v%d" % (n, v)] = {"ms": ms, "tflops": n * n * nocc / ms / 1e9, "relerr": err}
        print(n, v, res["n%d_v%d" % (n, v)], flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Cc = C[:, :nocc].contiguous()
    torch.matmul(Cc, Cc.t())
    a.record()
    for _ in range(3):
        torch.matmul(Cc, Cc.t())
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 3
    res["n%d_cublas_gemm" % n] = {"ms": ms, "tflops_full_gemm": 2.0 * n * n * nocc / ms / 1e9}
    print(n, "cublas", res["n%d_cublas_gemm" % n], flush=True)
    del C, dm, ref, Cc
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_k3.json"), "w"), indent=1)
