#!/usr/bin/env python3
"""Quick device-resident timings of every kernel (CUDA events on the library's stream) plus the
cuBLAS FP64 GEMM / SYRK rate used as the FP64 tensor-pipe denominator.  Development aid; the judged
numbers come from bench.py."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftbplus_b200 import workloads  # noqa: E402
from dftbplus_b200.api import HotPath, new_dense  # noqa: E402


ONCE = "--once" in sys.argv  # one warm-up + one timed call per kernel: the ncu capture run


def timeit(fn, stream, n=10, warm=3):
    if ONCE:
        n, warm = 1, 1
    for _ in range(warm):
        fn()
    stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        fn()
    b.record(stream)
    b.synchronize()
    return a.elapsed_time(b) / n


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "fec4096"
    dense = "--dense" in sys.argv
    dev = torch.device("cuda:0")
    out = {"workload": name}
    # cuBLAS FP64 rates
    st0 = torch.cuda.current_stream()
    n = 8192
    A = torch.randn(n, n, dtype=torch.float64, device=dev)
    B = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = timeit(lambda: torch.matmul(A, B), st0, n=5, warm=2)
    out["cublas_dgemm_8192_tflops"] = 2 * n ** 3 / ms / 1e9
    del A, B
    t0 = time.time()
    w = workloads.BY_NAME[name]()
    sk, geo, _ = workloads.prepare(w)
    out["prep_s"] = time.time() - t0
    hp = HotPath(0)
    stream = torch.cuda.Stream(device=dev)
    hp.set_stream(stream.cuda_stream)
    sk.upload(hp)
    hp.set_geometry(geo)
    ss, nOrb, nAtom = geo["sparseSize"], geo["nOrb"], geo["nAtom"]
    pairs = hp.num_pairs
    out.update(nAtom=nAtom, nOrb=nOrb, sparseSize=ss, pairs=pairs)
    f64 = dict(dtype=torch.float64, device=dev)
    ham, over = torch.zeros(ss, **f64), torch.zeros(ss, **f64)
    ms = timeit(lambda: hp.buildH0S(sk.selfegy, ham, over), stream)
    blocks = geo["nOrbAtom"][:, None] * 0
    out["k1_ms"] = ms
    out["k1_pairs_per_s"] = pairs / ms * 1e3
    out["k1_GBs"] = 16.0 * ss / ms / 1e6
    if "--k1-only" in sys.argv:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        def flushed():
            flush.zero_()
            hp.buildH0S(sk.selfegy, ham, over)
        with torch.cuda.stream(stream):
            msf = timeit(flushed, stream)
            msz = timeit(lambda: flush.zero_(), stream)
        out["k1_flushed_ms"] = msf - msz
        print(json.dumps(out))
        hp.close()
        return
    rho = over * torch.randn(ss, **f64)
    edm = over * torch.randn(ss, **f64)
    qq = torch.zeros((nAtom, hp.mOrb), **f64)
    ms = timeit(lambda: hp.mulliken(over, rho, qq=qq), stream)
    out["k4_ms"] = ms
    out["k4_GBs"] = 16.0 * ss / ms / 1e6
    deriv = torch.zeros((nAtom, 3), **f64)
    ms = timeit(lambda: hp.derivative_shift(rho, edm, deriv=deriv), stream, n=5)
    out["k5_ms"] = ms
    out["k5_pairs_per_s"] = pairs / ms * 1e3
    # SCC form, 2 spins: diagonal shift blocks (plain SCC / collinear spin) and full blocks (DFTB+U-like)
    rho2 = torch.stack([rho, over * torch.randn(ss, **f64)])
    mo = hp.mOrb
    shd = torch.zeros((2, nAtom, mo, mo), **f64)
    shd[:, :, torch.arange(mo), torch.arange(mo)] = 0.1 * torch.randn((2, nAtom, mo), **f64)
    ms = timeit(lambda: hp.derivative_shift(rho2, edm, shift=shd, deriv=deriv), stream, n=5)
    out["k5_scc_diag_ms"] = ms
    shf = 0.1 * torch.randn((2, nAtom, mo, mo), **f64)
    shf = 0.5 * (shf + shf.transpose(2, 3)).contiguous()
    ms = timeit(lambda: hp.derivative_shift(rho2, edm, shift=shf, deriv=deriv), stream, n=5)
    out["k5_scc_full_ms"] = ms
    ham2 = torch.zeros((2, ss), **f64)
    ms = timeit(lambda: hp.addShift(ham2, over, shd, is_input_zero=True), stream)
    out["k6_addshift_diag_ms"] = ms
    out["k6_GBs"] = 8.0 * ss * 5 / ms / 1e6
    ms = timeit(lambda: hp.addShift(ham2, over, shf, is_input_zero=True), stream, n=5)
    out["k6_addshift_full_ms"] = ms
    if dense:
        sq = new_dense(nOrb)
        ms = timeit(lambda: hp.unpackHS(ham, square=sq), stream, n=5)
        out["k2a_ms"] = ms
        out["k2a_GBs"] = (8.0 * ss + 8.0 * nOrb * nOrb) / ms / 1e6
        ms = timeit(lambda: hp.packHS(rho, sq), stream, n=5)
        out["k2b_ms"] = ms
        out["k2b_GBs"] = 24.0 * ss / ms / 1e6
    if dense and "--no-k3" not in sys.argv:
        C = new_dense(nOrb)
        C.copy_(torch.randn(nOrb, nOrb, **f64) / nOrb ** 0.5)
        nocc = int(sum(sum(sk.occ[s - 1]) for s in w["species0"]) / 2)
        fill = torch.zeros(nOrb, **f64)
        fill[:nocc] = 2.0
        ms = timeit(lambda: hp.getDensityMatrix(C, fill, dm=sq), stream, n=2, warm=1)
        out["k3_ms"] = ms
        out["k3_nocc"] = nocc
        out["k3_tflops"] = nOrb * nOrb * nocc / ms / 1e9
        # cuBLAS dsyrk-equivalent through torch: C[:, :nocc] @ C[:, :nocc].T (full GEMM, 2x flops)
        Cc = C[:, :nocc].contiguous()
        with torch.cuda.stream(stream):
            ms = timeit(lambda: torch.matmul(Cc, Cc.t()), stream, n=2, warm=1)
        out["cublas_gemm_same_shape_ms"] = ms
        out["cublas_gemm_same_shape_tflops"] = 2.0 * nOrb * nOrb * nocc / ms / 1e9
    # N2: neighbour list + sparse descriptor on the device (host wall clock: the call ends with the
    # read-back of the sizes), without and with the copies of the arrays back to the host, against
    # the same geometry through the host producer + set_geometry
    try:
        cutoff = sk.cutoff()
        def wall(fn, n=5):
            n = 1 if ONCE else n
            fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / n * 1e3
        hp.timer_enable(True)
        out["n2_update_wall_ms"] = wall(lambda: hp.updateNeighbourListAndSpecies(
            w["coord0"], w["species0"], cutoff, w["latvec"], adopt=False, fetch=False))
        t = hp.timer_read()["neighbourList"]
        out["n2_update_device_ms"] = t[0] / max(1, t[1])
        hp.timer_enable(False)
        out["n2_update_fetch_wall_ms"] = wall(lambda: hp.updateNeighbourListAndSpecies(
            w["coord0"], w["species0"], cutoff, w["latvec"], adopt=False, fetch=True), n=3)
        out["n2_update_adopt_wall_ms"] = wall(lambda: hp.updateNeighbourListAndSpecies(
            w["coord0"], w["species0"], cutoff, w["latvec"], adopt=True, fetch=False), n=3)
        out["set_geometry_from_host_wall_ms"] = wall(lambda: hp.set_geometry(geo), n=3)
        out["n2_candidate_images"] = int(len(w["coord0"]) * (27 if w["latvec"] is not None else 1))
    except Exception as e:  # keep the kernel numbers above if this young path fails
        out["n2_error"] = repr(e)
    out["launches"] = hp.kernel_launches
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe_%s.json" % name), "w") as fh:
        json.dump(out, fh, indent=1)
    hp.close()


if __name__ == "__main__":
    main()
