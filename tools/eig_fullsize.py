#!/usr/bin/env python3
"""The generalized eigensolve of one SCC iteration at the full size of a workload (BASELINE config 4:
n = 31744), device resident, through the C ABI: build H0 / S -> unpackHS x 2 -> dftbp_b200_hegv_real.
Prints ONE JSON line.  bench.py runs it in a child process WITHOUT torch: a process that has imported
torch resolves libcusolver.so.11 to the copy bundled with torch (CUDA 12.8), whose cusolverDnXsyevd
rejects n > ~24000; the library itself links the toolkit's cuSOLVER (12.9), which a Fortran host gets too."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftbplus_b200 import workloads  # noqa: E402
from dftbplus_b200.api import HotPath  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "fec4096"
    device = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    assert "torch" not in sys.modules
    rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaSetDevice(device)

    def dmalloc(nbytes):
        p = ctypes.c_void_p()
        rc = rt.cudaMalloc(ctypes.byref(p), nbytes)
        if rc != 0:
            raise MemoryError("cudaMalloc(%d) failed: %d" % (nbytes, rc))
        return p

    w = workloads.BY_NAME[name]()
    sk, geo, _ = workloads.prepare(w)
    hp = HotPath(device)
    sk.upload(hp)
    hp.set_geometry(geo)
    L, h = hp._L, hp._h
    n, ss = int(geo["nOrb"]), int(geo["sparseSize"])
    ham, over = dmalloc(8 * ss), dmalloc(8 * ss)
    H, S, eig = dmalloc(8 * n * n), dmalloc(8 * n * n), dmalloc(8 * n)
    import numpy as np
    selfegy = hp._selfegy(sk.selfegy)
    hp._ck(L.dftbp_b200_build_h0_s(h, ctypes.c_void_p(selfegy.ctypes.data), ham, over))
    out = {"workload": name, "nOrb": n, "solver": "cusolverDnDsygvd" if 2.1 * n * n <= 2.0e9 else
           "cusolverDnXpotrf + 2 x cublasDtrsm + cusolverDnXsyevd + cublasDtrsm (64-bit API)"}
    reps = 1 if n > 16000 else 3
    hp.timer_enable(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        hp._ck(L.dftbp_b200_unpack_hs_real(h, ham, H, n))
        hp._ck(L.dftbp_b200_unpack_hs_real(h, over, S, n))
        hp._ck(L.dftbp_b200_hegv_real(h, H, S, eig, n, n, b"V"))
    hp.synchronize()
    out["wall_s"] = (time.perf_counter() - t0) / reps
    t = hp.timer_read()
    out["ms"] = t["diagonalization"][0] / max(1, t["diagonalization"][1])
    out["unpack_ms"] = t["sparseToDense"][0] / max(1, t["sparseToDense"][1])
    ev = np.zeros(n)
    rt.cudaMemcpy(ctypes.c_void_p(ev.ctypes.data), eig, 8 * n, 2)
    out["eigenvalues_sorted_finite"] = bool(np.all(np.isfinite(ev)) and np.all(np.diff(ev) >= 0.0))
    out["eigenvalue_range"] = [float(ev[0]), float(ev[-1])]
    print(json.dumps(out))
    hp.close()


if __name__ == "__main__":
    main()
