#!/usr/bin/env python3
"""Timing and self-check of the device Coulomb 1/R matrix (second piece) on one workload, as
ONE JSON line: Ewald matrix (device-resident, CUDA events), shift product, and -- with --oracle -- the CPU
restatement on all host threads as this row's baseline plus the element-wise comparison."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftbplus_b200 import workloads  # noqa: E402
from dftbplus_b200.api import HotPath  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "water2046"
    from dftbplus_b200 import coulomb as pc          # host mirror of TCoulomb%updateLatVecs: the inputs of the device routine
    w = workloads.BY_NAME[name]()
    s = pc.ewald_setup(w["latvec"])
    n = len(w["coord0"])
    out = {"workload": name, "nAtom": n, "alpha": s["alpha"], "maxREwald": s["maxR"], "nRLat": int(len(s["rLat"])),
           "nG": int(len(s["gLat"]))}
    hp = HotPath(0)
    hp.use_torch_stream()
    coord = torch.from_numpy(np.ascontiguousarray(w["coord0"])).cuda()
    rlat = torch.from_numpy(s["rLat"]).cuda()
    glat = torch.from_numpy(s["gLat"]).cuda()
    invr = torch.zeros((n, n), dtype=torch.float64, device="cuda").t()
    dq = torch.from_numpy(np.random.default_rng(1).normal(size=n)).cuda()
    sh = torch.zeros_like(dq)

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    out["invr_periodic_ms"] = timed(lambda: hp.invRPeriodic(coord, rlat, glat, s["alpha"], s["volume"], s["maxR"], invr=invr))
    out["coulomb_shift_ms"] = timed(lambda: hp.coulombShift(invr, dq, shift=sh), reps=20)
    out["coulomb_shift_GBs"] = 8.0 * n * (n + 1) / 2 / out["coulomb_shift_ms"] / 1e6
    out["invr_cluster_ms"] = timed(lambda: hp.invRCluster(coord, invr=invr))
    hp.invRPeriodic(coord, rlat, glat, s["alpha"], s["volume"], s["maxR"], invr=invr)
    hp.synchronize()
    if "--oracle" in sys.argv:
        # the CPU baseline leg and checker (the only use of oracle/ here)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle
        from oracle import coulomb_model as cm
        t0 = time.perf_counter()
        ref = cm.invr_periodic(w["coord0"], s)
        out["cpu_baseline"] = {"value": (time.perf_counter() - t0) * 1e3, "unit": "ms", "cores": oracle.num_threads(),
                               "kind": "port", "sample": "the full workload once: invRPeriodic restated in C, OpenMP over atoms like the reference"}
        got = invr.cpu().numpy()
        out["max_rel_diff_vs_oracle"] = float(np.abs(got - ref).max() / np.abs(ref).max())
    hp.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
