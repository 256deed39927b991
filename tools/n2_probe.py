#!/usr/bin/env python3
"""Timing and self-check of the device neighbour list on one workload, as ONE JSON
line.  bench.py runs this in a child process at the end of the N = 1 run (own CUDA context, timeout),
so that a problem in this young path can neither hang nor poison the measured one.

Checks without the oracle (always): the arrays are compared with the host producer the bench already uses
(dftbplus_b200/geometry.py; identical unless two neighbours of an atom are equidistant within
1e-10 bohr^2, where that producer orders by image number), and S built on the adopted device list
is compared bit for bit with S built on the host-produced geometry.  With --oracle (bench.py passes it) the
oracle is timed on the same workload as this row's CPU baseline and every array is compared with it."""
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
    name = sys.argv[1] if len(sys.argv) > 1 else "fec4096"
    out = {"workload": name}
    w = workloads.BY_NAME[name]()
    t0 = time.perf_counter()
    sk, geo, _ = workloads.prepare(w)
    out["host_producer_s"] = time.perf_counter() - t0          # k-d tree producer of the harness (not the reference's scan)
    hp = HotPath(0)
    sk.upload(hp)
    cutoff = sk.cutoff()
    coord0, species0, lat = w["coord0"], w["species0"], w["latvec"]

    def update(**kw):
        return hp.updateNeighbourListAndSpecies(coord0, species0, cutoff, lat, **kw)

    def wall(fn, n=5):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / n * 1e3

    g = update(adopt=False, fetch=True)
    out.update(nAtom=int(g["nAtom"]), nAllAtom=int(g["nAllAtom"]), mNeigh=int(g["mNeigh"]),
               sparseSize=int(g["sparseSize"]), pairs=int(g["nNeighbourSK"].sum()),
               candidate_images=int(g["nAtom"] * g["cellVec"].shape[0]))
    keys = ("img2CentCell", "iCellVec", "species", "nNeighbourSK", "iNeighbour", "iPair", "iAtomStart")
    out["sizes_equal_host_producer"] = bool(g["nAllAtom"] == geo["nAllAtom"] and g["sparseSize"] == geo["sparseSize"]
                                            and g["mNeigh"] == geo["mNeigh"])
    out["arrays_equal_host_producer"] = bool(out["sizes_equal_host_producer"]
                                             and all(np.array_equal(g[k], geo[k]) for k in keys))
    out["coords_max_abs_diff"] = float(np.abs(g["coords"] - geo["coords"]).max()) if out["sizes_equal_host_producer"] else None
    hp.timer_enable(True)
    out["update_wall_ms"] = wall(lambda: update(adopt=False, fetch=False))
    t = hp.timer_read()["neighbourList"]
    out["update_device_ms"] = t[0] / max(1, t[1])               # events around the whole pass incl. the two size read-backs
    hp.timer_enable(False)
    out["update_fetch_wall_ms"] = wall(lambda: update(adopt=False, fetch=True), n=3)
    out["update_adopt_wall_ms"] = wall(lambda: update(adopt=True, fetch=False), n=3)
    out["kernel_launches_per_update"] = None
    n0 = hp.kernel_launches
    update(adopt=False, fetch=False)
    out["kernel_launches_per_update"] = int(hp.kernel_launches - n0)
    # chain: S on the adopted device list vs S on the host-produced geometry
    update(adopt=True, fetch=False)
    s_dev = hp.buildS(np.zeros(g["sparseSize"]))
    out["set_geometry_from_host_wall_ms"] = wall(lambda: hp.set_geometry(geo), n=3)
    s_host = hp.buildS(np.zeros(geo["sparseSize"]))
    out["S_on_adopted_list_equals_S_on_host_geometry"] = bool(out["arrays_equal_host_producer"]
                                                              and np.array_equal(s_dev, s_host))
    hp.close()
    if "--oracle" in sys.argv:
        # the checker and the CPU baseline of this row (bench.py's cpu_baseline leg): the oracle's
        # restatement of the reference's serial O(nCellVec nAtom^2) scan + heap sorts + descriptor
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle
        rcell = np.zeros((1, 3)) if lat is None else oracle.cell_translations(lat, cutoff)[1]
        t0 = time.perf_counter()
        ref = oracle.neighbour_list_full(sk.n_orb_species, coord0, species0, rcell, cutoff, cutoff)
        out["cpu_baseline"] = {"value": (time.perf_counter() - t0) * 1e3, "unit": "ms", "cores": 1, "kind": "port",
                               "sample": "the full workload once: updateNeighbourList + getNrOfNeighboursForAll + "
                                         "getSparseDescriptor restated in C (serial, like the reference without MPI)"}
        out["arrays_equal_oracle"] = bool(
            g["nAllAtom"] == ref["nAllAtom"] and g["mNeigh"] == ref["mNeigh"] and g["sparseSize"] == ref["sparseSize"]
            and all(np.array_equal(g[k], ref[k]) for k in keys + ("coords", "nNeighbour", "neighDist2")))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
