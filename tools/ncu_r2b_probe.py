#!/usr/bin/env python3
"""One launch of every kernel added in the second part of round 2, for `ncu --set full -k regex:...`:
the TMA-fed K3 (n = 16384), the Coulomb kernels (fec4096: structure factors, the Gram matrix through K3,
real-space Ewald, symmetric product), the short-range gamma kernels and the geo_* kernels of adopt."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dftbplus_b200 import coulomb as pc, workloads  # noqa: E402
from dftbplus_b200.api import HotPath, new_dense  # noqa: E402

hp = HotPath(0)
hp.use_torch_stream()
f64 = dict(dtype=torch.float64, device="cuda")
n, nocc = 16384, 8192
C = new_dense(n)
C.copy_(torch.randn(n, n, **f64) / n ** 0.5)
fill = torch.zeros(n, **f64)
fill[:nocc] = 2.0
dm = new_dense(n)
hp.getDensityMatrix(C, fill, dm=dm)
torch.cuda.synchronize()
del C, dm
w = workloads.BY_NAME["fec4096"]()
sk, geo, _ = workloads.prepare(w)
sk.upload(hp)
hubm = pc.unique_hubbard([[workloads.SPECIES[nm][2][0]] * len(sk.ang_shells[i]) for i, nm in enumerate(w["names"])],
                         [len(a) for a in sk.ang_shells])
cuts = pc.short_gamma_cutoffs(hubm)
ew = pc.ewald_setup(w["latvec"])
cutoff = max(sk.cutoff(), float(cuts.max()))
hp.updateNeighbourListAndSpecies(w["coord0"], w["species0"], cutoff, w["latvec"], sk_cutoff=sk.cutoff(), adopt=True, fetch=False)
hp.shortGammaUpdate(hubm["nHubbU"], hubm["uniqHubbU"], cuts)
na = len(w["coord0"])
coord = torch.from_numpy(np.ascontiguousarray(w["coord0"])).cuda()
invr = new_dense(na)
hp.invRPeriodic(coord, torch.from_numpy(ew["rLat"]).cuda(), torch.from_numpy(ew["gLat"]).cuda(), ew["alpha"], ew["volume"],
                ew["maxR"], invr=invr)
dq = torch.randn(na, **f64)
hp.coulombShift(invr, dq)
dqs = torch.zeros((na, hubm["mShell"]), **f64)
dqs[:, 0] = dq
hp.shortGammaShifts(hubm["iHubbU"], dqs, shift_shell=torch.zeros_like(dqs))
hp.invRCluster(coord, invr=invr)
torch.cuda.synchronize()
hp.close()
print("ok")
