"""The round-end scripts around the device neighbour list (tools/n2_probe.py, which bench.py runs in a
child process, and the last step of __graft_entry__.smoke()) executed on the CPU with a stand-in for the
library context whose list comes from the host run of the product's kernels -- so that a Python slip in
those scripts cannot pass for a GPU failure.  No GPU, no claims about the device: flow and keys only."""
import ctypes
import importlib.util
import json
import os
import sys

import numpy as np
import pytest

from helpers import oracle, case, ROOT
import test_neighbour_list_host as T


class _FakeHotPath:
    """Just enough of dftbplus_b200.api.HotPath for the two scripts."""
    S_REF = None

    def __init__(self, device=0, rank=0, nranks=1):
        self.geo, self.kernel_launches = None, 0
        self._lib = ctypes.CDLL(T.SO)

    def set_orbitals(self, ang_shells):
        self.nos = [sum(2 * l + 1 for l in s) for s in ang_shells]

    def set_sk_table(self, *a, **k):
        pass

    def timer_enable(self, on=True):
        pass

    def timer_read(self):
        return {"neighbourList": (1.0, 4)}

    def updateNeighbourListAndSpecies(self, coord0, species0, cutoff, latvec=None, sk_cutoff=None, adopt=True,
                                      fetch=True, symmetric=False):
        cell, rcell = (np.zeros((1, 3)), np.zeros((1, 3))) if latvec is None else oracle.cell_translations(latvec, cutoff)
        g = T.host_list(self._lib, coord0, species0, self.nos, cell, rcell, cutoff,
                        cutoff if sk_cutoff is None else sk_cutoff, symmetric=symmetric)
        g.update(cellVec=cell, rCellVec=rcell)
        self.kernel_launches += 19
        if not fetch:
            g = {k: g[k] for k in ("nAtom", "nAllAtom", "mNeigh", "sparseSize", "cellVec")}
        if adopt:
            self.geo = g
        return g

    def buildS(self, over=None):
        return np.array(self.S_REF, copy=True)

    def set_geometry(self, geo):
        self.geo = geo

    # N4 stand-ins (flow only): the values come from the checker itself
    def invRPeriodic(self, coord, r_lat, g_lat, alpha, volume, r_cutoff, invr=None):
        from oracle import coulomb_model as cm
        return cm.invr_periodic(coord, dict(rLat=r_lat, gLat=g_lat, alpha=alpha, volume=volume, maxR=r_cutoff))

    def coulombShift(self, invr, dq, shift=None):
        from oracle import coulomb_model as cm
        return cm.coulomb_shift(invr, dq)

    def close(self):
        pass


@pytest.fixture()
def fake(monkeypatch):
    T.build_host_so()
    w, sk, geo, model, selfegy = case("fec128")
    _FakeHotPath.S_REF = oracle.build_s(model, oracle.neighbour_geometry(model, w["coord0"], w["species0"], w["latvec"]))
    import dftbplus_b200.api as api
    monkeypatch.setattr(api, "HotPath", _FakeHotPath)
    return _FakeHotPath


def test_n2_probe_flow(fake, monkeypatch, capsys):
    import torch
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    spec = importlib.util.spec_from_file_location("n2_probe_dry", os.path.join(ROOT, "tools", "n2_probe.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(mod, "HotPath", fake)
    monkeypatch.setattr(sys, "argv", ["n2_probe.py", "fec128", "--oracle"])
    mod.main()
    out = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert out["arrays_equal_host_producer"] and out["arrays_equal_oracle"]
    assert out["S_on_adopted_list_equals_S_on_host_geometry"]
    assert out["cpu_baseline"]["kind"] == "port" and out["cpu_baseline"]["value"] > 0
    for k in ("update_wall_ms", "update_device_ms", "update_fetch_wall_ms", "update_adopt_wall_ms",
              "set_geometry_from_host_wall_ms", "kernel_launches_per_update", "candidate_images", "pairs"):
        assert k in out


def test_smoke_neighbour_list_flow(fake, capsys):
    spec = importlib.util.spec_from_file_location("graft_entry_dry", os.path.join(ROOT, "__graft_entry__.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod._smoke_neighbour_list()
    text = capsys.readouterr().out
    assert "smoke N2 device neighbour list ok" in text and "ERROR" not in text and "MISMATCH" not in text
