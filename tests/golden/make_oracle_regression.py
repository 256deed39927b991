#!/usr/bin/env python3
"""Regression vectors of the oracle itself (NOT of the Fortran reference, which cannot be built in
this image): outputs of oracle/ for two tiny workloads, so that a later change to the restatement
cannot drift unnoticed.  Usage: python tests/golden/make_oracle_regression.py  (rewrites the .npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from helpers import oracle, case  # noqa: E402


def vectors():
    out = {}
    for name in ("h2o", "si2"):
        w, sk, geo, model, selfegy = case(name)
        rng = np.random.default_rng(31)
        h0, s = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
        rho = s * rng.normal(size=geo["sparseSize"])
        erho = s * rng.normal(size=geo["sparseSize"])
        sh = 0.1 * rng.normal(size=(1, geo["nAtom"], model.mOrb, model.mOrb))
        sh = 0.5 * (sh + sh.transpose(0, 1, 3, 2))
        out[name + "_h0"], out[name + "_s"] = h0, s
        out[name + "_rho"], out[name + "_erho"], out[name + "_shift"] = rho, erho, sh
        out[name + "_q"] = oracle.mulliken_per_orbital(model, geo, s, rho)
        out[name + "_ham_scc"] = oracle.add_shift(model, geo, h0[None, :], s, sh)
        out[name + "_deriv"] = oracle.derivative_shift(model, geo, rho, erho, shift=sh)
        out[name + "_deriv_rich"] = oracle.derivative_shift(model, geo, rho, erho, diff_type=2)
        if w["latvec"] is not None:
            out[name + "_stress"] = oracle.stress(model, geo, rho, erho, abs(np.linalg.det(w["latvec"])), shift=sh)
            kp = w["kpoints"][3]
            out[name + "_hk"] = oracle.unpack_hs_cplx(geo, h0, kp)
        # neighbour list / sparse descriptor / fold / per-atom populations
        cutoff = sk.cutoff()
        rcell = np.zeros((1, 3)) if w["latvec"] is None else oracle.cell_translations(w["latvec"], cutoff)[1]
        nl = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, 0.8 * cutoff)
        for k in ("coords", "img2CentCell", "iCellVec", "nNeighbour", "nNeighbourSK", "iNeighbour", "neighDist2", "iPair"):
            out[name + "_nl_" + k] = np.asarray(nl[k], dtype=np.float64)
        out[name + "_q_atom"] = oracle.mulliken_per_atom(model, geo, s, rho)
        out[name + "_q_onsite"] = oracle.onsite_population(geo, rho)
        if w["latvec"] is not None:
            far = w["coord0"] + np.array([[3.0, -2.0, 5.0], [-1.0, 4.0, 0.0]]) @ np.asarray(w["latvec"]) + 0.123
            out[name + "_folded"] = oracle.fold_coords_to_cell(far, w["latvec"])
        w1, sk1, geo1, model1, selfegy1 = case(name, 1)
        out[name + "_s_old_interp"] = oracle.build_s(model1, geo1)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "oracle_regression.npz"), **vectors())
    print("wrote oracle_regression.npz")
