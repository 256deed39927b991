"""Host-side producers of the hot path's inputs (product code) against the oracle restatement of
the reference: neighbour list / sparse descriptor, SK-file reader + getFullTable, k-point mesh."""
import numpy as np
import pytest

from helpers import oracle, case
from oracle import oldskdata
from dftbplus_b200 import geometry, slako, workloads


@pytest.mark.parametrize("name", ["h2o", "ce_o", "si2", "fec128", "si512"])
def test_geometry_matches_reference_neighbour_list(name):
    w, sk, geo, model, selfegy = case(name)
    g2 = oracle.neighbour_geometry(model, w["coord0"], w["species0"], w["latvec"])
    for k in ("nAtom", "nAllAtom", "sparseSize", "nOrb"):
        assert geo[k] == g2[k], k
    for k in ("nNeighbourSK", "img2CentCell", "iCellVec", "species", "iAtomStart", "cellVec"):
        assert np.array_equal(geo[k], g2[k]), k
    assert np.allclose(geo["coords"], g2["coords"], atol=1e-12)
    assert np.allclose(geo["neighDist2"], g2["neighDist2"], atol=1e-9)      # both sorted by distance
    for i in range(geo["nAtom"]):       # same neighbour SET; order inside equal-distance shells is free
        n = geo["nNeighbourSK"][i]
        assert sorted(geo["iNeighbour"][i, :n + 1]) == sorted(g2["iNeighbour"][i, :n + 1])
    # half-list rule and sortedness (periodic.F90:596, 675-688)
    for i in range(geo["nAtom"]):
        n = geo["nNeighbourSK"][i]
        assert np.all(geo["img2CentCell"][geo["iNeighbour"][i, 1:n + 1] - 1] >= i + 1)
        assert np.all(np.diff(geo["neighDist2"][i, 1:n + 1]) >= -1e-10)


def test_cell_translations_origin_first_and_complete():
    lat = np.array([[10.0, 0, 0], [2.0, 9.0, 0], [0, 1.0, 12.0]])
    c1, r1 = geometry.cell_translations(lat, 11.0)
    c2, r2 = oracle.cell_translations(lat, 11.0)
    assert np.array_equal(c1, c2) and np.allclose(r1, r2)
    assert np.all(c1[0] == 0) and len(np.unique(c1, axis=0)) == len(c1)


def test_skf_reader_and_full_table_match_oracle(slako_dir):
    names = ["Ce", "Fe", "O", "H"]
    sk, d = workloads.make_slako_set(names, slako_dir)
    model, selfegy = oldskdata.load_model(d, names, sk.ang_shells)
    for (i, j), (dist, ham, over) in sk.tables.items():
        assert ham.shape[1] == model.n_integrals(i, j) == slako.n_sk_integrals(sk.ang_shells[i - 1], sk.ang_shells[j - 1])
        assert np.array_equal(ham, model.tables[(0, i, j)][1]) and np.array_equal(over, model.tables[(1, i, j)][1])
        assert dist == model.tables[(0, i, j)][0]
    assert selfegy == sk.selfegy
    assert abs(sk.cutoff() - model.cutoff()) < 1e-15
    # l1 > l2 columns come from the B-A file (parser.F90:4110-4114): O(p)-H(s) uses H-O's sp-sigma
    ho = slako.read_skf(d + "/H-O.skf", homo=False)
    dist, ham, over = sk.tables[(3, 4)]        # O atom, H neighbour: shells (s,p) x (s) -> [ss, ps]
    assert np.array_equal(ham[:, 1], ho.skHam[:, 19 - 1])
    # classic 10-column and extended 20-column files are both in the set
    assert open(d + "/Ce-Ce.skf").readline().startswith("@")
    assert not open(d + "/O-H.skf").readline().startswith("@")
    assert ho.nGrid == 500 and ho.dist == 0.02


def test_monkhorst_pack_444():
    k, w = geometry.monkhorst_pack((4, 4, 4))
    assert len(k) == 32 and abs(w.sum() - 1.0) < 1e-14
    k0, w0 = geometry.monkhorst_pack((4, 4, 4), shift=(0, 0, 0))
    assert len(k0) == 36 and abs(w0.sum() - 1.0) < 1e-14


def test_synthetic_overlap_is_positive_definite():
    for name in ("h2o", "ce_o", "fec128"):
        w, sk, geo, model, selfegy = case(name)
        S = oracle.unpack_hs_real(geo, oracle.build_s(model, geo))
        ev = np.linalg.eigvalsh(np.tril(S) + np.tril(S, -1).T)
        assert ev.min() > 1e-3, (name, ev.min())


def test_generated_rotation_code_is_in_sync(tmp_path):
    """dftbplus_b200/csrc/sk_gaunt_code.cuh is generated (tools/gen_sk_gaunt.py: Gaunt coefficients
    by exact quadrature, self-checked against the rotation-matrix definition); the committed file
    must be what the generator writes today."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_sk_gaunt", os.path.join(root, "tools", "gen_sk_gaunt.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    tables = gen.build()
    assert gen.self_check(tables) < 1e-13
    committed = open(gen.OUT_CODE).read()
    gen.OUT_CODE = str(tmp_path / "sk_gaunt_code.cuh")
    gen.emit_code(tables)
    assert open(gen.OUT_CODE).read() == committed
