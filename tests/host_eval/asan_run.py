"""Driver of tests/test_neighbour_list_host.py::test_kernels_are_clean_under_address_sanitizer: the
host build of the neighbour-list kernels with -fsanitize=address,undefined (every buffer has exactly the
size the launch sequence allocates on the GPU, so an index that would be an illegal address there is
caught here).  Usage: LD_PRELOAD=libasan.so python asan_run.py <libhostnl_asan.so>"""
import sys, ctypes, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import test_neighbour_list_host as T
from helpers import oracle, case
lib = ctypes.CDLL(sys.argv[1])
n=0
for name in ("h2o","ce_o","scan","si2","fec128","si512"):
    w, sk, geo, model, selfegy = case(name)
    cutoff = sk.cutoff()
    cell, rcell = T.cells_of(w, cutoff)
    for sym in (False, True):
        ref = oracle.neighbour_list_full(sk.n_orb_species, w["coord0"], w["species0"], rcell, cutoff, 0.7*cutoff, symmetric=sym)
        got = T.host_list(lib, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, 0.7*cutoff, n_threads=300, seed=3, symmetric=sym)
        T.assert_same_list(got, ref); n+=1
    # smaller follow-up update reusing the (now larger) buffers
    got = T.host_list(lib, w["coord0"][:1], w["species0"][:1], sk.n_orb_species, cell, rcell, cutoff, cutoff)
    out = np.zeros(1, dtype=np.int32)
    lib.host_nl_nr_of_neighbours(ctypes.c_double(3.0), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 7, 1)
lib.host_nl_set_rank_max(5)
w, sk, geo, model, selfegy = case("si512"); cutoff=sk.cutoff(); cell, rcell = T.cells_of(w, cutoff)
T.host_list(lib, w["coord0"], w["species0"], sk.n_orb_species, cell, rcell, cutoff, cutoff)
c = np.random.default_rng(1).uniform(-30,30,(50,3)); lat=np.ascontiguousarray(np.diag([7.,8.,9.]))
lib.host_nl_fold(50, c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), lat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 9, 2)
# per-atom population kernels and the stand-alone sort / bisection entry points
w, sk, geo, model, selfegy = case("ce_o")
rho = np.random.default_rng(2).normal(size=geo["sparseSize"])
ipair = np.ascontiguousarray(geo["iPair"], dtype=np.int32); start = np.ascontiguousarray(geo["iAtomStart"], dtype=np.int32)
q = np.zeros(geo["nAtom"]); dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
lib.host_pop_onsite(geo["nAtom"], ipair.shape[1], ipair.ctypes.data_as(ip), start.ctypes.data_as(ip), rho.ctypes.data_as(dp), q.ctypes.data_as(dp), 5, 1)
qo = np.random.default_rng(3).normal(size=(geo["nAtom"], 16))
lib.host_pop_atom_sum(geo["nAtom"], 16, qo.ctypes.data_as(dp), q.ctypes.data_as(dp), 5, 1)
for m in (0, 1, 2, 33):
    keys = np.arange(1, m + 1, dtype=np.int32); vals = np.random.default_rng(m).uniform(size=max(m, 1))[:m].copy()
    lib.host_heap_sort_records(m, keys.ctypes.data_as(ip), vals.ctypes.data_as(dp), ctypes.c_double(1e-10))
    lib.host_search_asc_real_geq(m, vals.ctypes.data_as(dp), ctypes.c_double(0.5), ctypes.c_double(1e-10))
print("asan/ubsan run finished,", n, "lists compared")
