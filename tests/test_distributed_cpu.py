"""The N>1 path on CPU: two gloo ranks split the atoms exactly like the reference's MPI ranks
(getChunkRanges / getIndicesWithWorkload), (k,spin) pairs round-robin (TParallelKS), and
assemble_chunks (= assembleChunks, an all-reduce SUM) restores the single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import oracle, case
    from dftbplus_b200 import schedule
    w, sk, geo, model, selfegy = case("si2")
    # K1: contiguous atom chunk per rank, zero elsewhere, then all-reduce
    first, last = schedule.get_chunk_ranges(world, rank, 1, geo["nAtom"])
    part = oracle.build_s(model, geo, rank=rank, nranks=world)
    own = np.zeros(geo["sparseSize"], dtype=bool)
    for a in range(first - 1, last):
        n = geo["nNeighbourSK"][a]
        end = geo["iPair"][a + 1, 0] if a + 1 < geo["nAtom"] else geo["sparseSize"]
        own[geo["iPair"][a, 0]:end] = True
    assert np.all(part[~own] == 0.0)
    s = schedule.assemble_chunks(torch.from_numpy(part.copy())).numpy()
    # (k,spin) sharding of the density step + all-reduce of the sparse rho (main.F90:4327)
    h0 = oracle.build_h0(model, geo, selfegy)
    import scipy.linalg
    rho = np.zeros(geo["sparseSize"])
    for (iK, iS) in schedule.parallel_ks(world, rank, len(w["kpoints"]), 1):
        kp, kw = w["kpoints"][iK - 1], w["kweights"][iK - 1]
        Hk, Sk = oracle.unpack_hs_cplx(geo, h0, kp), oracle.unpack_hs_cplx(geo, s, kp)
        eig, C = scipy.linalg.eigh(Hk, Sk, lower=True)
        f = np.zeros(len(eig)); f[:4] = 2.0
        oracle.pack_hs_cplx(geo, rho, oracle.density_matrix(np.asfortranarray(C), f), kp, kw)
    rho = schedule.assemble_chunks(torch.from_numpy(rho)).numpy()
    # K4 / K5: workload-balanced atom sets + all-reduce
    idx = schedule.get_indices_with_workload(world, rank, 1, geo["nAtom"], geo["nNeighbourSK"])
    q = schedule.assemble_chunks(torch.from_numpy(oracle.mulliken_per_orbital(model, geo, s, rho, iter_indices=idx))).numpy()
    d = schedule.assemble_chunks(torch.from_numpy(oracle.derivative_shift(model, geo, rho, 0.1 * rho, iter_indices=idx))).numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "r.npz"), s=s, rho=rho, q=q, d=d)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank(tmp_path):
    world, port = 2, 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle, case
    import scipy.linalg
    w, sk, geo, model, selfegy = case("si2")
    r = np.load(os.path.join(str(tmp_path), "r.npz"))
    s = oracle.build_s(model, geo)
    assert np.array_equal(r["s"], s)
    h0 = oracle.build_h0(model, geo, selfegy)
    rho = np.zeros(geo["sparseSize"])
    for kp, kw in zip(w["kpoints"], w["kweights"]):
        eig, C = scipy.linalg.eigh(oracle.unpack_hs_cplx(geo, h0, kp), oracle.unpack_hs_cplx(geo, s, kp), lower=True)
        f = np.zeros(len(eig)); f[:4] = 2.0
        oracle.pack_hs_cplx(geo, rho, oracle.density_matrix(np.asfortranarray(C), f), kp, kw)
    assert np.allclose(r["rho"], rho, atol=1e-13)
    assert np.allclose(r["q"], oracle.mulliken_per_orbital(model, geo, s, rho), atol=1e-12)
    assert abs(r["q"].sum() - 8.0) < 1e-10                       # 2 Si atoms x 4 electrons
    assert np.allclose(r["d"], oracle.derivative_shift(model, geo, rho, 0.1 * rho), atol=1e-12)


def _panel_worker(rank, world, port, out_dir):
    """More ranks than (k,spin) pairs: the ranks of a group split ONE density matrix into column
    panels (schedule.dm_panels); a rank keeps only its panel's columns, packs only its panel's atoms
    (here: the oracle's pack on a matrix zeroed outside the panel), and the all-reduce assembles rho."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.linalg
    from helpers import oracle, case, fermi_fillings, n_electrons
    from dftbplus_b200 import schedule
    w, sk, geo, model, selfegy = case("fec128")
    h0, s = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    eig, C = scipy.linalg.eigh(oracle.unpack_hs_real(geo, h0), oracle.unpack_hs_real(geo, s), lower=True)
    f = fermi_fillings(eig, n_electrons(w, sk))
    a0, a1, c0, c1 = schedule.dm_panels(geo["iAtomStart"], world)[rank]
    dm = oracle.density_matrix(np.asfortranarray(C), f)
    panel = np.zeros_like(dm)
    panel[:, c0 - 1:c1] = dm[:, c0 - 1:c1]                      # what *_density_matrix_*_cols leaves valid
    # pack only the blocks of atoms a0..a1 (packHS_real, sparse2dense.F90:1024-1096, restricted to the panel)
    part = np.zeros(geo["sparseSize"])
    for i in range(a0 - 1, a1):
        for k in range(0, geo["nNeighbourSK"][i] + 1):
            j = geo["img2CentCell"][geo["iNeighbour"][i, k] - 1] - 1
            n1, n2 = geo["nOrbAtom"][i], geo["nOrbAtom"][j]
            ii, jj = geo["iAtomStart"][i] - 1, geo["iAtomStart"][j] - 1
            blk = panel[jj:jj + n2, ii:ii + n1].copy()
            if i == j:                                            # on-site: symmetrised from the lower triangle
                blk = np.tril(blk) + np.tril(blk, -1).T
            part[geo["iPair"][i, k]:geo["iPair"][i, k] + n1 * n2] += blk.T.reshape(-1)
    rho = schedule.assemble_chunks(torch.from_numpy(part)).numpy()
    if rank == 0:
        np.savez(os.path.join(out_dir, "p.npz"), rho=rho)
    dist.barrier()
    dist.destroy_process_group()


def test_column_panels_over_two_gloo_ranks(tmp_path):
    world, port = 2, 30100 + os.getpid() % 500
    mp.spawn(_panel_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scipy.linalg
    from helpers import oracle, case, fermi_fillings, n_electrons
    w, sk, geo, model, selfegy = case("fec128")
    h0, s = oracle.build_h0(model, geo, selfegy), oracle.build_s(model, geo)
    eig, C = scipy.linalg.eigh(oracle.unpack_hs_real(geo, h0), oracle.unpack_hs_real(geo, s), lower=True)
    f = fermi_fillings(eig, n_electrons(w, sk))
    ref = oracle.pack_hs_real(geo, np.zeros(geo["sparseSize"]), oracle.density_matrix(np.asfortranarray(C), f))
    got = np.load(os.path.join(str(tmp_path), "p.npz"))["rho"]
    assert np.allclose(got, ref, atol=1e-14)
