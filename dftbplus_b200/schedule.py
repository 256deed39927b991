"""Host-side partition logic of the multi-GPU path (one process per GPU).

Mirrors the reference's own decomposition so that N GPUs split the work exactly like N MPI
ranks/groups do: contiguous atom chunks for the sparse builds (getChunkRanges,
common/schedule.F90:230-263), greedy workload-balanced atom sets for Mulliken and forces
(getIndicesWithWorkload, :267-307), and round-robin (k-point, spin) pairs over groups
(TParallelKS_init, type/parallelks.F90:46-89).  assembleChunks (schedule.F90:170-182, an in-place
MPI_Allreduce SUM) becomes a torch.distributed all_reduce over NCCL (gloo in the CPU tests).
"""
import numpy as np


def get_chunk_ranges(group_size, my_rank, global_first, global_last):
    range_length = global_last - global_first + 1
    n_local = range_length // group_size
    remainder = range_length % group_size
    if my_rank < remainder:
        n_local += 1
        local_first = global_first + my_rank * n_local
    else:
        local_first = global_first + remainder * (n_local + 1) + (my_rank - remainder) * n_local
    return local_first, min(local_first + n_local - 1, global_last)


def get_indices_with_workload(group_size, my_rank, global_first, global_last, workload):
    """workload is indexed 1-based like the Fortran array: workload[ii-1]."""
    rank_workload = [0] * group_size
    out = []
    for ii in range(global_first, global_last + 1):
        rank = min(range(group_size), key=lambda r: (rank_workload[r], r))   # minloc: first minimum
        rank_workload[rank] += max(1, int(workload[ii - 1]))
        if rank == my_rank:
            out.append(ii)
    return np.array(out, dtype=np.int32)


def parallel_ks(n_group, my_group, n_kpoint, n_spin):
    """[(iK, iS), ...] (1-based) handled by group my_group."""
    return [(iK, iS) for iS in range(1, n_spin + 1) for iK in range(1, n_kpoint + 1)
            if ((iS - 1) * n_kpoint + iK - 1) % n_group == my_group]


def assemble_chunks(tensor, group=None):
    """assembleChunks: in-place SUM over the ranks (NCCL on GPUs, gloo on CPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor


def dm_panels(i_atom_start, n_panel):
    """Column panels of one density matrix for n_panel GPUs ("beyond (k,spin)
    sharding"): contiguous atom ranges whose dense columns carry about equal shares of the LOWER
    triangle (column c has nOrb - c rows below it, so the boundaries follow
    nOrb * (1 - sqrt(1 - p / n_panel))), snapped to atom starts.
    i_atom_start: iAtomStart(nAtom + 1), 1-based.  Returns [(atomFirst, atomLast, colFirst, colLast)]
    (1-based, inclusive), one per panel; a panel may be empty (atomFirst > atomLast)."""
    start = np.asarray(i_atom_start, dtype=np.int64)
    n_atom, n_orb = len(start) - 1, int(start[-1] - 1)
    cuts = [0]
    for p in range(1, n_panel):
        col = n_orb * (1.0 - np.sqrt(1.0 - p / n_panel))        # 0-based column boundary
        a = int(np.searchsorted(start[:-1] - 1, col, side="left"))
        cuts.append(min(max(a, cuts[-1]), n_atom))
    cuts.append(n_atom)
    out = []
    for p in range(n_panel):
        a0, a1 = cuts[p] + 1, cuts[p + 1]
        out.append((a0, a1, int(start[a0 - 1]) if a0 <= n_atom else n_orb + 1, int(start[a1]) - 1 if a1 >= 1 else 0))
    return out
