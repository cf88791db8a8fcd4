"""Host-side plumbing of the row-sharded multi-GPU path (one process per GPU, torch.distributed for the
rendezvous; the data-path collective -- the all-gather of the Krylov vector slices -- runs inside
liboptimet_b200.so on NCCL).  Replaces the reference's MPI/BLACS layer (srcAna/mpi, srcAna/scalapack):
contiguous particle block-rows with the remainder rule of srcAna/PreconditionedMatrix.cpp:418-424.

Backend-agnostic on purpose: the same functions run over `gloo` in the CPU tests (tests/test_sharding_gloo.py)
and over `nccl` in bench.py.
"""
import numpy as np

from . import capi


def row_range(nobj, world, rank, blk):
    """Element range [lo, hi) of the replicated length-(blk*nobj) vectors owned by `rank`."""
    first, count = capi.Library.partition(nobj, world, rank)
    return first * blk, (first + count) * blk


def exchange_unique_id(dist, rank):
    """Rank 0 creates the NCCL unique id, everybody receives it (128 bytes)."""
    uid = [capi.Library.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return uid[0]


def attach(solver, dist, rank, world):
    """Give a host.Solver its NCCL communicator (no-op for world == 1)."""
    if world > 1:
        solver.comm_init(exchange_unique_id(dist, rank), rank, world)


def allgather_slices(dist, torch, vec, nobj, world, blk):
    """Host/tensor twin of the library's slice all-gather: every rank holds `vec` (full length) with only its
    own row range valid; afterwards all ranges are valid everywhere.  Handles uneven partitions the same way
    the library does (one broadcast per owner)."""
    t = vec if isinstance(vec, torch.Tensor) else torch.from_numpy(vec)
    for r in range(world):
        lo, hi = row_range(nobj, world, r, blk)
        if hi > lo:
            dist.broadcast(t[lo:hi], src=r)
    return vec


def sum_partials(dist, torch, values, device=None):
    """Cross sections are linear in the per-particle terms: each rank returns the partial sums of its own
    particles (the reference gathers the same partial sums with MPI_Gather, Simulation.cpp:510-572)."""
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def max_over_ranks(dist, torch, value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
