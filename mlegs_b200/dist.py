"""Multi-GPU plumbing: one process per GPU, slab decomposition.

torch.distributed is used for exactly one thing -- all-gathering the 64-byte CUDA-IPC handles of the
per-rank exchange windows.  The data path (pencil transposes, scalar all-reduces) then runs inside
libmlegs_b200.so as one-sided puts over NVLink peer memory (csrc/dist.cu); it replaces set_comm_grps +
MPI_Alltoallw of submodules/mlegs_scalar_dist.f90:468-578.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


def decompose(nsize: int, nprocs: int, proc: int):
    """submodules/mlegs_envir_mpi.f90:6-31: (count, offset) of rank `proc`."""
    q, r = divmod(nsize, nprocs)
    return (q + 1, (q + 1) * proc) if r > proc else (q, q * proc + r)


def m_owned(npdim: int, nprocs: int, proc: int) -> range:
    """Azimuthal columns of rank `proc` when m is distributed: cyclic, m = proc, proc + nprocs, ... (the work per
    column falls linearly with m, so contiguous blocks would be unbalanced; include/mlegs_b200.h)."""
    return range(proc, npdim, nprocs)


def attach(group=None):
    """Export this rank's window, all-gather the IPC handles over `group`, map the peers' windows."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    handle = (C.c_ubyte * 64)()
    ptr = C.c_void_p()
    nbytes = C.c_size_t()
    check(_lib.lib().mlegs_b200_dist_window(C.byref(ptr), C.byref(nbytes), handle))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(handle), group=group)
    blob = b"".join(handles)
    buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
    check(_lib.lib().mlegs_b200_dist_attach(buf))
    dist.barrier(group=group)


def detach():
    check(_lib.lib().mlegs_b200_dist_detach())


def allreduce(values) -> np.ndarray:
    """Sum of host doubles over all ranks (rank-ordered, identical everywhere); identity on one rank."""
    a = np.ascontiguousarray(values, dtype=np.float64).copy()
    check(_lib.lib().mlegs_b200_dist_allreduce(a.ctypes.data_as(C.c_void_p), a.size))
    return a


def put_map(direction: int, rank: int, nranks: int, nrdim: int, npdim: int, nz: int):
    """Host-only exchange plan of one rank (no CUDA): (dst_rank, dst_index) per local element.  direction 0:
    exchange(2,1); 1: exchange(1,2); 2: exchange(1,2) into the transit layout of the fused exchanges."""
    if direction == 0:
        n = decompose(nrdim, nranks, rank)[0] * npdim * nz
    else:
        n = nrdim * len(m_owned(npdim, nranks, rank)) * nz
    dst_rank = np.zeros(n, dtype=np.int32)
    dst_index = np.zeros(n, dtype=np.int64)
    check(_lib.lib().mlegs_b200_dist_put_map(direction, rank, nranks, nrdim, npdim, nz,
                                             dst_rank.ctypes.data_as(C.c_void_p),
                                             dst_index.ctypes.data_as(C.c_void_p)))
    return dst_rank, dst_index


def stage_map(rank: int, nranks: int, nrdim: int, npdim: int, nz: int):
    """Host-only plan of the staged exchange(1,2) of one rank (no CUDA): stage_index per local element, and
    (ship_rank, ship_index) per staging index."""
    n = nrdim * len(m_owned(npdim, nranks, rank)) * nz
    stage_index = np.zeros(n, dtype=np.int64)
    ship_rank = np.full(n, -1, dtype=np.int32)
    ship_index = np.full(n, -1, dtype=np.int64)
    check(_lib.lib().mlegs_b200_dist_stage_map(rank, nranks, nrdim, npdim, nz,
                                               stage_index.ctypes.data_as(C.c_void_p),
                                               ship_rank.ctypes.data_as(C.c_void_p),
                                               ship_index.ctypes.data_as(C.c_void_p)))
    return stage_index, ship_rank, ship_index
