"""Host halves of msave / mload and of assemble / disassemble.

The reference gathers the global array on rank 0 (`assemble`, submodules/mlegs_scalar_dist.f90:70-203) and writes
it there (`msave_scalar`, submodules/mlegs_scalar_io.f90:6-115).  Here every rank writes / reads the byte ranges
of its own slab (csrc/field_io.cu); these helpers expose that host code for arrays that already live on the host
(no CUDA needed), plus the global-array gather for callers that really want one.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import Field, check
from .dist import decompose


def slab_meta(glb_sz, rank: int, nranks: int, space: str = "PPP", ln: float = 0.0, offsets=(0, 0, 0)) -> Field:
    """mlegs_field metadata of rank `rank`'s slab: PPP -> r sharded (axis_comm 1,0,2), else m sharded (2,1,0)."""
    f = Field()
    physical = space[:3] == "PPP"
    axis = 0 if physical else 1
    cnt, off = decompose(glb_sz[axis], nranks, rank)
    for a in range(3):
        f.glb_sz[a] = glb_sz[a]
        f.loc_sz[a] = cnt if a == axis else glb_sz[a]
        f.loc_st[a] = off if a == axis else 0
    for a, v in enumerate((1, 0, 2) if physical else (2, 1, 0)):
        f.axis_comm[a] = v
    f.ln = ln
    f.nrchop_offset, f.npchop_offset, f.nzchop_offset = offsets
    f.space = space.encode()
    return f


def msave_part(meta: Field, block: np.ndarray, fn: str, is_binary=False, is_global=True, rank=0, create=True):
    a = np.asfortranarray(block, dtype=np.complex128)
    assert a.shape == tuple(meta.loc_sz)
    check(_lib.lib().mlegs_b200_msave_part(C.byref(meta), a.ctypes.data_as(C.c_void_p), str(fn).encode(),
                                           int(is_binary), int(is_global), rank, int(create)))


def mload_part(fn: str, meta: Field, is_binary=False, is_global=True, rank=0) -> np.ndarray:
    out = np.zeros(tuple(meta.loc_sz), dtype=np.complex128, order="F")
    check(_lib.lib().mlegs_b200_mload_part(str(fn).encode(), C.byref(meta), out.ctypes.data_as(C.c_void_p),
                                           int(is_binary), int(is_global), rank))
    return out


def msave_global(meta: Field, block: np.ndarray, fn: str, is_binary=False, group=None):
    """Cooperative global save of host blocks under torch.distributed: rank 0 lays the file out, then every rank
    writes its slab at its byte offsets (the order is the only synchronisation needed)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank == 0:
        msave_part(meta, block, fn, is_binary, True, 0, True)
    if dist.is_initialized():
        dist.barrier(group=group)
    if rank != 0:
        msave_part(meta, block, fn, is_binary, True, rank, False)
    if dist.is_initialized():
        dist.barrier(group=group)


def assemble(s, group=None) -> np.ndarray:
    """scalar_assemble (dist:70-203): the global array, on every rank (slabs summed into zeros over `group`)."""
    glb = np.zeros(s.glb_sz, dtype=np.complex128, order="F")
    glb[s.global_slices()] = s.download()
    try:
        import torch
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            t = torch.from_numpy(glb.view(np.float64).reshape(-1, order="A").copy())
            if dist.get_backend(group) == "nccl":
                t = t.cuda()
            dist.all_reduce(t, group=group)
            glb = np.asfortranarray(t.cpu().numpy().view(np.complex128).reshape(s.glb_sz, order="F"))
    except ImportError:   # pragma: no cover
        pass
    return glb


def disassemble(s, glb: np.ndarray):
    """scalar_disassemble (dist:205-368) without the scatter: every rank keeps its slab of the global array."""
    return s.upload_global(glb)
