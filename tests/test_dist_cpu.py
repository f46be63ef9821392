"""CPU (gloo, world_size 2) coverage of the multi-rank host logic: the slab decomposition and the exchange
plan that the device put kernel executes (mlegs_b200_dist_put_map runs the kernel's own addressing code on the
host), checked against the reference meaning of scalar_exchange as restated by the oracle
(submodules/mlegs_scalar_dist.f90:6-67, 395-504; mlegs_envir_mpi.f90:6-31)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, glb_sz, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from mlegs_b200 import dist as mdist
    from oracle import mlegs_oracle as mo
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        nrdim, npdim, nz = glb_sz
        # index-encoded fill like src/apps/assemble.f90:49 (e = i*10^4 + j*10^2 + k)
        i, j, k = np.meshgrid(np.arange(nrdim), np.arange(npdim), np.arange(nz), indexing="ij")
        full = np.asfortranarray((i * 1e4 + j * 1e2 + k) + 1j * (k * 1e4 + i * 1e2 + j))
        errs = []
        # r is distributed with the reference's decompose (mlegs_envir_mpi.f90:6-31); m cyclically (rank q owns
        # m = q, q + P, ...: include/mlegs_b200.h), so "the block of rank p" along m is full[:, p::P, :]
        def block(axis, p):
            sl = [slice(None)] * 3
            if axis == 1:
                c, o = mdist.decompose(nrdim, world, p)
                assert (c, o) == mo.decompose(nrdim, world, p)
                sl[0] = slice(o, o + c)
            else:
                own = mdist.m_owned(npdim, world, p)
                sl[1] = slice(own.start, own.stop, own.step)
            return np.asfortranarray(full[tuple(sl)])

        for direction, (axis_old, axis_new) in enumerate([(2, 1), (1, 2), (1, 2)]):
            # before: all of axis_old local, axis_new distributed; direction 2 = (1,2) into the transit layout
            mine = block(axis_new, rank)
            dst_rank, dst_index = mdist.put_map(direction, rank, world, nrdim, npdim, nz)
            flat = mine.ravel(order="F")
            msgs = [(dst_index[dst_rank == p], flat[dst_rank == p]) for p in range(world)]
            # the all-to-all: every rank receives the messages addressed to it
            gathered = [None] * world
            dist.all_gather_object(gathered, msgs)
            want = block(axis_old, rank)       # reference meaning: all of axis_new, this rank's share of axis_old
            if direction == 2:                 # columns grouped by the rank that owns them
                order = np.concatenate([np.arange(npdim)[p::world] for p in range(world)])
                want = np.asfortranarray(want[:, order, :])
            got = np.full(want.size, np.nan + 0j)
            for src in range(world):
                idx, val = gathered[src][rank]
                assert not np.isfinite(got[idx]).any(), "two sources wrote the same element"
                got[idx] = val
            got = got.reshape(want.shape, order="F")
            errs.append(float(np.abs(got - want).max()))
        # with contiguous ownership the same plan is the reference's scalar_exchange (dist:6-67): on ONE rank
        blocks = [np.asfortranarray(full)]
        assert np.array_equal(mo.exchange_global(blocks, glb_sz, 2, 1)[0], full)
        q.put((rank, errs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("glb_sz", [(19, 9, 4), (35, 9, 8), (11, 13, 1)])
@pytest.mark.parametrize("world", [2, 3])
def test_exchange_plan_matches_reference_semantics(glb_sz, world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, glb_sz, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in results:
        assert errs == [0.0, 0.0, 0.0], (rank, errs)   # bit-exact data movement


def test_decompose_covers_the_axis():
    from mlegs_b200 import dist as mdist
    for n in (1, 7, 64, 131, 264, 520):
        for p in (1, 2, 3, 4, 8):
            segs = [mdist.decompose(n, p, r) for r in range(p)]
            assert segs[0][1] == 0 and sum(c for c, _ in segs) == n
            for (c0, o0), (c1, o1) in zip(segs, segs[1:]):
                assert o1 == o0 + c0 and c0 >= c1    # first mod(n,p) ranks get one extra (mlegs_envir_mpi.f90:20-28)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_staged_exchange_equals_direct_put(world):
    """The staged exchange(1,2) (Legendre epilogue -> local staging buffer in destination order -> ship kernel's
    contiguous runs) must deliver every element exactly where the direct put into the transit layout does; the staging map is a
    permutation of the local block and every ship run is contiguous on both sides."""
    from mlegs_b200 import dist as mdist
    nrdim, npdim, nz = 35, 9, 4
    for rank in range(world):
        dst_rank, dst_index = mdist.put_map(2, rank, world, nrdim, npdim, nz)
        stage, ship_rank, ship_index = mdist.stage_map(rank, world, nrdim, npdim, nz)
        n = stage.size
        assert n == dst_rank.size
        assert np.array_equal(np.sort(stage), np.arange(n)), "staging map is not a permutation"
        assert (ship_rank >= 0).all() and (ship_index >= 0).all(), "ship runs do not cover the staging buffer"
        assert np.array_equal(ship_rank[stage], dst_rank)
        assert np.array_equal(ship_index[stage], dst_index)
