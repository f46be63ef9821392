"""msave / mload file formats (submodules/mlegs_scalar_io.f90:6-250) through the host halves of the C ABI: no GPU.

The binary stream layout is checked byte for byte against an independent numpy writer of the reference's write
statements; the formatted layout against the reference's edit descriptors ((3(1X,I10)), 1PE24.15E3) -- the
list-directed trailer (write(fo,*) s%ln, ...) has compiler-chosen field widths, so there the claim is only that the
reference's list-directed read accepts it; and the cooperative multi-rank write must produce exactly the file a single
rank writes."""
import os
import re
import socket
import struct
import sys

import numpy as np
import pytest

import mlegs_b200 as mb
from mlegs_b200 import io as mio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLB = (11, 5, 4)


def _global_array(glb_sz=GLB, seed=3):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal(glb_sz) * 10.0 ** rng.integers(-120, 120, glb_sz) + 1j * rng.standard_normal(glb_sz)
    a[0, 0, 0] = 0.0
    a[1, 0, 0] = complex(-1.0, 1e-300)
    return np.asfortranarray(a)


def _slab(a, meta):
    st, sz = tuple(meta.loc_st), tuple(meta.loc_sz)
    return np.asfortranarray(a[st[0]:st[0] + sz[0], st[1]:st[1] + sz[1], st[2]:st[2] + sz[2]])


def _reference_binary(a, ln, offs, space):
    """write(fo) size(a,1..3); write(fo) a; write(fo) s%ln, offsets; write(fo) s%space   (io.f90:47-52)"""
    return (struct.pack("<3i", *a.shape) + a.tobytes(order="F") + struct.pack("<d3i", ln, *offs) + space.encode())


@pytest.mark.parametrize("space", ["PPP", "FFF"])
@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_binary_global_file_is_the_reference_stream(tmp_path, space, nranks):
    a = _global_array()
    fn = str(tmp_path / f"fld_{space}_{nranks}.bin")
    for rank in range(nranks):               # rank 0 lays the file out, the others fill in their slabs
        meta = mio.slab_meta(GLB, rank, nranks, space, ln=0.25, offsets=(3, 0, -1))
        mio.msave_part(meta, _slab(a, meta), fn, is_binary=True, is_global=True, rank=rank, create=(rank == 0))
    assert open(fn, "rb").read() == _reference_binary(a, 0.25, (3, 0, -1), space)
    # every rank reads its own slab back, with the metadata
    for rank in range(nranks):
        meta = mio.slab_meta(GLB, rank, nranks, space)
        got = mio.mload_part(fn, meta, is_binary=True, is_global=True, rank=rank)
        assert np.array_equal(got, _slab(a, meta))
        assert (meta.ln, meta.nrchop_offset, meta.npchop_offset, meta.nzchop_offset) == (0.25, 3, 0, -1)
        assert meta.space.decode() == space


def test_binary_local_files(tmp_path):
    a = _global_array()
    fn = str(tmp_path / "loc")
    for rank in range(2):
        meta = mio.slab_meta(GLB, rank, 2, "FFF", ln=-1.5)
        blk = _slab(a, meta)
        mio.msave_part(meta, blk, fn, is_binary=True, is_global=False, rank=rank)
        raw = open(f"{fn}_{rank}", "rb").read()
        # write(fo) s%glb_sz, s%loc_sz, s%loc_st; write(fo) s%e; trailer   (io.f90:86-90)
        want = (struct.pack("<9i", *meta.glb_sz, *meta.loc_sz, *meta.loc_st) + blk.tobytes(order="F")
                + struct.pack("<d3i", -1.5, 0, 0, 0) + b"FFF")
        assert raw == want
        m2 = mio.slab_meta(GLB, rank, 2, "PPP")
        m2.loc_sz[:] = meta.loc_sz[:]
        m2.loc_st[:] = meta.loc_st[:]
        assert np.array_equal(mio.mload_part(fn, m2, is_binary=True, is_global=False, rank=rank), blk)
        assert m2.space.decode() == "FFF" and m2.ln == -1.5


def _parse_formatted(text, n1, n2, n3):
    """What the reference's formatted reader does (io.f90:160-181): header by (3(1X,I10)), then per (k, i) one record
    of 2*n2 numbers in 1PE24.15E3, one skipped record between planes."""
    lines = text.split("\n")
    assert re.fullmatch(r"( [ \d]{10}){3} ?", lines[0])
    assert [int(lines[0][1 + 11 * q: 11 + 11 * q]) for q in range(3)] == [n1, n2, n3]
    a = np.zeros((n1, n2, n3), dtype=np.complex128)
    ln = 1
    for k in range(n3):
        for i in range(n1):
            rec = lines[ln]
            ln += 1
            assert len(rec.rstrip(" ")) == 48 * n2
            for j in range(n2):
                re_s, im_s = rec[48 * j: 48 * j + 24], rec[48 * j + 24: 48 * j + 48]
                for tok in (re_s, im_s):
                    assert re.fullmatch(r" *-?\d\.\d{15}E[+-]\d{3}", tok), tok
                a[i, j, k] = complex(float(re_s), float(im_s))
        if k < n3 - 1:
            assert lines[ln].strip() == ""
            ln += 1
    trailer = lines[ln].split()
    space = lines[ln + 1].strip()
    return a, float(trailer[0]), [int(t) for t in trailer[1:4]], space


@pytest.mark.parametrize("space", ["PPP", "FFF"])
@pytest.mark.parametrize("nranks", [1, 2])
def test_formatted_global_file(tmp_path, space, nranks):
    a = _global_array()
    fn = str(tmp_path / f"fld_{space}_{nranks}.dat")
    for rank in range(nranks):
        meta = mio.slab_meta(GLB, rank, nranks, space, ln=0.125, offsets=(0, 1, 2))
        mio.msave_part(meta, _slab(a, meta), fn, is_binary=False, is_global=True, rank=rank, create=(rank == 0))
    text = open(fn).read()
    got, ln, offs, sp = _parse_formatted(text, *GLB)
    # 1PE24.15E3 keeps 16 significant digits
    assert np.allclose(got.real, a.real, rtol=1e-15, atol=0) and np.allclose(got.imag, a.imag, rtol=1e-15, atol=0)
    assert (ln, offs, sp) == (0.125, [0, 1, 2], space)
    if nranks > 1:      # cooperative write == single-rank write
        fn1 = str(tmp_path / "single.dat")
        meta = mio.slab_meta(GLB, 0, 1, space, ln=0.125, offsets=(0, 1, 2))
        mio.msave_part(meta, a, fn1, is_binary=False, is_global=True)
        assert open(fn1).read() == text
    for rank in range(nranks):
        meta = mio.slab_meta(GLB, rank, nranks, space)
        back = mio.mload_part(fn, meta, is_binary=False, is_global=True, rank=rank)
        assert np.array_equal(back, _slab(got, meta))
        assert meta.space.decode() == space and meta.ln == 0.125 and meta.nzchop_offset == 2


def test_formatted_local_file_and_special_values(tmp_path):
    a = _global_array((4, 3, 2))
    a[2, 1, 1] = complex(np.inf, np.nan)
    fn = str(tmp_path / "loc.dat")
    meta = mio.slab_meta((4, 3, 2), 0, 1, "PFP")
    mio.msave_part(meta, a, fn, is_binary=False, is_global=False, rank=0)
    lines = open(fn + "_0").read().split("\n")
    assert [int(t) for t in lines[0].split()] == [4, 3, 2] and [int(t) for t in lines[2].split()] == [0, 0, 0]
    back = mio.mload_part(fn, meta, is_binary=False, is_global=False, rank=0)
    ok = np.isfinite(a)
    assert np.allclose(back[ok], a[ok], rtol=1e-15, atol=0)
    assert np.isinf(back[2, 1, 1].real) and np.isnan(back[2, 1, 1].imag)


def test_load_errors(tmp_path):
    a = _global_array()
    fn = str(tmp_path / "x.bin")
    meta = mio.slab_meta(GLB, 0, 1, "FFF")
    mio.msave_part(meta, a, fn, is_binary=True)
    with pytest.raises(mb.MlegsError, match="mload_scalar: cannot open"):
        mio.mload_part(str(tmp_path / "missing.bin"), meta, is_binary=True)
    wrong = mio.slab_meta((GLB[0] + 1, GLB[1], GLB[2]), 0, 1, "FFF")
    with pytest.raises(mb.MlegsError, match="mloadc: size inconsistency between data and array"):
        mio.mload_part(fn, wrong, is_binary=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from mlegs_b200 import io as mio2
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        a = _global_array()
        for space, binary in (("PPP", True), ("FFF", True), ("FFF", False)):
            meta = mio2.slab_meta(GLB, rank, world, space, ln=0.5)
            mio2.msave_global(meta, _slab(a, meta), f"{fn}_{space}_{int(binary)}", is_binary=binary)
        q.put(rank)
    finally:
        dist.destroy_process_group()


def test_cooperative_write_under_gloo(tmp_path):
    """world_size 2: both processes write their slabs into ONE file; result == the reference stream."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    fn = str(tmp_path / "coop")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fn, q)) for r in range(2)]
    for p in procs:
        p.start()
    assert sorted(q.get(timeout=120) for _ in range(2)) == [0, 1]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a = _global_array()
    for space in ("PPP", "FFF"):
        assert open(f"{fn}_{space}_1", "rb").read() == _reference_binary(a, 0.5, (0, 0, 0), space)
    got, ln, offs, sp = _parse_formatted(open(f"{fn}_FFF_0").read(), *GLB)
    assert np.allclose(got.real, a.real, rtol=1e-15, atol=0) and sp == "FFF" and ln == 0.5
