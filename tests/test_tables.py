"""The product's host table builder (C++, binary128) against the oracle's 50-digit mpmath restatement
of sinit:181-300, and against the closed forms of docs/tutorial/transformation.md."""
import math

import numpy as np
import pytest

import mlegs_b200 as mb
from oracle import mlegs_oracle as mo
from helpers import oracle_params


@pytest.mark.parametrize("dims", [(32, 48, 1, 32, 25, 1, 1.0), (48, 8, 4, 40, 5, 3, 3.0), (20, 12, 6, 20, 7, 4, 4.0)])
def test_tables_bitwise_vs_mpmath(dims):
    nr, np_, nz, nrc, npc, nzc, ell = dims
    p = mb.make_params(nr, np_, nz, nrc, npc, nzc, ell=ell, zlen=2 * math.pi)
    k = mb.TfmKit.build_tables(p)
    ok = mo.kit_init(oracle_params(p))
    for name in ("x", "w", "ln", "r", "lognorm", "at0", "at1", "ak"):
        assert np.array_equal(getattr(k, name), getattr(ok, name)), name
    # binary128 (34 digits) vs 50 digits can differ by one ulp only where the polynomial nearly vanishes
    diff = k.pf != ok.pf
    assert diff.mean() < 1e-3
    assert np.max(np.abs(k.pf - ok.pf)) < 1e-30


def test_tables_closed_form():
    p = mb.make_params(32, 48, 1, 32, 25, 1, ell=1.0, zlen=1.0)
    k = mb.TfmKit.build_tables(p)
    r = k.r[:16]
    f1 = math.sqrt(5.0 / 12.0) * (-6.0 * r * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 2)
    f2 = math.sqrt(7.0 / 240.0) * (60.0 * r ** 2 * (r ** 2 - 1.0) / (r ** 2 + 1.0) ** 3)
    assert np.max(np.abs(k.pf[:, 1, 1] - f1)) < 5e-15
    assert np.max(np.abs(k.pf[:, 1, 2] - f2)) < 5e-15


def test_large_table_is_finite_and_orthonormal():
    # 256^3-class table (128 x 270 x 129): binary128 keeps the un-normalised recurrence in range
    p = mb.make_params(256, 256, 2, 256, 129, 2, ell=4.0, zlen=2 * math.pi)
    k = mb.TfmKit.build_tables(p)
    assert np.all(np.isfinite(k.pf))
    nrh = 128
    for m in (0, 64, 128):
        nn = 256 - m
        full = np.zeros((256, nn))
        par = (-1.0) ** np.arange(nn)
        full[:nrh] = k.pf[:, :nn, m]
        full[::-1][:nrh] = k.pf[:, :nn, m] * par[None, :]
        gram = full.T @ (k.w[:, None] * full)
        lim = min(nn, 256 - m - 2)
        assert np.max(np.abs(gram[:lim, :lim] - np.eye(lim))) < 1e-11   # limited by the double-precision GL nodes (sinit:187)


def test_table_cache_round_trip(tmp_path):
    """the on-disk table cache returns bit-identical tables, recomputes the physics-dependent
    ones (ln, r, ak) for the caller's ell / zlen / nz, and rebuilds a damaged file instead of trusting it."""
    import mlegs_b200 as mb
    p = mb.make_params(24, 12, 8, 22, 7, 5, ell=3.0, zlen=2 * np.pi)
    fresh = mb.TfmKit.build_tables(p)
    first = mb.TfmKit.build_tables(p, cache_dir=str(tmp_path))
    assert not first.from_cache
    files = list(tmp_path.iterdir())
    assert [f.name for f in files] == ["mlegs_tables_24_22_7.bin"]
    p2 = mb.make_params(24, 12, 16, 22, 7, 9, ell=1.5, zlen=4 * np.pi)      # other nz / ell / zlen: same file
    second = mb.TfmKit.build_tables(p2, cache_dir=str(tmp_path))
    assert second.from_cache
    for name in ("x", "w", "lognorm", "pf", "at0", "at1"):
        assert np.array_equal(getattr(first, name), getattr(fresh, name)), name
        assert np.array_equal(getattr(second, name), getattr(fresh, name)), name
    direct = mb.TfmKit.build_tables(p2)
    for name in ("ln", "r", "ak"):
        assert np.array_equal(getattr(second, name), getattr(direct, name)), name
    raw = bytearray(files[0].read_bytes())
    raw[len(raw) // 2] ^= 0x40                                              # flip one bit of the payload
    files[0].write_bytes(bytes(raw))
    third = mb.TfmKit.build_tables(p, cache_dir=str(tmp_path))
    assert not third.from_cache and np.array_equal(third.pf, fresh.pf)
    assert mb.TfmKit.build_tables(p, cache_dir=str(tmp_path)).from_cache    # rewritten
