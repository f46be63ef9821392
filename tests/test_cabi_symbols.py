"""The C-ABI library loads and exports every symbol include/mlegs_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mlegs_b200 as mb
from mlegs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mlegs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mlegs_b200_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 45
    l = ctypes.CDLL(mb.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/mlegs_b200.h but not exported"
    assert names == _lib.exported_symbols()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = mb.make_params(32, 48, 1, 32, 25, 1, ell=1.0, zlen=1.0)
    kit = mb.TfmKit.build_tables(p)          # host-only table build works
    with pytest.raises(mb.MlegsError, match="no CUDA device"):
        kit.upload()
    with pytest.raises(mb.MlegsError, match="not initialized"):
        mb.Scalar("PPP")


def test_param_validation_messages():
    # the reference's stop strings (sinit:35-62)
    bad = [((31, 48, 1, 31, 25, 1), "nr must be even"),
           ((32, 14, 1, 32, 8, 1), "np must only have factors of 2, 3 and 5"),
           ((32, 48, 7, 32, 25, 1), "nz must be even"),
           ((32, 48, 1, 33, 25, 1), "nrchop must be smaller than or equal to nr"),
           ((32, 48, 1, 32, 26, 1), "npchop <= np/2 \\+ 1"),
           ((32, 48, 8, 32, 25, 6), "nzchop <= nz/2 \\+ 1")]
    for dims, msg in bad:
        with pytest.raises(mb.MlegsError, match=msg):
            mb.TfmKit.build_tables(mb.make_params(*dims))
