"""The CUDA path against the reference's PUBLISHED definition of the transform, with no oracle in between.

tests/definition.py evaluates the triple sum of docs/tutorial/initialization.md:154 and its inverse quadrature term by
term with third-party Legendre functions (mpmath legenp / scipy lpmv), numpy's exp and numpy's Gauss-Legendre weights.
tests/test_oracle_independent.py holds the ORACLE to those sums on the CPU; here the same sums check what
mlegs_b200_trans returns through the C ABI -- exponent signs, the placement of the 1/N factors, the phi-pair packing, the
FFT order of k, the parity fold of the Legendre stage and the device library's own binary128 tables all at once.

Shapes cover both Legendre code paths (TMA/DMMA kernels for nr a multiple of 4, the cp.async kernels otherwise) and both
FFT families (register-resident kernels from length 32, the shared-memory Stockham kernels for the other lengths).
(The file name sorts last on purpose: these checks came after the round's last GPU session.)"""
import numpy as np
import pytest

import mlegs_b200 as mb
import definition as dfn

pytestmark = pytest.mark.gpu

TOL = 1.0e-12   # BASELINE.json north_star; the sums themselves agree with the oracle to 1e-14 .. 1e-13 (CPU test)

# (nr, np, nz, nrchop, npchop, nzchop, ell, hyperpow): shapes of tests/test_gpu_trans.py, so every kernel variant used
# here is one the oracle parity tests exercise as well
CASES = {
    "gate3d": (32, 16, 8, 32, 9, 5, 4.0, 8),           # tools/validate_tutorials.py:222-238
    "radix35": (36, 30, 20, 30, 12, 9, 2.0, 4),        # factors 3 and 5; chops below the maximum
    "nr30": (30, 16, 8, 30, 9, 5, 2.0, 0),             # nr/2 odd: the cp.async Legendre kernels
    "cube64": (64, 64, 64, 64, 33, 33, 4.0, 0),        # register-resident FFT kernels, TMA/DMMA Legendre kernels
}


@pytest.mark.parametrize("case", list(CASES))
def test_device_transform_equals_the_published_triple_sum(case):
    nr, npts, nz, nrc, npc, nzc, ell, hp = CASES[case]
    kit = mb.TfmKit.init(mb.make_params(nr, npts, nz, nrc, npc, nzc, ell=ell, zlen=2 * np.pi, hyperpow=hp,
                                        hypervisc=(1e-6 if hp else 0.0)))
    glb = tuple(kit.glb_sz)
    # scipy's lpmv as the third-party Legendre function (the CPU test also runs mpmath's legenp at 30 digits)
    B = dfn.basis_at_all_nodes(kit.x, nrc, npc, "scipy")

    # synthesis FFF -> PPP
    a = dfn.random_triangular(glb, nrc, npc, nz, seed=11)
    s = mb.Scalar("FFF").upload(a)
    mb.trans(s, "PPP")
    assert s.space == "PPP"
    got = dfn.unpack_ppp(s.download(), nr, npts, nz)
    want = dfn.synthesis_by_definition(a, B, npts, nz)
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < TOL

    # analysis PPP -> FFF
    f = np.random.default_rng(12).standard_normal((nr, npts, nz))
    s = mb.Scalar("PPP").upload(dfn.pack_ppp(f, glb))
    mb.trans(s, "FFF")
    assert s.space == "FFF"
    e = s.download()
    want = dfn.analysis_by_definition(f, B, kit.x)
    scale = np.max(np.abs(want))
    for m in range(npc):
        nn = max(nrc - m, 0)
        assert np.max(np.abs(e[:nn, m, :nz] - want[:nn, m])) / np.max(np.abs(want[:nn, m])) < TOL, m
        if nn < e.shape[0]:
            assert np.max(np.abs(e[nn:, m, :nz])) <= TOL * scale      # nothing beyond the triangular truncation
    mb.finalize()
