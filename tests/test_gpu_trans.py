"""GPU parity of the spectral transforms (trans, ops:157-235) against the oracle, through the C ABI."""
import numpy as np
import pytest

import mlegs_b200 as mb
from oracle import mlegs_oracle as mo
from helpers import oracle_kit, random_fff, random_ppp, rel_l2

pytestmark = pytest.mark.gpu

# (nr, np, nz, nrchop, npchop, nzchop, ell, hyperpow)
CASES = {
    "gate2d": (32, 48, 1, 32, 25, 1, 1.0, 0),          # tools/validate_tutorials.py:255-269
    "gate2d_nr64": (64, 48, 1, 64, 25, 1, 1.0, 0),     # BASELINE.json configs[0] as quoted (NR=64)
    "gate3d": (32, 16, 8, 32, 9, 5, 4.0, 8),           # tools/validate_tutorials.py:222-238
    "radix35": (36, 30, 20, 30, 12, 9, 2.0, 4),        # np, nz with factors 3 and 5; chops below the maximum
    "cube64": (64, 64, 64, 64, 33, 33, 4.0, 0),
    # nr/2 odd: table rows are not 16-byte aligned, the cp.async Legendre kernels (legendre.cu) take over
    "nr30": (30, 16, 8, 30, 9, 5, 2.0, 0),
    # several row tiles of the TMA-fed Legendre kernels, radial tails in the last stage (nr/2 = 100), 2 n-tiles
    "nr200": (200, 16, 8, 200, 9, 5, 4.0, 0),
    # every length of the register-resident FFT kernels (fft_reg.cu): np/2 and nz in {32 .. 1024}
    "fft256": (16, 512, 256, 16, 9, 129, 4.0, 0),
    "fftp512": (8, 1024, 32, 8, 5, 17, 4.0, 0),
    "fftz512": (8, 64, 512, 8, 5, 257, 4.0, 0),
    "fftp1024": (8, 2048, 32, 8, 5, 17, 4.0, 0),
    "fftz1024": (8, 64, 1024, 8, 5, 513, 4.0, 0),
}
TOL = 1.0e-12   # BASELINE.json north_star: 1e-12 relative L2 per transform


def _setup(case):
    nr, np_, nz, nrc, npc, nzc, ell, hp = CASES[case]
    p = mb.make_params(nr, np_, nz, nrc, npc, nzc, ell=ell, zlen=2 * np.pi, hyperpow=hp,
                       hypervisc=(1e-6 if hp else 0.0))
    kit = mb.TfmKit.init(p)
    return kit, oracle_kit(kit)


@pytest.mark.parametrize("case", list(CASES))
def test_stagewise_forward_and_backward(case):
    kit, ok = _setup(case)
    e0 = random_ppp(ok, seed=1)
    s = mb.Scalar("PPP").upload(e0)
    so = mo.Scalar(e=e0.copy(order="F"), space="PPP")
    for sp in ("PFP", "FFP", "FFF", "FFP", "PFP", "PPP"):
        mb.trans(s, sp)
        mo.trans(so, sp, ok)
        assert s.space == sp
        got = s.download()
        assert rel_l2(got, so.e) < TOL, (case, sp)
        # continue both chains from the oracle's state so that stage errors do not accumulate
        s.upload(so.e)


@pytest.mark.parametrize("case", list(CASES))
def test_full_transforms_and_ln_term(case):
    kit, ok = _setup(case)
    e0 = random_fff(ok, seed=2)
    for ln in (0.0, 0.37):
        s = mb.Scalar("FFF").upload(e0)
        s.ln = ln
        so = mo.Scalar(e=e0.copy(order="F"), space="FFF", ln=ln)
        mb.trans(s, "PPP")
        mo.trans(so, "PPP", ok)
        assert rel_l2(s.download(), so.e) < TOL
        mb.trans(s, "FFF")
        mo.trans(so, "FFF", ok)
        assert rel_l2(s.download(), so.e) < TOL
        # forward(backward(x)) == x once x is the spectrum of a real field (Im of the m=0 and Nyquist
        # columns is dropped by the Hermitian inverse, external/ffte-7.0/zdfft2d.f:119-128)
        e1 = s.download()
        mb.trans(s, "PPP")
        mb.trans(s, "FFF")
        assert rel_l2(s.download(), e1) < 50 * TOL


@pytest.mark.parametrize("case", ["cube64", "gate3d"])
def test_axial_stage_alone_transforms_every_line(case):
    """Forward trans does not chop: FFP -> FFF on its own must transform rows beyond the
    truncation too, while inside a full trans only the retained lines are touched (compact path)."""
    kit, ok = _setup(case)
    rng = np.random.default_rng(7)
    e0 = np.asfortranarray(rng.standard_normal(kit.glb_sz) + 1j * rng.standard_normal(kit.glb_sz))
    for a, b in (("FFP", "FFF"), ("FFF", "FFP")):
        s = mb.Scalar(a).upload(e0)
        so = mo.Scalar(e=e0.copy(order="F"), space=a)
        mb.trans(s, b)
        mo.trans(so, b, ok)
        assert rel_l2(s.download(), so.e) < TOL, (case, a, b)


def test_chop_offsets_in_rtrans():
    kit, ok = _setup("gate3d")
    e0 = random_ppp(ok, seed=3)
    s = mb.Scalar("PPP").upload(e0)
    so = mo.Scalar(e=e0.copy(order="F"), space="PPP")
    s.chop_offset(3)
    so.chop_offset(3)
    mb.trans(s, "FFF")
    mo.trans(so, "FFF", ok)
    assert rel_l2(s.download(), so.e) < TOL
    s.chop_offset(kit.nrdim)        # nrc > nrdim
    with pytest.raises(mb.MlegsError, match="rtrans_backward: chopping in r too large"):
        mb.trans(s, "PPP")


def test_trans_rejects_bad_space():
    kit, ok = _setup("gate2d")
    s = mb.Scalar("PPP")
    with pytest.raises(mb.MlegsError, match="only taking PPP, PFP, FFP and FFF"):
        mb.trans(s, "XYZ")
    s.space = "PPF"
    with pytest.raises(mb.MlegsError, match="scalar space info corrupted"):
        mb.trans(s, "FFF")


def test_roundtrip_128_config_and_host_entry():
    """BASELINE.json configs[1]: 3D scalar PPP<->FFF round trip, NR=NP=NZ=128 (SURVEY.md section 8d input 2)."""
    p = mb.make_params(128, 128, 128, 128, 65, 65, ell=4.0, zlen=2 * np.pi, hyperpow=0)
    kit = mb.TfmKit.init(p)
    ok = oracle_kit(kit)
    e0 = random_fff(ok, seed=0)
    s = mb.Scalar("FFF").upload(e0)
    so = mo.Scalar(e=e0.copy(order="F"), space="FFF")
    mb.trans(s, "PPP")
    mo.trans(so, "PPP", ok)
    ppp = so.e.copy(order="F")
    assert rel_l2(s.download(), ppp) < TOL
    mb.trans(s, "FFF")
    mo.trans(so, "FFF", ok)
    assert rel_l2(s.download(), so.e) < TOL
    # the reference-facing host-buffer call gives the same numbers
    h = e0.copy(order="F")
    mb.trans_host(h, "FFF", "PPP")
    assert rel_l2(h, ppp) < TOL
    mb.trans_host(h, "PPP", "FFF")
    assert rel_l2(h, so.e) < TOL
    assert mb.launch_count() > 0
    # the pipelined batch entry gives the same numbers, field by field
    hs = [random_fff(ok, seed=sd).copy(order="F") for sd in (0, 3, 4, 5, 6)]
    want = []
    for h0 in hs:
        t = mo.Scalar(e=h0.copy(order="F"), space="FFF")
        mo.trans(t, "PPP", ok)
        want.append(t.e)
    mb.trans_host_batch(hs, "FFF", "PPP")
    for got, w in zip(hs, want):
        assert rel_l2(got, w) < TOL
    assert rel_l2(hs[0], ppp) < TOL


@pytest.mark.parametrize("case", ["gate2d", "gate3d", "radix35", "cube64"])
@pytest.mark.parametrize("nf", [2, 3, 11, 33])
def test_trans_many_is_bit_identical_to_trans(case, nf):
    """mlegs_b200_trans_many: one launch per stage over all scalars; same arithmetic per scalar as trans()."""
    kit, ok = _setup(case)
    es = [random_fff(ok, seed=10 + i) for i in range(nf)]
    lns = [0.0 if i % 2 == 0 else 0.1 * i for i in range(nf)]
    single, many = [], []
    for e, ln in zip(es, lns):
        for lst in (single, many):
            s = mb.Scalar("FFF").upload(e)
            s.ln = ln
            lst.append(s)
    for targets in (("PPP", "FFF"), ("FFP", "PFP", "PPP", "PFP", "FFP", "FFF")):
        for sp in targets:
            for s in single:
                mb.trans(s, sp)
            mb.trans_many(many, sp)
            for a, b in zip(single, many):
                assert b.space == sp
                assert np.array_equal(a.download(), b.download()), (case, nf, sp)
    # parity against the oracle for one member of the batch
    for s, e in zip(many, es):
        s.upload(e)
    so = mo.Scalar(e=es[1].copy(order="F"), space="FFF", ln=lns[1])
    mo.trans(so, "PPP", ok)
    mb.trans_many(many, "PPP")
    assert rel_l2(many[1].download(), so.e) < TOL


def test_trans_many_mixed_states_and_duplicates():
    kit, ok = _setup("gate3d")
    a = mb.Scalar("FFF").upload(random_fff(ok, seed=1))
    b = mb.Scalar("FFF").upload(random_fff(ok, seed=2))
    want = []
    for s in (a, b):
        t = s.copy()
        mb.trans(t, "PPP")
        want.append(t.download())
    mb.trans(b, "FFP")                      # mixed states: falls back to one trans() per scalar
    mb.trans_many([a, b], "PPP")
    assert np.array_equal(a.download(), want[0]) and np.array_equal(b.download(), want[1])
    with pytest.raises(mb.MlegsError, match="same scalar appears twice"):
        mb.trans_many([a, a], "FFF")
