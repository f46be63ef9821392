"""Host mirror of the transform kit (type tfm_kit_3d, modules/mlegs_spectfm.f90:14-63).

``TfmKit.init()`` does what ``tfm%init()`` does (submodules/mlegs_spectfm_init.f90:6-154): builds
the Gauss-Legendre nodes, normalisation logs and the P_L^m table on the host (C++, binary128
recurrence, through the C ABI) and uploads them once; they stay resident in HBM.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Params, check


def make_params(nr, np_, nz, nrchop, npchop, nzchop, ell=4.0, zlen=2.0 * np.pi, visc=1.0e-3, hyperpow=0,
                hypervisc=0.0, is_svv=True, svv_cutoff=0.75, svv_target=2.0e-2, svv_strength=0.12,
                svv_relax=0.25) -> Params:
    return Params(nr, np_, nz, nrchop, npchop, nzchop, ell, zlen, visc, hyperpow, hypervisc, int(bool(is_svv)),
                  svv_cutoff, svv_target, svv_strength, svv_relax)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


@dataclass
class TfmKit:
    params: Params
    x: np.ndarray
    w: np.ndarray
    ln: np.ndarray
    r: np.ndarray
    lognorm: np.ndarray
    pf: np.ndarray
    at0: np.ndarray
    at1: np.ndarray
    ak: np.ndarray
    from_cache: bool = False

    @property
    def nrdim(self):
        return self.params.nr + max(3, self.params.hyperpow)

    @property
    def npdim(self):
        return self.params.np // 2 + 1

    @property
    def nzdim(self):
        return self.params.nz

    @property
    def glb_sz(self):
        return (self.nrdim, self.npdim, self.nzdim)

    def tables(self) -> dict:
        return dict(x=self.x, w=self.w, lognorm=self.lognorm, pf=self.pf, at0=self.at0, at1=self.at1)

    @staticmethod
    def build_tables(p: Params, cache_dir: str | None = None) -> "TfmKit":
        """Host-only part of tfm%init(): no GPU needed.  cache_dir (default: $MLEGS_TABLE_CACHE if set) keeps the
        (nr, nrchop, npchop)-dependent tables on disk between runs (mlegs_b200_tfm_tables_cached)."""
        import os
        cache_dir = cache_dir if cache_dir is not None else os.environ.get("MLEGS_TABLE_CACHE", "")
        ne = p.nrchop + 14
        x = np.zeros(p.nr)
        w = np.zeros(p.nr)
        ln = np.zeros(p.nr)
        r = np.zeros(p.nr)
        lognorm = np.zeros((ne, p.npchop), order="F")
        pf = np.zeros((p.nr // 2, ne, p.npchop), order="F")
        at0 = np.zeros(p.nrchop)
        at1 = np.zeros(p.nrchop)
        ak = np.zeros(p.nz)
        hit = C.c_int(0)
        if cache_dir:
            os.makedirs(cache_dir, exist_ok=True)
        check(_lib.lib().mlegs_b200_tfm_tables_cached(C.byref(p), cache_dir.encode(), _ptr(x), _ptr(w), _ptr(ln),
                                                      _ptr(r), _ptr(lognorm), _ptr(pf), _ptr(at0), _ptr(at1),
                                                      _ptr(ak), C.byref(hit)))
        kit = TfmKit(p, x, w, ln, r, lognorm, pf, at0, at1, ak)
        kit.from_cache = bool(hit.value)
        return kit

    def upload(self, rank: int = 0, nranks: int = 1) -> "TfmKit":
        check(_lib.lib().mlegs_b200_init(C.byref(self.params), _ptr(self.x), _ptr(self.w), _ptr(self.lognorm),
                                         _ptr(self.pf), _ptr(self.at0), _ptr(self.at1), rank, nranks))
        return self

    @staticmethod
    def init(p: Params, rank: int = 0, nranks: int = 1, cache_dir: str | None = None) -> "TfmKit":
        return TfmKit.build_tables(p, cache_dir).upload(rank, nranks)


def finalize():
    check(_lib.lib().mlegs_b200_finalize())
