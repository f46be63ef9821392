"""Shared helpers of the parity tests: one set of tables feeds both the device library and the oracle."""
from __future__ import annotations

import numpy as np

from oracle import mlegs_oracle as mo


def oracle_params(p) -> mo.Params:
    return mo.Params(nr=p.nr, np=p.np, nz=p.nz, nrchop=p.nrchop, npchop=p.npchop, nzchop=p.nzchop, ell=p.ell,
                     zlen=p.zlen, visc=p.visc, hyperpow=p.hyperpow, hypervisc=p.hypervisc, is_svv=bool(p.is_svv),
                     svv_cutoff=p.svv_cutoff, svv_target=p.svv_target, svv_strength=p.svv_strength,
                     svv_relax=p.svv_relax)


def oracle_kit(kit) -> mo.Kit:
    """Oracle kit sharing the device kit's tables bit for bit (SURVEY.md section 7 'Tables')."""
    return mo.kit_init(oracle_params(kit.params), tables=kit.tables())


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    nb = np.linalg.norm(b)
    err = float(np.linalg.norm(a - b) / (nb if nb > 0 else 1.0))
    _log_parity(err)
    return err


def _log_parity(err: float):
    """MLEGS_PARITY_LOG=<file>: append every measured relative error with the test and source line that asked for it
    (the evidence behind the tolerances written in the tests; profiles/r2/parity_errors.jsonl)."""
    import os
    path = os.environ.get("MLEGS_PARITY_LOG")
    if not path:
        return
    import inspect
    import json
    fr = inspect.stack()[2]
    rec = {"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "where": f"{os.path.basename(fr.filename)}:{fr.lineno}",
           "code": (fr.code_context[0].strip() if fr.code_context else ""), "rel_l2": err}
    with open(path, "a") as fh:
        fh.write(json.dumps(rec) + "\n")


def random_fff(okit: mo.Kit, seed: int = 0, decay: float = 8.0) -> np.ndarray:
    """SURVEY.md section 8d input 2: iid N(0,1)+iN(0,1) coefficients scaled by exp(-(n/nrchop)^2*decay), chopped."""
    rng = np.random.default_rng(seed)
    shp = okit.glb_sz
    e = rng.standard_normal(shp) + 1j * rng.standard_normal(shp)
    n = np.arange(shp[0])[:, None, None]
    e = e * np.exp(-((n / okit.p.nrchop) ** 2) * decay)
    s = mo.Scalar(e=np.asfortranarray(e), space="FFF")
    mo.chop(s, okit)
    return s.e


def random_ppp(okit: mo.Kit, seed: int = 0) -> np.ndarray:
    """A smooth physical field: backward transform of random_fff (padding column/rows as the reference leaves them)."""
    s = mo.Scalar(e=random_fff(okit, seed), space="FFF")
    # make it the spectrum of a real field in z so the packed-phi layout is meaningful either way
    mo.trans(s, "PPP", okit)
    return s.e
