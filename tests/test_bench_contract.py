"""bench.py's contract, as far as it can be exercised without a GPU: the reference arm (the oracle port on the host
cores) prints ONE JSON line with the keys the driver reads and the same `config` object the native arm prints; the
native arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    env = dict(os.environ)
    env.pop("WORLD_SIZE", None)
    env.pop("RANK", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"] == "transform_roundtrip_gdofs" and d["unit"] == "GDOF/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    cfg = d["config"]
    assert "128x128x128" in cfg["workload"] and "configs[1]" in cfg["workload"]
    assert (cfg["nr"], cfg["np"], cfg["nz"]) == (128, 128, 128) and "model" not in cfg
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # whole-job throughput: fields x DOF / time
    dof = cfg["nr"] * cfg["np"] * cfg["nz"] * cfg["fields_per_step"]
    assert d["value"] == pytest.approx(dof / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)


def test_native_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("--gpus", "1", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")], "no result line may be printed"


def test_reference_arm_does_not_load_the_product_library():
    """The reference arm's record must be clean: it may not import mlegs_b200 or dlopen libmlegs_b200.so."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'];\n"
            "try:\n    runpy.run_path('bench.py', run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "assert not [m for m in sys.modules if m.split('.')[0] == 'mlegs_b200'], 'mlegs_b200 imported'\n"
            "maps = open('/proc/self/maps').read()\nassert 'libmlegs_b200' not in maps, 'product library mapped'\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
