"""CPU tier: the reference arm of bench.py runs and prints the contract's JSON line -- with the numpy port forced,
and with the unmodified reference module when oracle/stage_reference.py has staged it (oracle/_ref travels with the
snapshot; /root/reference itself is never read at run time).  Small grid / sample so it takes seconds; no GPU involved."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _run(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--grid", "24",
                          "--steps", "2", "--warmup", "1", "--cpu-sample", "100000", *extra],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def _check_contract(d):
    assert d["impl"] == "reference" and d["metric"] == "interp queries/sec (value+grad)" and d["unit"] == "queries/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64" and d["steps"] == 2
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_string(24, "norm")      # the string the GPU arm prints too


def test_reference_arm_prints_contract_line_port():
    d = _run("--force-port")
    _check_contract(d)
    assert d["cpu_baseline"]["kind"] == "port"


def _staged():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        from stage_reference import module_path
    finally:
        sys.path.pop(0)
    return module_path(os.path.join(ROOT, "oracle", "_ref"))


@pytest.mark.skipif(_staged() is None, reason="oracle/_ref not staged (oracle/stage_reference.py needs /root/reference)")
def test_reference_arm_uses_the_unmodified_module_when_staged():
    d = _run()
    _check_contract(d)
    assert d["cpu_baseline"]["kind"] == "reference" and "oracle/_ref" in d["cpu_baseline"]["sample"]


@pytest.mark.skipif(_staged() is None, reason="oracle/_ref not staged")
def test_staged_module_is_the_reference_not_the_dropin_and_matches_golden():
    """bench.load_reference_module() must load the staged reference by path (the repository root holds drop-in packages
    of the same names), and that module reproduces a committed golden vector -- it is the code the goldens came from."""
    sys.path.insert(0, ROOT)
    import bench
    mod, path = bench.load_reference_module()
    assert mod is not None and os.sep + os.path.join("oracle", "_ref") + os.sep in path
    assert "arbinterp_b200" not in open(path).read()
    g = load_golden("tri_12x10x9")
    obj = mod.tricubic(g["field"].copy(), "quiet", mode="norm")
    q = g["norm_q_in"].copy()
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        norms, grads = obj.Query(q)
    assert np.array_equal(norms, g["norm_out0"], equal_nan=True) and np.array_equal(grads, g["norm_out1"], equal_nan=True)


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=60, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
