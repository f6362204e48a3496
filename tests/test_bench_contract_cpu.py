"""CPU: the committed bench lines (profiles/) carry every key of the driver's bench.py contract, with consistent arithmetic -
a guard against silently dropping a field when bench.py changes."""
import json
import os

import pytest

from tests.conftest import ROOT

LINES = ["r01_bench_v22_default.json", "r01_bench_v19_2gpu.json"]


@pytest.mark.parametrize("name", LINES)
def test_bench_line_contract(name):
    line = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k, t in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int), ("ms_per_step", float),
                 ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str), ("config", dict), ("e2e", dict),
                 ("gpu_launches", int), ("clocks", dict), ("roofline", dict)):
        assert isinstance(line[k], t), (k, type(line[k]))
    assert "vs_baseline" in line and line["vs_baseline"] is None                 # BASELINE.md holds no published number for this metric
    assert line["warmup"] >= 3 and line["higher_is_better"] is True and line["scaling"] == "weak" and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0 and line["e2e"]["value"] <= line["value"] * 1.02
    assert line["gpu_launches"] > 0
    assert set(line["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"])
    r = line["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # value = whole-job steps/s: n_gpus per-rank steps over the max-over-ranks time
    assert abs(line["value"] - line["n_gpus"] * 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"]
    if line["n_gpus"] == 1:
        c = line["cpu_baseline"]
        assert set(c) >= {"value", "unit", "cores", "kind", "sample"} and c["kind"] in ("reference", "port") and c["cores"] >= 1


def test_multi_rank_processes_load_cuda_modules_eagerly(monkeypatch):
    """WORLD_SIZE > 1: bench.py / the package set CUDA_MODULE_LOADING=EAGER before any CUDA context exists (lazy loading of a kernel
    variant's first launch inside an optimiser tail dead-locked against the peer-waiting all-reduce); single-rank
    runs leave the variable alone."""
    import importlib, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "import os, sys; sys.path.insert(0, %r); import bench; import comat_b200; print(os.environ.get('CUDA_MODULE_LOADING'))" % root
    for ws, want in (("2", "EAGER"), ("1", "None")):
        env = dict(os.environ, WORLD_SIZE=ws)
        env.pop("CUDA_MODULE_LOADING", None)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert out.stdout.strip().splitlines()[-1] == want, (ws, out.stdout, out.stderr[-500:])
