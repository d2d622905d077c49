"""bench.py end to end on a GPU with the driver's own arguments (`--steps 20 --warmup 5`), on a small mesh."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + extra, capture_output=True, text=True, timeout=900,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_bench_runs_with_the_drivers_arguments():
    d = _run(["--gpus", "1", "--steps", "20", "--warmup", "5", "--n", "16", "--cpu-budget", "3"])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["steps"] == 20 and d["warmup"] == 5 and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["value"] <= d["value"] * 1.0001
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 20 * 50 and d["all_steps_converged"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["achieved"] > 0 and rf["peak"] > 1000 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == len(os.sched_getaffinity(0)) and cb["value"] > 0
    assert set(d["config"]["newton_steps_per_sec_on_sample_meshes"]) == {"12", "16", "20", "24", "28", "32"}


def test_bench_outlasting_the_dt_schedule():
    d = _run(["--steps", "2", "--warmup", "40", "--n", "8", "--no-cpu-baseline", "--no-same-config"])
    assert d["all_steps_converged"] and d["config"]["newton_iters"] >= 2


def test_bench_config_5_line():
    """BASELINE.json configs[4] as a second bench line: HCP KMBalD, B-bar + EA + NRLS, cyclic loading (small mesh here)"""
    d = _run(["--config", "5", "--steps", "3", "--warmup", "10", "--n", "12", "--krylov-iter", "600"])
    assert d["config"]["baseline_config"] == 5 and "HCP" in d["config"]["workload"] and d["all_steps_converged"]
    assert d["roofline"]["kernel"].startswith("k_ea_mult_p") and d["roofline"]["achieved"] > 0
    assert d["config"]["model_setups"] >= 3 * d["config"]["newton_iters"]      # line search: 3 residuals per iteration
