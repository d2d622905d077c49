"""bench.py's reference arm runs on the CPU and must print exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-n", "12"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "newton_steps_per_sec" and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["value"] > 0
    # value is derived from exactly the timed steps: (Newton iterations / seconds on the sample) / slowdown
    c = d["config"]
    assert c["sample_mesh"] == [12, 12, 12] and c["newton_iters"] >= 2 and c["pcg_iters"] > c["newton_iters"]
    assert abs(c["newton_steps_per_sec_on_sample"] / c["slowdown_sample_to_full"] - d["value"]) <= 1e-12 * d["value"]
    assert c["pcg_per_newton_full"] <= 1000.0 and c["element_ratio"] == (128 / 12) ** 3


def test_reference_arm_uses_every_host_thread_under_a_launcher_env():
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not be throttled by it
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--ref-n", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) == d["config"]["host_threads"]
    # the other ranks exit 0 without work or output
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--ref-n", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_driver_arguments_are_accepted():
    """The driver runs `bench.py --gpus N --steps 20 --warmup 5` (round 1 died on a 20-entry dt schedule)."""
    sys.path.insert(0, ROOT)
    import bench
    import numpy as np
    for gpus in (1, 2, 4, 8):
        a = bench.parse_args(["--gpus", str(gpus), "--steps", "20", "--warmup", "5"])
        dts = bench.dt_schedule(a.warmup + a.steps)
        assert len(dts) == 25 and all(dt > 0 for dt in dts)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "exaconstit_goldens.npz"))
    assert np.array_equal(np.array(bench.DT_SCHEDULE), gold["custom_dt"])       # test/data/custom_dt.txt, all 40 rows
    long = bench.dt_schedule(100)                                               # a run may outlast the schedule
    assert len(long) == 100 and long[39:] == [bench.DT_SCHEDULE[-1]] * 61
    assert bench.grains_for(128) == (2000, 1282000) and bench.grains_for(32) == (100, 32100) and bench.grains_for(64)[0] == 500


def test_our_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0          # no CPU fallback: the CUDA path is the only path
    assert r.stdout.strip() == ""


def test_fingerprint_covers_the_driver_run():
    """bench.py compares every run with the stored 1-GPU history (parity_fingerprint): the driver's
    `--steps 20 --warmup 5` needs 25 stored steps of the default workload, each with six stress components."""
    import json
    import bench
    fp = json.load(open(bench.FINGERPRINT))
    n = 128
    g, _ = bench.grains_for(n)
    ent = fp["%d:%d:1000:0" % (n, g)]
    assert len(ent["avg_stress"]) >= 25 and len(ent["newton_iters"]) == len(ent["avg_stress"])
    assert all(len(row) == 6 for row in ent["avg_stress"])
