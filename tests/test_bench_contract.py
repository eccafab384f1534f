"""bench.py's line contract, the parts that run without a GPU: the reference arm (`--impl reference`, the
reference's own CPU implementation from oracle/_ref on a bounded sample) prints one JSON line with the keys the
driver reads, ranks other than 0 of a torchrun launch print nothing, both arms describe the workload with the
same `config` object, and `--span day` selects the one-day window of BASELINE.md section 3."""
import json
import os
import subprocess
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
reflib = pytest.importorskip("reflib")


def run_bench(*args, env=None):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, **(env or {})))
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_line():
    if not reflib.available(False):
        pytest.skip("oracle/_ref not present")
    out = run_bench("--impl", "reference", "--size", "tiny", "--steps", "3", "--warmup", "1")
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sim-days/s" and d["higher_is_better"] is True
    assert d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "sim-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["ms_per_step"] * 1e-3 * 3 - (3 * 60.0 / 86400.0) / d["value"]) < 1e-9


def test_reference_arm_other_ranks_are_silent():
    assert run_bench("--impl", "reference", "--size", "tiny", "--steps", "1", "--warmup", "1",
                     env={"RANK": "1", "WORLD_SIZE": "2"}).strip() == ""


def test_both_arms_share_the_config_object_and_the_day_span():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config("1M", False, 1000000, 20422, 1)
    assert a == bench.workload_config("1M", False, 1000000, 20422, 1, "storm")
    assert "rain pulse" in a["workload"] and a["nsv"] == 3 * 1000000 + 2 * 20422
    args = types.SimpleNamespace(span="day", steps=20, warmup=5, no_strong=False, no_cpu=False)
    t0 = bench.T0
    try:
        bench.span_settings(args)
        assert (args.warmup, args.steps, bench.T0) == (60, 1380, 0.0) and args.no_strong and args.no_cpu
        assert "one simulated day" in bench.workload_config("1M", False, 1000000, 20422, 1, "day")["workload"]
    finally:
        bench.T0 = t0
