"""bench.py's reference arm (CPU only): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, ACFB200_BENCH_CPU_PRODUCTS="4e8")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "lag-products/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") 
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == cores  # every host core runs the serial reference on its own slice
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["one_core_value"] > 1e8
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 1e8
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", ACFB200_BENCH_CPU_PRODUCTS="4e8")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "0"], capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""
