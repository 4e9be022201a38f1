"""bench.py host logic that needs no GPU: the per-config algorithmic bytes and the
`--impl reference` arm's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_per_config():
    import bench
    want = {"c2": 159.0, "c3": 77.0, "c4": 73.875, "c5": 78.3}          # SURVEY 8d, bytes per cell-step
    for name, b in want.items():
        wl = bench.Workload(name, 8 if name == "c5" else 1)
        assert wl.bytes_per_cell_step == pytest.approx(b, rel=2e-3), (name, wl.bytes_per_cell_step)
    assert bench.Workload("c5", 8).shape == (2048, 1024, 1024)
    assert bench.Workload("c5", 2).scaling == "weak" and bench.Workload("c4", 2).scaling == "strong"
    assert bench.algorithmic_bytes_per_cell_step(1024, 4) == pytest.approx(73.875)


@pytest.mark.parametrize("config", ["c1", "c4", "c5"])
def test_reference_arm_prints_one_json_line(config):
    env = dict(os.environ, OMP_NUM_THREADS="1")                # as under torchrun: the arm must not be starved by it
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config,
                        "--steps", "2", "--warmup", "0", "--cpu-size", "24"], capture_output=True, text=True, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "Mcell-updates/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["config"] == config


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
