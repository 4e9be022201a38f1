"""The reference's OWN test-suite, collected where it lies, run against fdtd_b200 aliased as `fdtd` (tests/refshim.py).
Build container only: /root/reference does not exist on the GPU box (the test is skipped there)."""
import os
import re
import subprocess
import sys

import pytest

REF_TESTS = os.path.join(os.environ.get("FDTD_REFERENCE", "/root/reference"), "tests")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference is not mounted here")
def test_the_reference_suite_passes_against_this_package(tmp_path):
    env = dict(os.environ, PYTHONPATH=HERE + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-m", "pytest", REF_TESTS, "-q", "-p", "refshim", "-p", "no:cacheprovider",
                        "--rootdir", str(tmp_path)], cwd=str(tmp_path), env=env, capture_output=True, text=True,
                       timeout=900)
    tail = r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0, tail
    m = re.search(r"(\d+) passed", r.stdout)
    assert m and int(m.group(1)) >= 27, tail          # every non-slow test of the reference (the slow one needs CUDA + hours)
    assert "failed" not in r.stdout, tail


EXAMPLES = os.path.join(os.environ.get("FDTD_REFERENCE", "/root/reference"), "examples")
NOTEBOOKS = ["00-quick-start", "01-basic-example", "02-absorbing-object", "03-objects-of-arbitrary-shape",
             "04-performance-profiling", "05-lenses-and-analysing-lensing-actions",
             "06-GRIN-medium-and-analysing-refraction"]


@pytest.mark.skipif(not os.path.isdir(EXAMPLES), reason="the reference is not mounted here")
def test_the_reference_example_notebooks_run_unchanged(tmp_path):
    """every example notebook of the reference executes cell by cell against this package (04's line_profiler, like
    matplotlib, is a stand-in): grids, PMLs, periodic boundaries, objects of arbitrary shape built from hundreds of
    registrations, lenses, GRIN media, detectors, visualize / save_simulation / save_data / dB_map_2D /
    plot_detection calls (plotting goes to a stand-in)."""
    paths = [os.path.join(EXAMPLES, n + ".ipynb") for n in NOTEBOOKS]
    r = subprocess.run([sys.executable, os.path.join(HERE, "run_notebook.py")] + paths, cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=1200,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("\nOK ") + r.stdout.startswith("OK ") == len(NOTEBOOKS), r.stdout[-3000:]
