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


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference is not mounted here")
def test_grid_indexing_matches_the_reference():
    """differential fuzz of `grid[key] = thing` (ints, floats in metres, slices of both, index lists, negative
    values) for every plug-in type: the registered x / y / z and the exception types equal the reference's.
    Known, deliberate differences are excluded: a BlockDetector whose inclusive ranges leave the grid is refused at
    registration here (the reference fails with IndexError at the first step), and Objects with reversed / empty
    ranges (the reference's behaviour there depends on incidental numpy broadcasting)."""
    import random
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    from emu.harness import use_emu
    import numpy as np
    ref = make_golden.load_reference()
    ref.set_backend("numpy")
    fd = use_emu("float64")
    rnd = random.Random(7)
    N, sp = (12, 9, 7), 50e-9

    def rand_index(n):
        kind = rnd.choice(["int", "float", "slice", "list"])
        if kind == "int":
            return rnd.randrange(-n, n)
        if kind == "float":
            return rnd.uniform(0, (n - 1) * sp)
        if kind == "slice":
            return slice(rnd.choice([None, rnd.randrange(-n, n), rnd.uniform(0, (n - 1) * sp)]),
                         rnd.choice([None, rnd.randrange(-n, n + 1), rnd.uniform(0, n * sp)]))
        return [rnd.randrange(0, n) for _ in range(rnd.randrange(1, 4))]

    def describe(obj):
        out = []
        for a in "xyz":
            v = getattr(obj, a)
            out.append(("slice", v.start, v.stop) if isinstance(v, slice) else
                       ("list", [int(t) for t in v]) if isinstance(v, (list, tuple, np.ndarray)) else ("val", int(v)))
        return out

    makers = {"LineDetector": lambda m: m.LineDetector(), "BlockDetector": lambda m: m.BlockDetector(),
              "PointSource": lambda m: m.PointSource(period=10), "LineSource": lambda m: m.LineSource(period=10),
              "PlaneSource": lambda m: m.PlaneSource(period=10), "Object": lambda m: m.Object(permittivity=2.0),
              "PML": lambda m: m.PML(), "PeriodicBoundary": lambda m: m.PeriodicBoundary()}
    compared = 0
    for _ in range(160):
        name = rnd.choice(list(makers))
        key = tuple(rand_index(n) for n in N)
        res = []
        for m in (ref, fd):
            g = m.Grid(shape=N, grid_spacing=sp)
            try:
                thing = makers[name](m)
                g[key] = thing
                res.append(("ok", describe(thing)))
            except Exception as exc:     # noqa: BLE001
                res.append(("err", type(exc).__name__))
        if name == "BlockDetector" and res[0][0] == "ok" and res[1] == ("err", "IndexError"):
            continue
        if name == "Object" and (res[0][0] == "err" or any(d[0] == "slice" and d[2] <= d[1] for d in res[0][1])):
            continue
        compared += 1
        assert res[0] == res[1], f"{name} {key}: reference {res[0]}, here {res[1]}"
    assert compared > 110


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference is not mounted here")
def test_oracle_equals_the_reference_on_random_scenes():
    """beyond the committed golden scenes: the oracle against the live reference (numpy float64) on seeded random
    registrations (tests/fuzz_scenes.py) -- final E, H and every detector trace bit for bit.  (150 seeds were run
    during development; a dozen stay in the suite.)"""
    import numpy as np
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    import scenes
    from fuzz_scenes import random_scene
    from oracle import yee_oracle as yo
    ref = make_golden.load_reference()
    ref.set_backend("numpy")
    yo.set_backend("numpy", "float64")
    for seed in range(300, 312):
        build, steps = random_scene(seed)
        g = build(ref)
        g.run(steps, progress_bar=False)
        want = scenes.dump(g)
        o = build(yo)
        o.run(steps)
        got = scenes.dump(o)
        assert set(got) == set(want)
        for k in want:
            assert np.array_equal(got[k], want[k], equal_nan=True), f"seed {seed} {k}: {scenes.rel_l2(got[k], want[k]):.3e}"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference is not mounted here")
def test_user_plugins_against_the_live_reference():
    """the user-defined plug-ins of tests/scenes.py (written against the reference's duck-typed protocol only) on the
    unmodified reference and on this package: final fields, the built-in detector and the user's probe agree bit for
    bit -- the drop-in holds for plug-ins the package has never seen."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_golden
    import numpy as np
    import scenes
    from emu.harness import use_emu
    ref = make_golden.load_reference()
    ref.set_backend("numpy")
    r = scenes.user_plugins(ref)
    r.run(30, progress_bar=False)
    want = scenes.dump(r)
    g = scenes.user_plugins(use_emu("float64"))
    g.run(30, progress_bar=False)
    got = scenes.dump(g)
    assert g._engine._hooked
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(np.array(g.detectors[1].E), np.array(r.detectors[1].E))
    assert np.array_equal(np.array(g.detectors[1].H), np.array(r.detectors[1].H))
    assert float(np.abs(np.array(r.detectors[1].E)).max()) > 0
