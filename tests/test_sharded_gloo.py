"""x-slab sharding + halo exchange on 2 and 3 ranks (gloo, CPU): sharded == unsharded, bit for bit.

The ranks run the same host code as on the GPUs (Partition, clipping of PML / objects / sources /
detectors to the slab, ghost planes, the split half-step with the exchange in between) against the
serial-interpreter build of the kernels (tests/emu, test infrastructure)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import scenes
from emu.harness import use_emu

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def launch(world, backend, dtype, scene, steps, out, **extra_env):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(free_port()), WORLD_SIZE=str(world),
               OMP_NUM_THREADS="1", **extra_env)
    procs = []
    for r in range(world):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_shard_worker.py"), backend, dtype,
                                       scene, str(steps), out], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]


CASES = [(2, "pml3d", "float64"), (2, "objects3d", "float64"), (3, "periodic3d", "float64"), (2, "c4small", "float32"),
         (3, "slab2d_xz", "float64"), (2, "overlaps3d", "float64"), (3, "overlaps3d", "float32"), (2, "ring3d", "float64"),
         (3, "ring3d", "float32"), (4, "ring3d", "float64"), (2, "feed50", "float64"), (3, "feed50", "float32"),
         (2, "objects3d", "float32x"), (3, "periodic3d", "float32x"), (2, "stacked3d", "float64"),
         (3, "stacked3d", "float32")]
STEPS = 24


@pytest.fixture(scope="module")
def sharded_runs(tmp_path_factory):
    """every case of one world size in ONE job of that many ranks (the ranks' start-up dominates a single case)"""
    cache = {}

    def get(world):
        if world not in cache:
            items = [(s, d) for w, s, d in CASES if w == world]
            out = str(tmp_path_factory.mktemp(f"world{world}") / "sharded")
            launch(world, "gloo", "-", ",".join(f"{s}@{d}" for s, d in items), STEPS, out)
            cache[world] = {item: dict(np.load(f"{out}.{n}.npz")) for n, item in enumerate(items)}
        return cache[world]
    return get


@pytest.mark.parametrize("world,scene,dtype", CASES)
def test_sharded_equals_single(sharded_runs, world, scene, dtype):
    got = sharded_runs(world)[(scene, dtype)]
    fd = use_emu(dtype)
    g = scenes.SCENES[scene][0](fd)
    assert not g._part.sharded
    g.run(STEPS, progress_bar=False)
    want = scenes.dump(g)
    assert set(got) == set(want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


@pytest.mark.parametrize("seed", [3, 7, 11, 19, 23, 42])
def test_sharded_random_scene(tmp_path, seed):
    """seeded random registrations (tests/fuzz_scenes.py) on 2 ranks: sharded == unsharded, bit for bit."""
    from fuzz_scenes import random_scene
    build, steps = random_scene(seed)
    out = str(tmp_path / "sharded.npz")
    launch(2, "gloo", "float64", f"fuzz:{seed}", steps, out)
    got = dict(np.load(out))
    if "skipped" in got:
        pytest.skip(f"not shardable: {got['skipped']}")
    fd = use_emu("float64")
    g = build(fd)
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


def test_current_detector_on_a_slab_boundary(tmp_path):
    """feed50 on 2 ranks cuts the grid at x = 9, exactly where the impedance port and its CurrentDetector sit: the
    loop of H around that cell reaches into the left neighbour's last plane of the SAME half-step, so the ranks
    exchange the H ghost plane before sampling.  Bit-identical to the unsharded run."""
    steps = 40
    out = str(tmp_path / "sharded.npz")
    launch(2, "gloo", "float64", "feed50", steps, out, FDTD_TEST_EXPECT_LATE="1")
    got = dict(np.load(out))
    fd = use_emu("float64")
    g = scenes.feed50(fd)
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    assert float(np.abs(want["det0_I"]).max()) > 0
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


def test_sharded_running_dft(tmp_path):
    """the device-side running DFT of detectors whose points are spread over the ranks: gathered spectra
    equal the unsharded ones bit for bit (every point accumulates its own record in order)."""
    steps, scene = 24, "objects3d"
    out = str(tmp_path / "sharded.npz")
    launch(2, "gloo", "float64", scene, steps, out, FDTD_TEST_TRACK="1")
    got = dict(np.load(out))
    fd = use_emu("float64")
    g = scenes.SCENES[scene][0](fd)
    scenes.track_all(g, steps)
    g.run(steps, progress_bar=False)
    want = scenes.dump_tracked(g)
    assert want and all(k in got for k in want)
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert np.abs(want[k]).max() > 0, k
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


@pytest.mark.parametrize("world", [2, 3])
def test_ring_capacity_is_the_same_on_every_rank(tmp_path, world):
    """a detector whose points all live on one slab, with detector rings so small that a run flushes them several
    times: every flush is collective, so every rank must reach it at the same step (the capacity was
    derived from the rank-local point count and the ranks deadlocked)."""
    steps = 70
    out = str(tmp_path / "sharded.npz")
    # 40 points x 2 fields x 3 components x 8 bytes = 1920 bytes per step on the owning rank
    launch(world, "gloo", "float64", "slabdet", steps, out, FDTD_TEST_RING_BYTES=str(1920 * 20))
    got = dict(np.load(out))
    fd = use_emu("float64")
    g = scenes.slabdet(fd)
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    assert float(np.abs(want["det0_E"]).max()) > 0
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"


@pytest.mark.parametrize("world", [2, 3])
def test_current_detector_on_global_plane_zero(tmp_path, world):
    """a CurrentDetector cell at x = 0 reads H[x-1] = H[-1] (python wrap-around, fdtd/detectors.py:432-447): on an
    x-sharded grid that plane lives on the LAST slab and travels into the first slab's low ghost before sampling."""
    steps = 30
    out = str(tmp_path / "sharded.npz")
    launch(world, "gloo", "float64", "current_x0", steps, out, FDTD_TEST_EXPECT_LATE="1")
    got = dict(np.load(out))
    assert "skipped" not in got, got
    fd = use_emu("float64")
    g = scenes.current_x0(fd)
    g.run(steps, progress_bar=False)
    want = scenes.dump(g)
    assert float(np.abs(want["det0_I"]).max()) > 0
    for k in want:
        assert np.array_equal(got[k], want[k]), f"{k}: rel-L2 {scenes.rel_l2(got[k], want[k]):.3e}"
    # ... and the unsharded run is the reference's (oracle, bit for bit)
    from oracle import yee_oracle as yo
    yo.set_backend("numpy", "float64")
    o = scenes.current_x0(yo)
    o.run(steps)
    ref = scenes.dump(o)
    for k in ref:
        assert np.array_equal(want[k], ref[k]), k
