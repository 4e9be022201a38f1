"""Seeded random scenes (tests/fuzz_scenes.py): the kernels (serial-interpreter build, CPU) against the oracle.
Bit-equality throughout, however many objects of whatever kinds share a cell."""
import numpy as np
import pytest

import scenes
from emu.harness import use_emu  # noqa: E402
from fuzz_scenes import random_scene
from oracle import yee_oracle as yo


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("seed", range(24))
def test_random_scene(seed, dtype):
    build, steps = random_scene(seed)
    fd = use_emu(dtype)
    g = build(fd)
    g.run(steps // 2, progress_bar=False)
    for _ in range(steps - steps // 2):
        g.step()
    got = scenes.dump(g)
    yo.set_backend("numpy" if dtype == "float64" else "torch", dtype)
    try:
        o = build(yo)
        o.run(steps)
        want = scenes.dump(o)
    finally:
        yo.set_backend("numpy", "float64")
    for k in want:
        assert got[k].shape == want[k].shape, k
        finite = np.isfinite(want[k]).all()
        assert finite, f"{k}: the scene blew up in the oracle"
        err = scenes.rel_l2(got[k], want[k])
        assert err <= (1e-12 if dtype == "float64" else 1e-5), f"seed {seed} {k}: rel-L2 {err:.3e}"
        assert np.array_equal(got[k], want[k]), f"seed {seed} {k}: not bit-identical ({err:.3e})"


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("seed", range(10))
def test_random_fused_scene(seed, dtype, monkeypatch):
    """the single-pass E+H kernel (TMA-staged tiles and z-slab psi, CPML tables in shared memory, early x-slab psi,
    non-uniform x-chunks, split launches) against the two half-steps on random eligible scenes: bit-identical."""
    from fuzz_scenes import random_fused_scene
    build, steps, x_chunk, split = random_fused_scene(seed)
    fd = use_emu(dtype)
    outs = []
    for fuse in (1, 0):
        monkeypatch.setenv("FDTD_B200_FUSE_SPLIT_TEST", "1" if (fuse and split) else "0")
        g = build(fd)
        g._fuse_eh = fuse
        g._x_chunk = x_chunk
        g.run(steps, progress_bar=False)
        g.step()
        g.run(2, progress_bar=False)
        assert bool(g._engine.lib.fdtd_fuse_eh_active(g._engine.desc)) == bool(fuse), "the scene must be eligible"
        outs.append(scenes.dump(g))
    monkeypatch.delenv("FDTD_B200_FUSE_SPLIT_TEST")
    assert float(np.abs(outs[1]["E"]).max()) > 0
    for k in outs[1]:
        assert np.array_equal(outs[0][k], outs[1][k]), f"seed {seed} {k}: rel-L2 {scenes.rel_l2(outs[0][k], outs[1][k]):.3e}"
