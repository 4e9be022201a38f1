"""Seeded random scenes (tests/fuzz_scenes.py): the kernels (serial-interpreter build, CPU) against the oracle.
Bit-equality, except where three objects (or a plain and an anisotropic one) share a cell."""
import numpy as np
import pytest

import scenes
from emu.harness import use_emu  # noqa: E402
from fuzz_scenes import random_scene
from oracle import yee_oracle as yo


def inexact_overlaps(g):
    """Two objects on one cell are reproduced exactly whatever their kinds (first and second coefficient layer);
    three on one cell only to rounding (the third is summed into the second layer)."""
    objs = g.objects
    boxes = [(o.x, o.y, o.z) for o in objs]

    def meet(*bs):
        return all(max(s.start for s in axis) < min(s.stop for s in axis) for axis in zip(*bs))

    n = len(objs)
    return any(meet(boxes[a], boxes[b], boxes[c])
               for a in range(n) for b in range(a + 1, n) for c in range(b + 1, n))


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
    exact = not inexact_overlaps(o)
    for k in want:
        assert got[k].shape == want[k].shape, k
        finite = np.isfinite(want[k]).all()
        assert finite, f"{k}: the scene blew up in the oracle"
        err = scenes.rel_l2(got[k], want[k])
        assert err <= (1e-12 if dtype == "float64" else 1e-5), f"seed {seed} {k}: rel-L2 {err:.3e}"
        if exact:
            assert np.array_equal(got[k], want[k]), f"seed {seed} {k}: not bit-identical ({err:.3e})"
