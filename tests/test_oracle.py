"""The oracle against the reference's golden vectors and against outputs of the reference itself.

CPU only.  `tests/golden/*.npz` were produced by tests/golden/make_golden.py from the
unmodified reference; `curls.npz` holds the reference's curl_E / curl_H on the seeds of its own
golden-vector tests (reference tests/test_grid.py:47-164), three literals of which are repeated
here verbatim as an independent pin.
"""
import glob
import os

import numpy as np
import pytest

import scenes
from oracle import yee_oracle as yo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_curl_golden_vectors():
    yo.set_backend("numpy", "float64")
    g = np.load(os.path.join(GOLD, "curls.npz"))
    E = np.random.RandomState(0).randn(3, 3, 3, 3)
    H = np.random.RandomState(1).randn(3, 3, 3, 3)
    assert np.array_equal(E, g["E"]) and np.array_equal(H, g["H"])
    cE, cH = yo.curl_E(E), yo.curl_H(H)
    assert np.array_equal(cE, g["curl_E"])
    assert np.array_equal(cH, g["curl_H"])
    # literals from the reference's test_curl_E (tests/test_grid.py:52-54, 57)
    assert cE[0, 0, 0] == pytest.approx([-0.99186526, -0.01377993, 2.48607585])
    assert cE[0, 0, 1] == pytest.approx([3.44005631, -1.38029691, -0.00954])
    assert cE[0, 1, 0] == pytest.approx([-3.98489477, 2.19203955, 1.15586708])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*_f64.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_bitwise_vs_reference_f64(path):
    scene = os.path.basename(path)[:-8]
    gold = np.load(path)
    yo.set_backend("numpy", "float64")
    build, steps = scenes.SCENES[scene]
    assert int(gold["steps"]) == steps
    g = build(yo)
    g.run(steps)
    out = scenes.dump(g)
    for k, v in out.items():
        assert v.shape == gold[k].shape, k
        assert np.array_equal(v, gold[k]), f"{scene}:{k} rel_l2={scenes.rel_l2(v, gold[k]):.3e}"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*_f32.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_vs_reference_true_f32(path):
    """torch flavour: bit-identical where the host's torch build matches the generator's;
    the stated float32 tolerance (1e-5 rel-L2) is the bar, bit-equality is reported."""
    scene = os.path.basename(path)[:-8]
    gold = np.load(path)
    yo.set_backend("torch", "float32")
    try:
        build, steps = scenes.SCENES[scene]
        g = build(yo)
        g.run(steps)
        out = scenes.dump(g)
    finally:
        yo.set_backend("numpy", "float64")
    for k, v in out.items():
        assert v.dtype == np.float32 or k.startswith("src")     # recorded source voltages are host floats
        assert scenes.rel_l2(v, gold[k]) <= 1e-5, k


def test_default_courant_and_shape():
    """reference tests/test_grid.py:17-44."""
    yo.set_backend("numpy", "float64")
    assert yo.Grid(shape=(3, 3, 3)).shape == (3, 3, 3)
    g = yo.Grid(shape=(10.0e-9, 10.0e-9, 3), grid_spacing=5.0e-9)
    assert g.shape == (2, 2, 3)
    assert yo.Grid(shape=(3, 1, 1)).courant_number == pytest.approx(1.0, rel=0.02)
    assert yo.Grid(shape=(3, 3, 1)).courant_number == pytest.approx(0.5 ** 0.5, rel=0.02)
    assert yo.Grid(shape=(3, 3, 3)).courant_number == pytest.approx((1 / 3) ** 0.5, rel=0.02)
