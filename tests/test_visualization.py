"""fdtd_b200.visualization (fdtd/visualization.py): the numerical parts against plain numpy on the same
fields, the outline primitives against the registered scene, and the matplotlib part against a recording
stand-in for pyplot (matplotlib is not in this image; the reference's own test file for it is empty)."""
import sys
import types

import numpy as np
import pytest

import scenes
from emu.harness import use_emu
from test_sharded_gloo import launch


def _ran(fd, builder=scenes.objects3d, steps=20):
    g = builder(fd)
    g.run(steps, progress_bar=False)
    return g


def test_energy_slice_matches_full_grid_energy():
    fd = use_emu("float64")
    from fdtd_b200.visualization import energy_slice
    g = _ran(fd)
    E, H = g.E.cpu().numpy(), g.H.cpu().numpy()
    energy = (E ** 2 + H ** 2).sum(-1)
    assert np.array_equal(energy_slice(g, x=7), energy[7, :, :])
    assert np.array_equal(energy_slice(g, y=9), energy[:, 9, :].T)
    assert np.array_equal(energy_slice(g, z=-3), energy[:, :, -3])
    assert energy_slice(g, z=4).max() > 0
    with pytest.raises(ValueError):
        energy_slice(g)
    with pytest.raises(ValueError):
        energy_slice(g, x=1, y=2)
    with pytest.raises(ValueError):
        energy_slice(g, x=1.5)
    with pytest.raises(IndexError):
        energy_slice(g, x=1000)


def test_energy_slice_sharded(tmp_path):
    out = str(tmp_path / "sharded.npz")
    launch(2, "gloo", "float64", "objects3d", 20, out, FDTD_TEST_SLICES="7,9,4")
    got = dict(np.load(out))
    energy = (got["E"] ** 2 + got["H"] ** 2).sum(-1)
    assert np.array_equal(got["slice_x"], energy[7])
    assert np.array_equal(got["slice_y"], energy[:, 9, :].T)
    assert np.array_equal(got["slice_z"], energy[:, :, 4])


def test_scene_outline_primitives():
    fd = use_emu("float64")
    from fdtd_b200.visualization import scene_outline
    g = scenes.quickstart2d(fd)                           # 161 x 97 x 1: object, line source, detector, 4 PMLs, periodic z
    items = scene_outline(g, z=0)
    roles = [i["role"] for i in items]
    assert roles.count("pml") == 4 and roles.count("object") == 2
    assert roles.count("source") == 1 and roles.count("detector") == 1 and roles.count("periodic") == 0
    obj = next(i for i in items if i["role"] == "object")         # the first one, grid[11:32, 30:84, 0]
    assert obj["xy"] == (30 - 0.5, 11 - 0.5) and obj["width"] == 54 and obj["height"] == 21   # grid[11:32, 30:84, 0]
    low_x = [i for i in items if i["role"] == "pml" and i["xy"] == (-0.5, -0.5)]
    assert {(i["width"], i["height"]) for i in low_x} == {(97, 10), (10, 161)}
    g3 = scenes.periodic3d(fd)
    items = scene_outline(g3, x=3)                        # rows = y, cols = z: both periodic, x-PMLs not visible
    assert [i["role"] for i in items].count("periodic") == 2 and "pml" not in [i["role"] for i in items]
    block = next(i for i in scene_outline(scenes.pml3d(fd), z=1) if i["role"] == "detector" and len(i["h"]) == 5)
    assert block["h"][0] == block["h"][3] and block["v"][0] == block["v"][1]


class _Node:
    """an attribute of the stand-in: callable (the call is logged) and itself has attributes (plt.gca().add_patch,
    cbar.ax.set_ylabel)."""

    def __init__(self, name, log):
        self._name, self._log = name, log

    def __call__(self, *a, **kw):
        self._log.append((self._name, a, kw))
        return self

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Node(attr, self._log)


class _Recorder(types.ModuleType):
    """stands in for matplotlib.pyplot / patches / colors: records every call."""

    def __init__(self, name, log):
        super().__init__(name)
        self._log = log

    def __getattr__(self, attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Node(attr, self._log)


@pytest.fixture
def fake_pyplot(monkeypatch):
    log = []
    plt, ptc, colors = _Recorder("matplotlib.pyplot", log), _Recorder("matplotlib.patches", log), \
        _Recorder("matplotlib.colors", log)
    root = types.ModuleType("matplotlib")
    root.pyplot, root.patches, root.colors = plt, ptc, colors
    for name, mod in (("matplotlib", root), ("matplotlib.pyplot", plt), ("matplotlib.patches", ptc),
                      ("matplotlib.colors", colors)):
        monkeypatch.setitem(sys.modules, name, mod)
    return log


def test_visualize_draws_the_scene(fake_pyplot):
    fd = use_emu("float64")
    g = _ran(fd, scenes.pml3d, 10)
    fig = g.visualize(z=6, norm="log")
    calls = [c[0] for c in fake_pyplot]
    assert fig is not None and "imshow" in calls and "figlegend" in calls and "LogNorm" in calls
    shown = next(c for c in fake_pyplot if c[0] == "imshow")[1][0]
    assert shown.shape == (g.Nx, g.Ny) and (shown >= 0).all()
    assert calls.count("Rectangle") == 4                                   # x and y PMLs in a z-projection
    src = g.sources[0]
    assert shown[src.x, src.y] == 0                                        # source cell blanked
    with pytest.raises(ValueError):
        g.visualize(z=6, norm="sqrt")
    with pytest.raises(ValueError):
        g.visualize()


def test_db_map_and_arrivals(fake_pyplot):
    fd = use_emu("float64")
    from fdtd_b200.visualization import envelope_arrivals, peak_to_peak_dB
    g = fd.Grid(shape=(24, 20, 1), grid_spacing=100e-9)
    g[0:4, :, :] = fd.PML(); g[-4:, :, :] = fd.PML(); g[:, 0:4, :] = fd.PML(); g[:, -4:, :] = fd.PML()
    g[8, 10, 0] = fd.PointSource(period=12, pulse=True, cycle=3, hanning_dt=4.0, name="src")
    g[10:15, 8:12, 0] = fd.BlockDetector(name="block")
    g[12, 10:11, 0] = fd.LineDetector(name="near")
    g[18, 10:11, 0] = fd.LineDetector(name="far")
    g.run(90, progress_bar=False)
    rec = np.array(g.block.E)
    db = peak_to_peak_dB(rec, choose_axis=2)
    tr = rec[:, :, :, 0, 2]
    swing = tr.max(0) - tr.min(0)
    assert db.shape == (6, 5) and np.allclose(db, 10 * np.log10(swing / swing.min())) and db.min() == 0
    assert fd.dB_map_2D(rec, show=False) is not None
    with pytest.raises(ValueError):
        peak_to_peak_dB(rec[0])
    records = {"near (E)": np.array(g.near.E), "far (E)": np.array(g.far.E), "block (E)": rec}
    arr = envelope_arrivals(records, specific_plot="Ez", verbose=False)
    got = {name: step for name, _, step in arr["E"][2]}
    assert set(got) == {"near (E)", "far (E)"} and got["far (E)"] > got["near (E)"]   # the pulse arrives later
    assert fd.plot_detection(records, show=False) is not None
