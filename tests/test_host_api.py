"""The reference's own unit tests for this path, re-stated against fdtd_b200's host layer
(reference tests/test_grid.py:17-44, test_boundaries.py:19-54, test_sources.py:17-43,
test_detectors.py:91-107, test_backend.py).  CPU: the host layer runs against the serial-interpreter
build of the kernels (tests/emu)."""
import numpy as np
import pytest

from emu.harness import use_emu


@pytest.fixture
def fd():
    return use_emu("float64")


@pytest.fixture
def grid(fd):
    return fd.Grid(shape=(10, 10, 10), grid_spacing=100e-9, permittivity=1.0, permeability=1.0)


# --- tests/test_grid.py -----------------------------------------------------------------------
def test_grid_shape_of_ints(fd):
    g = fd.Grid(shape=(3, 3, 3))
    assert (g.Nx, g.Ny, g.Nz) == (3, 3, 3)


def test_grid_shape_of_floats(fd):
    g = fd.Grid(shape=(10.0e-9, 10.0e-9, 10.0e-9), grid_spacing=5.0e-9)
    assert (g.Nx, g.Ny, g.Nz) == (2, 2, 2)


def test_grid_shape_mix_of_floats_and_ints(fd):
    g = fd.Grid(shape=(10.0e-9, 10.0e-9, 3), grid_spacing=5.0e-9)
    assert (g.Nx, g.Ny, g.Nz) == (2, 2, 3)


def test_default_courant_numbers(fd):
    assert fd.Grid(shape=(3, 1, 1)).courant_number == pytest.approx(1.0, rel=0.02)
    assert fd.Grid(shape=(3, 3, 1)).courant_number == pytest.approx((1.0 / 2.0) ** 0.5, rel=0.02)
    assert fd.Grid(shape=(3, 3, 3)).courant_number == pytest.approx((1.0 / 3.0) ** 0.5, rel=0.02)
    with pytest.raises(ValueError):
        fd.Grid(shape=(3, 3, 3), courant_number=0.9)


def test_curl_golden_vectors(fd):
    import os
    import torch
    from fdtd_b200.grid import curl_E, curl_H
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "curls.npz"))
    assert np.array_equal(curl_E(torch.from_numpy(g["E"])).numpy(), g["curl_E"])
    assert np.array_equal(curl_H(torch.from_numpy(g["H"])).numpy(), g["curl_H"])


def test_fields_are_views(grid):
    assert tuple(grid.E.shape) == (10, 10, 10, 3)
    grid.E[1, 2, 3, 2] = 5.0
    assert float(grid._E[2, 2, 2, 3]) == 5.0          # SoA storage with one ghost plane
    grid.E *= 0
    assert float(grid.E.abs().max()) == 0.0
    grid.H = np.ones((10, 10, 10, 3))
    assert float(grid.H.sum()) == 3000.0
    grid.reset()
    assert float(grid.H.abs().max()) == 0.0 and grid.time_steps_passed == 0


# --- tests/test_boundaries.py -------------------------------------------------------------------
def test_periodic_boundary_in_grid_boundary_list(fd, grid):
    pb = fd.PeriodicBoundary(name="pb")
    grid[0, :, :] = pb
    assert pb in grid.boundaries and grid.pb is pb


def test_periodic_boundary_raises_error_when_indexed_with_slice(fd, grid):
    with pytest.raises(IndexError):
        grid[0:2, :, :] = fd.PeriodicBoundary()


def test_periodic_boundary_placed_in_middle_of_grid(fd, grid):
    with pytest.raises(IndexError):
        grid[2, :, :] = fd.PeriodicBoundary()


def test_pml_in_grid_boundary_list(fd, grid):
    pml = fd.PML(name="PML")
    grid[0:3, :, :] = pml
    assert pml in grid.boundaries and pml.thickness == 3 and pml.axis == 0


def test_pml_placed_in_middle_of_grid(fd, grid):
    with pytest.raises(IndexError):
        grid[2:4, :, :] = fd.PML()


def test_two_boundaries_on_one_face(fd, grid):
    grid[0:3, :, :] = fd.PML()
    with pytest.raises(AttributeError):
        grid[0:2, :, :] = fd.PML()
    with pytest.raises(AttributeError):
        grid[0, :, :] = fd.PeriodicBoundary()


def test_duplicate_names_are_refused(fd, grid):
    grid[0:3, :, :] = fd.PML(name="edge")
    with pytest.raises(ValueError):
        grid[5, 5, 5] = fd.PointSource(name="edge")


def test_DomainBorderPML(fd, grid):
    with pytest.raises(IndexError):
        fd.DomainBorderPML(grid, grid.Nx // 2 + 1)
    fd.DomainBorderPML(grid, 3)
    assert len(grid.boundaries) == 6
    grid.run(3, progress_bar=False)


# --- tests/test_sources.py ------------------------------------------------------------------------
def test_PlaneSource_polarization_error(fd, grid):
    with pytest.raises(ValueError):
        grid[0, :, :] = fd.PlaneSource(polarization="x")
    with pytest.raises(ValueError):
        grid[:, 0, :] = fd.PlaneSource(polarization="y")
    with pytest.raises(ValueError):
        grid[:, :, 0] = fd.PlaneSource(polarization="z")


def test_PlaneSource_polarization_inference(fd, grid):
    for key, pol, e, h in (((0, slice(None), slice(None)), "y", 1, 2), ((0, slice(None), slice(None)), "z", 2, 1),
                           ((slice(None), 0, slice(None)), "x", 0, 2), ((slice(None), 0, slice(None)), "z", 2, 0),
                           ((slice(None), slice(None), 0), "x", 0, 1), ((slice(None), slice(None), 0), "y", 1, 0)):
        ps = fd.PlaneSource(polarization=pol)
        grid[key] = ps
        assert ps._Epol == e and ps._Hpol == h


def test_point_source_needs_a_single_cell(fd, grid):
    with pytest.raises(ValueError):
        grid[2:4, 3, 3] = fd.PointSource()


# --- tests/test_detectors.py ------------------------------------------------------------------------
def test_CurrentDetector_shape(fd, grid):
    edetector = fd.BlockDetector()
    cdetector = fd.CurrentDetector()
    grid[4, 4, 4] = cdetector
    grid[4, 4, 5] = edetector
    grid.run(10, progress_bar=False)
    assert np.array(cdetector.I).shape[0:2] == np.array(edetector.E).shape[0:2]
    assert np.array(cdetector.I).shape[0] == 10


def test_SAPS_detector_register(fd, grid):
    source = fd.SoftArbitraryPointSource(np.zeros(1), impedance=50.0)
    grid[4, 4, 4] = source
    grid.run(100, progress_bar=False)
    assert isinstance(grid.detectors[0], fd.CurrentDetector)
    assert len(source.source_voltage) == 100 and len(source.input_voltage) == 100


def test_detector_values_and_line_detector_shapes(fd, grid):
    det = fd.LineDetector(name="det")
    grid[2:8, 5, 5] = det
    blk = fd.BlockDetector()
    grid[2:3, 2:4, 5:5] = blk                      # inclusive ranges: 2 x 3 x 1 points
    grid[5, 5, 5] = fd.PointSource(period=10)
    grid.run(7, progress_bar=False)
    vals = det.detector_values()
    assert len(vals["E"]) == 7 and vals["E"][0].shape == (6, 3)
    assert np.array(blk.H).shape == (7, 2, 3, 1, 3)
    with pytest.raises(IndexError):
        grid[2:, 2:4, 5] = fd.BlockDetector()      # open upper bound -> index Nx (fdtd/detectors.py:236)


# --- tests/test_backend.py ---------------------------------------------------------------------------
def test_backend_names():
    import torch
    import fdtd_b200
    from fdtd_b200.backend import _NAMES
    assert set(_NAMES) >= {"cuda", "cuda.float32", "cuda.float64"}
    with pytest.raises(ValueError):
        fdtd_b200.set_backend("cuda.float16")
    with pytest.raises(ValueError):
        fdtd_b200.set_backend("numpy")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            fdtd_b200.set_backend("cuda")


def test_str_of_grid(fd, grid):
    grid[0:3, :, :] = fd.PML(name="pml")
    grid[5, 5, 5] = fd.PointSource(period=10, name="src")
    grid[2:8, 5, 5] = fd.LineDetector(name="det")
    grid[4:6, 4:6, 4:6] = fd.Object(permittivity=2.0, name="obj")
    s = str(grid)
    assert s.startswith("Grid(shape=(10,10,10), grid_spacing=1.00e-07, courant_number=0.57)")
    for part in ("sources:", "detectors:", "boundaries:", "objects:", "PointSource(period=10", "@ x=0:3, y=:, z=:",
                 "Object(name='obj')", "@ x=4:6, y=4:6, z=4:6"):
        assert part in s, part


def test_save_data_npz_layout(fd, grid, tmp_path, monkeypatch):
    """reference fdtd/grid.py:490-514: one "<name> (E)" and "<name> (H)" array per detector."""
    monkeypatch.chdir(tmp_path)
    grid[2:8, 5, 5] = fd.LineDetector(name="det")
    grid[5, 5, 5] = fd.PointSource(period=10)
    with pytest.raises(Exception):
        grid.save_data()
    folder = grid.save_simulation("unit")
    grid.run(6, progress_bar=False)
    grid.save_data()
    data = np.load(folder + "/detector_readings.npz")
    assert set(data.files) == {"det (E)", "det (H)"}
    assert data["det (E)"].shape == (6, 6, 3)


def test_generate_video_needs_a_folder_and_ffmpeg(fd, grid, tmp_path, monkeypatch):
    """reference fdtd/grid.py:441-488: frames in the simulation folder -> ffmpeg; clear errors otherwise."""
    with pytest.raises(Exception):
        grid.generate_video()
    monkeypatch.chdir(tmp_path)
    folder = grid.save_simulation("video")
    open(folder + "/file0000.png", "wb").close()
    calls = []
    import subprocess
    monkeypatch.setattr(subprocess, "check_call", lambda cmd, cwd=None: calls.append((cmd, cwd)))
    name = grid.generate_video(delete_frames=True)
    assert name.startswith("fdtd_sim_video_") and name.endswith("(video).mp4")
    assert calls[0][0][0] == "ffmpeg" and calls[0][1] == folder and "file%04d.png" in calls[0][0]
    import os
    assert not os.path.exists(folder + "/file0000.png")

    def missing(cmd, cwd=None):
        raise FileNotFoundError("ffmpeg")
    monkeypatch.setattr(subprocess, "check_call", missing)
    with pytest.raises(RuntimeError):
        grid.generate_video()


def test_zero_step_run_and_empty_detector(fd, grid):
    grid[3:3, 4, 4] = fd.LineDetector(name="empty")        # zero points
    grid.run(0, progress_bar=False)
    assert grid.time_steps_passed == 0 and grid.empty.E == []
    grid.run(2.5 * grid.time_step, progress_bar=False)      # float seconds truncate to steps (fdtd/grid.py:259-261)
    assert grid.time_steps_passed == 2
    assert np.array(grid.empty.E).shape == (2, 0, 3)


def test_partition_balances_plane_cost():
    """x-slabs: equal plane counts by default, equal cumulative cost when planes are priced (x-PML planes move
    13 instead of 9 words per cell and half-step)."""
    from fdtd_b200.sharding import Partition

    class P(Partition):
        def __init__(self, Nx, world, cost=None):
            self.Nx, self.world, self.rank = Nx, world, 0
            self._cuts = self._balanced_cuts(cost)

    assert P(1024, 8)._cuts == [0, 128, 256, 384, 512, 640, 768, 896, 1024]
    assert P(10, 3)._cuts == [0, 4, 7, 10]
    cost = [13 / 9 if (i < 10 or i >= 1014) else 1.0 for i in range(1024)]
    cuts = P(1024, 8, cost)._cuts
    sizes = [b - a for a, b in zip(cuts, cuts[1:])]
    assert sizes[0] == sizes[-1] < sizes[3] and sum(sizes) == 1024
    work = [sum(cost[a:b]) for a, b in zip(cuts, cuts[1:])]
    assert max(work) - min(work) < 1.5                                   # within one plane of each other
    assert P(5, 5, [1, 1, 10, 1, 1])._cuts == [0, 1, 2, 3, 4, 5]         # never an empty slab
    with pytest.raises(ValueError):
        P(4, 2, [1, 1, 1])
