"""Scene builders shared by the golden generator, the oracle tests and the GPU parity tests.

Every builder takes `fd`, any module exposing the reference's public API
(`Grid`, `PML`, `PeriodicBoundary`, `Object`, ... ) -- the unmodified reference
(golden generation, build container only), `oracle.yee_oracle`, or the product
`fdtd_b200` -- and registers exactly the same things in exactly the same order.
Sizes are reductions of BASELINE.json's configs.
"""
import numpy as np


def _rand_materials(shape, seed=0):
    rs = np.random.RandomState(seed)
    eps = 1.0 + rs.rand(*shape, 3)
    mu = 1.0 + 0.5 * rs.rand(*shape, 3)
    return eps, mu


def quickstart2d(fd):
    """configs[0]: README quick-start, 161x97x1, fdtd README.md:137-297."""
    g = fd.Grid(shape=(25e-6, 15e-6, 1), grid_spacing=155e-9, permittivity=1.0, permeability=1.0)
    g[11:32, 30:84, 0] = fd.Object(permittivity=1.7 ** 2, name="object")
    g[13e-6:18e-6, 5e-6:8e-6, 0] = fd.Object(permittivity=1.5 ** 2)
    g[7.5e-6:8.0e-6, 11.8e-6:13.0e-6, 0] = fd.LineSource(period=1550e-9 / (3e8), name="source")
    g[12e-6, :, 0] = fd.LineDetector(name="detector")
    g[0:10, :, :] = fd.PML(name="pml_xlow")
    g[-10:, :, :] = fd.PML(name="pml_xhigh")
    g[:, 0:10, :] = fd.PML(name="pml_ylow")
    g[:, -10:, :] = fd.PML(name="pml_yhigh")
    g[:, :, 0] = fd.PeriodicBoundary(name="zbounds")
    return g


def pml3d(fd, n=(24, 22, 20), t=5):
    """configs[1] shape: six PMLs, PointSource, BlockDetector."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[0:t, :, :] = fd.PML(name="xl")
    g[-t:, :, :] = fd.PML(name="xh")
    g[:, 0:t, :] = fd.PML(name="yl")
    g[:, -t:, :] = fd.PML(name="yh")
    g[:, :, 0:t] = fd.PML(name="zl")
    g[:, :, -t:] = fd.PML(name="zh")
    c = [v // 2 for v in n]
    g[c[0], c[1], c[2]] = fd.PointSource(period=20, name="src")
    g[c[0] + 2:c[0] + 4, c[1]:c[1] + 2, c[2]:c[2] + 2] = fd.BlockDetector(name="det")
    return g


def objects3d(fd, n=(24, 20, 18), t=5):
    """configs[2] shape: random anisotropic eps/mu, PlaneSource, Absorbing + Anisotropic objects,
    an array-eps Object that overlaps a PML, a PointSource, Line + Block detectors."""
    eps, mu = _rand_materials(n, seed=0)
    g = fd.Grid(shape=n, grid_spacing=77.5e-9, permittivity=eps, permeability=mu)
    g[0:t, :, :] = fd.PML(name="xl")
    g[-t:, :, :] = fd.PML(name="xh")
    g[:, 0:t, :] = fd.PML(name="yl")
    g[:, -t:, :] = fd.PML(name="yh")
    g[:, :, 0:t] = fd.PML(name="zl")
    g[:, :, -t:] = fd.PML(name="zh")
    g[6, :, :] = fd.PlaneSource(period=20, polarization="z", name="plane")
    g[8:11, 6:14, 6:13] = fd.AbsorbingObject(permittivity=2.5, conductivity=1.5e4, name="absorber")
    rs = np.random.RandomState(3)
    lens = 1.0 + rs.rand(5, 8, 8, 3) * 1.5
    g[12:17, 6:14, 5:13] = fd.AnisotropicObject(permittivity=lens, name="lens")
    slab = 1.0 + rs.rand(4, 6, 7)
    g[18:22, 1:7, 9:16] = fd.Object(permittivity=slab, name="slab_in_pml")
    g[15, 3, 4] = fd.PointSource(period=14, amplitude=0.7, phase_shift=0.3, name="pt")
    g[2:22, 10, 9] = fd.LineDetector(name="line")
    g[13:15, 9:10, 8:10] = fd.BlockDetector(name="block")
    return g


def periodic3d(fd, n=(26, 14, 12), t=5):
    """configs[4] shape: x-PMLs + periodic y and z (one registered BEFORE the PMLs), GRIN object,
    pulsed LineSource, PointSource on a periodic plane, PlaneSource, both detectors."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[:, 0, :] = fd.PeriodicBoundary(name="ybounds")
    g[0:t, :, :] = fd.PML(name="xl")
    g[-t:, :, :] = fd.PML(name="xh")
    g[:, :, 0] = fd.PeriodicBoundary(name="zbounds")
    ramp = (1.0 + 1.25 * np.arange(n[1]) / (n[1] - 1.0)).reshape(1, n[1], 1)
    g[8:20, :, :] = fd.Object(permittivity=ramp, name="grin")
    g[6, :, :] = fd.PlaneSource(period=20, polarization="z", name="plane")
    g[10:16, 3:9, 5] = fd.LineSource(period=12, pulse=True, cycle=3, hanning_dt=4.0, name="pulse")
    g[12, n[1] - 1, 4] = fd.PointSource(period=9, name="on_periodic_plane")
    g[7:23, 7, 6] = fd.LineDetector(name="line")
    g[10:12, 0:1, 10:11] = fd.BlockDetector(name="block")
    return g


def vacuum_aniso(fd, n=(12, 10, 9)):
    """no boundaries: random anisotropic eps/mu cavity, PointSource, LineDetector."""
    eps, mu = _rand_materials(n, seed=5)
    g = fd.Grid(shape=n, grid_spacing=50e-9, permittivity=eps, permeability=mu)
    g[5, 4, 4] = fd.PointSource(period=11, name="pt")
    g[1:11, 5, 3] = fd.LineDetector(name="line")
    return g


def slab2d_xz(fd, n=(40, 1, 36), t=6):
    """a 2-D grid in the xz plane (Ny = 1) with PMLs and a dielectric block."""
    g = fd.Grid(shape=n, grid_spacing=100e-9)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[12:20, 0, 10:30] = fd.Object(permittivity=2.0, name="block")
    g[10, 0, 8:28] = fd.LineSource(period=16, name="src")
    g[30, 0, :] = fd.LineDetector(name="line")
    return g


def c4small(fd, n=(32, 32, 32), t=6):
    """configs[3] shape (the bench workload) reduced: six PMLs, centre PointSource, LineDetector."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[n[0] // 2, n[1] // 2, n[2] // 2] = fd.PointSource(period=20, name="src")
    g[2:n[0] - 2, n[1] // 2, n[2] // 2 + 3] = fd.LineDetector(name="line")
    return g


def feed50(fd, n=(18, 16, 14), t=4):
    """a SoftArbitraryPointSource with series impedance (gaussian-pulse voltage; Z = 0.5
    in simulation units -- the explicit one-step feedback is unstable from Z ~ 1 on this grid) with its paired
    CurrentDetector, a second CurrentDetector block away from it, an absorbing slab and a BlockDetector."""
    g = fd.Grid(shape=n, grid_spacing=1e-3)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()
    steps = np.arange(60)
    wave = np.exp(-((steps - 20.0) ** 2) / (2 * 6.0 ** 2)) * 1e-3
    # (the reference does not re-export this class at package level: fdtd/__init__.py:6-14)
    saps = getattr(fd, "SoftArbitraryPointSource", None) or fd.sources.SoftArbitraryPointSource
    g[9, 8, 7] = saps(wave, impedance=0.5)
    g[6:7, 9:10, 5:6] = fd.CurrentDetector(name="probe")
    g[11:13, 5:11, 5:9] = fd.AbsorbingObject(permittivity=2.0, conductivity=50.0)
    g[9:10, 8:9, 8:9] = fd.BlockDetector(name="block")
    return g


# name -> (builder, steps)
SCENES = {
    "quickstart2d": (quickstart2d, 300),
    "pml3d": (pml3d, 60),
    "objects3d": (objects3d, 60),
    "periodic3d": (periodic3d, 80),
    "vacuum_aniso": (vacuum_aniso, 40),
    "slab2d_xz": (slab2d_xz, 120),
    "c4small": (c4small, 50),
    "feed50": (feed50, 80),
}


def _np(a):
    """array / tensor / nested list of them -> numpy."""
    if isinstance(a, np.ndarray):
        return a
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    if isinstance(a, (list, tuple)):
        return np.stack([_np(v) for v in a]) if len(a) else np.zeros((0,))
    return np.asarray(a)


def dump(grid):
    """final fields and every detector trace as numpy arrays."""
    out = {"E": _np(grid.E), "H": _np(grid.H)}
    for n, det in enumerate(grid.detectors):
        if hasattr(det, "I"):
            out[f"det{n}_I"] = _np(det.I)
        else:
            out[f"det{n}_E"] = _np(det.E)
            out[f"det{n}_H"] = _np(det.H)
    for n, src in enumerate(grid.sources):
        if hasattr(src, "source_voltage"):
            out[f"src{n}_Vin"] = np.asarray(_np(src.input_voltage), dtype=np.float64).reshape(-1)
            out[f"src{n}_Vout"] = np.asarray(_np(src.source_voltage), dtype=np.float64).reshape(-1)
    return out


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.sqrt((b ** 2).sum())
    num = np.sqrt(((a - b) ** 2).sum())
    return num / den if den > 0 else num
