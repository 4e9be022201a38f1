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


def ring3d(fd, n=(20, 12, 14), t=4):
    """a ring along x: periodic x boundary registered BETWEEN a periodic y boundary and the z-PMLs (one PML
    after it -> a late correction on both sides of the wrap), sources and detectors on both wrap planes, an
    object reaching the last x plane.  On an x-sharded grid the copies E[0] = E[-1] / H[-1] = H[0] cross the
    first and the last slab."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[:, :, 0:t] = fd.PML(name="zl")
    g[:, 0, :] = fd.PeriodicBoundary(name="ybounds")
    g[0, :, :] = fd.PeriodicBoundary(name="xbounds")
    g[:, :, -t:] = fd.PML(name="zh")
    rs = np.random.RandomState(2)
    g[14:20, 3:9, 5:10] = fd.Object(permittivity=1.0 + rs.rand(6, 6, 5), name="block_at_the_end")
    g[2:5, 2:6, 5:9] = fd.AbsorbingObject(permittivity=1.5, conductivity=4e3, name="lossy")
    g[n[0] - 1, 6, 7] = fd.PointSource(period=10, name="on_last_plane")
    g[0, 3, 6] = fd.PointSource(period=13, amplitude=0.6, name="on_first_plane")
    g[6:15, 4:10, 7] = fd.LineSource(period=16, name="line_src")
    g[0:20, 7, 8] = fd.LineDetector(name="all_x")
    g[0:1, 5:6, 6:7] = fd.BlockDetector(name="first")
    g[n[0] - 2:n[0] - 1, 5:6, 6:7] = fd.BlockDetector(name="last")
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


def slabdet(fd, n=(24, 40, 10), t=4):
    """a 40-point LineDetector that lives on ONE x-plane (one rank of a sharded grid holds all of it, the others
    none) next to a one-point BlockDetector elsewhere: the ring capacity, and with it the collective flushes,
    must not depend on a rank's share of the points."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[n[0] // 2, n[1] // 2, n[2] // 2] = fd.PointSource(period=12)
    g[n[0] - 6, :, n[2] // 2] = fd.LineDetector(name="plane_line")
    g[3:3, 5:5, 2:2] = fd.BlockDetector(name="one_point")
    return g


def current_x0(fd, n=(18, 12, 10), t=3):
    """CurrentDetectors on global plane x = 0 (H[x-1] wraps to the LAST plane, fdtd/detectors.py:432-447 -- another
    slab when sharded) and on the first plane of an inner slab, next to a source that makes H differ there."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[:, 0:t, :] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[n[0] - 2, 6, 5] = fd.PointSource(period=9, name="near_the_end")
    g[1, 5, 4] = fd.PointSource(period=7, amplitude=0.6, name="near_the_start")
    g[0, 5, 5] = fd.CurrentDetector(name="on_plane_0")
    g[0:1, 4:6, 3:4] = fd.CurrentDetector(name="block_on_plane_0")
    g[n[0] // 2, 6, 5] = fd.CurrentDetector(name="on_a_cut")
    g[0:n[0], 6, 5] = fd.LineDetector(name="line")
    return g


def fusedslab(fd, n=(40, 30, 136), t=5):
    """a homogeneous grid (non-vacuum background) with six PMLs, sources in a slab corner, on a face and next to the
    middle x-planes (the cuts of a 2- or 4-rank partition), detectors across the cuts: what the temporally fused
    step of an x-sharded slab has to get right."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9, permittivity=1.3, permeability=1.1)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t + 1] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[n[0] // 2, n[1] // 2, n[2] // 2] = fd.PointSource(period=17, name="at_the_cut")
    g[n[0] // 2 - 1, n[1] // 2 + 2, n[2] // 3] = fd.PointSource(period=11, amplitude=0.4, name="left_of_the_cut")
    g[1, 2, 3] = fd.PointSource(period=13, amplitude=0.3, name="corner")
    g[n[0] // 4, n[1] // 2, 0] = fd.PointSource(period=9, amplitude=0.5, name="on_z_face")
    g[2:n[0] - 2, n[1] // 2 + 1, n[2] // 2 + 2] = fd.LineDetector(name="across")
    g[n[0] // 2 - 1:n[0] // 2, 3:5, n[2] - 3:n[2] - 2] = fd.BlockDetector(name="at_cut")
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


def overlaps3d(fd, n=(22, 20, 20), t=4):
    """every pair of object kinds on the same cells (the reference updates each object in registration
    order, fdtd/grid.py:285-287): absorber+absorber, anisotropic+absorber, plain+anisotropic,
    absorber+plain, anisotropic+anisotropic, some of them reaching into the PMLs; no cell has three."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()
    rs = np.random.RandomState(11)
    g[4:9, 4:10, 4:10] = fd.AbsorbingObject(permittivity=2.0, conductivity=2.0e4, name="P")
    g[7:11, 7:13, 6:12] = fd.AbsorbingObject(permittivity=1.0 + rs.rand(4, 6, 6, 1), conductivity=6.0e3, name="Q")
    g[11:15, 4:10, 4:10] = fd.AnisotropicObject(permittivity=1.0 + rs.rand(4, 6, 6, 3), name="R")
    g[13:18, 7:12, 7:13] = fd.AbsorbingObject(permittivity=1.5, conductivity=1.0e4, name="S")
    g[4:10, 13:18, 4:10] = fd.Object(permittivity=1.0 + rs.rand(6, 5, 6), name="T")
    g[8:13, 14:19, 8:14] = fd.AnisotropicObject(permittivity=1.0 + rs.rand(5, 5, 6, 3), name="U")
    g[11:16, 13:17, 12:16] = fd.AbsorbingObject(permittivity=2.2, conductivity=3.0e4, name="V")
    g[14:18, 14:18, 13:17] = fd.Object(permittivity=3.0, name="W")
    g[16:20, 3:7, 13:17] = fd.AnisotropicObject(permittivity=1.0 + rs.rand(4, 4, 4, 3), name="X")
    g[17:21, 5:9, 15:19] = fd.AnisotropicObject(permittivity=1.0 + rs.rand(4, 4, 4, 3), name="Y")
    g[5, :, :] = fd.PlaneSource(period=18, polarization="y", name="plane")
    g[10, 10, 10] = fd.PointSource(period=12, amplitude=0.8, name="pt")
    g[2:20, 8, 8] = fd.LineDetector(name="line")
    g[11:12, 14:16, 12:13] = fd.BlockDetector(name="block")
    g[17:18, 5:6, 15:16] = fd.BlockDetector(name="block2")
    return g


def overlaps3d_stable(fd, n=(22, 20, 20), t=4):
    """the object pairs of overlaps3d in a run that stays bounded: where two objects cover a cell the reference adds
    both updates, i.e. the cell sees the SUM of their eps^-1 (and the border fix puts the grid's vacuum value into an
    object's last planes), which overlaps3d's permittivities push past the Courant limit (|E| ~ 1e17 after 60 steps:
    fine as a bit-equality check, useless as a relative-error one).  Permittivities >= 2 and a Courant number of
    0.35 keep every cell stable, so rel-L2 means something over the 300 steps of this scene."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9, courant_number=0.35)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, -t:, :] = fd.PML()
    g[:, :, 0:t] = fd.PML()
    g[:, :, -t:] = fd.PML()
    rs = np.random.RandomState(12)
    g[4:9, 4:10, 4:10] = fd.AbsorbingObject(permittivity=2.0, conductivity=2.0e4, name="P")
    g[7:11, 7:13, 6:12] = fd.AbsorbingObject(permittivity=2.0 + rs.rand(4, 6, 6, 1), conductivity=6.0e3, name="Q")
    g[11:15, 4:10, 4:10] = fd.AnisotropicObject(permittivity=2.0 + rs.rand(4, 6, 6, 3), name="R")
    g[13:18, 7:12, 7:13] = fd.AbsorbingObject(permittivity=2.5, conductivity=1.0e4, name="S")
    g[4:10, 13:18, 4:10] = fd.Object(permittivity=2.0 + rs.rand(6, 5, 6), name="T")
    g[8:13, 14:19, 8:14] = fd.AnisotropicObject(permittivity=2.0 + rs.rand(5, 5, 6, 3), name="U")
    g[11:16, 13:17, 12:16] = fd.AbsorbingObject(permittivity=2.2, conductivity=3.0e4, name="V")
    g[14:18, 14:18, 13:17] = fd.Object(permittivity=3.0, name="W")
    g[16:20, 3:7, 13:17] = fd.AnisotropicObject(permittivity=2.0 + rs.rand(4, 4, 4, 3), name="X")
    g[17:21, 5:9, 15:19] = fd.AnisotropicObject(permittivity=2.0 + rs.rand(4, 4, 4, 3), name="Y")
    g[5, :, :] = fd.PlaneSource(period=18, polarization="y", name="plane")
    g[10, 10, 10] = fd.PointSource(period=12, amplitude=0.8, name="pt")
    g[2:20, 8, 8] = fd.LineDetector(name="line")
    g[11:12, 14:16, 12:13] = fd.BlockDetector(name="block")
    g[17:18, 5:6, 15:16] = fd.BlockDetector(name="block2")
    return g


def stacked3d(fd, n=(20, 18, 16), t=3):
    """objects stacked three, four and five deep on the same cells, absorbers among them at every depth, one stack
    reaching into the PMLs and across the middle of the grid (an x-slab boundary when sharded)."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9, courant_number=0.25)
    g[0:t, :, :] = fd.PML()
    g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, :, -t:] = fd.PML()
    rs = np.random.RandomState(21)
    g[3:12, 3:12, 3:12] = fd.Object(permittivity=3.0, name="objA")
    g[5:14, 5:14, 4:11] = fd.AbsorbingObject(permittivity=2.5 + rs.rand(9, 9, 7, 1), conductivity=8.0e3, name="objB")
    g[7:13, 2:10, 5:12] = fd.AnisotropicObject(permittivity=3.0 + rs.rand(6, 8, 7, 3), name="objC")      # third
    g[8:16, 6:12, 6:10] = fd.AbsorbingObject(permittivity=4.0, conductivity=2.0e4, name="objD")          # fourth
    g[9:11, 7:16, 2:14] = fd.Object(permittivity=3.0 + rs.rand(2, 9, 12), name="objE")                   # fifth
    g[0:6, 0:6, 8:16] = fd.Object(permittivity=3.5, name="objF")                                         # in three PMLs
    g[1:5, 1:7, 9:15] = fd.AbsorbingObject(permittivity=3.0, conductivity=1.0e4, name="objG")
    g[2:8, 2:5, 10:16] = fd.AnisotropicObject(permittivity=3.0 + rs.rand(6, 3, 6, 3), name="objH")       # third, in PMLs
    g[4, :, :] = fd.PlaneSource(period=16, polarization="y", name="plane")
    g[10, 8, 8] = fd.PointSource(period=11, amplitude=0.7, name="pt")                                 # on a 5-deep cell
    g[2:18, 8, 7] = fd.LineDetector(name="line")
    g[9:10, 7:9, 7:8] = fd.BlockDetector(name="block")
    return g


def patch_antenna(fd, patch=(20, 14), border=4, steps=240):
    """the reference's probe-fed patch antenna (tests/test_antenna_impedance.py:18-99) at reduced size and with
    one copper object per region (the original registers ONE AbsorbingObject instance three times, which
    fails in the reference itself: the conductivity broadcast of the second shape, fdtd/objects.py:194).
    Ground plane, substrate in four pieces around the feed column, top plane, a copper via that overlaps the
    top plane, a SoftArbitraryPointSource with series impedance and a gaussian pulse, PML on all faces.  The
    original's Z = 50 diverges in the reference itself (the explicit one-step feedback V_out = V_in + Z*I is
    unstable from Z ~ 1 in simulation units: 1e110 after 240 steps); Z = 0.5 is used here."""
    nsub = 3
    spacing = 1.27e-3 / nsub
    n = (patch[0] + 2 * border, patch[1] + 2 * border, nsub + 2 * border + 2)
    g = fd.Grid(shape=n, grid_spacing=spacing, permittivity=1.0, permeability=1.0, courant_number=None)
    pml = fd.PML
    g[0:border, :, :] = pml(); g[-border:, :, :] = pml()
    g[:, 0:border, :] = pml(); g[:, -border:, :] = pml()
    g[:, :, 0:border] = pml(); g[:, :, -border:] = pml()
    sl = slice(border, -border)
    ground, top = border, border + 1 + nsub

    def copper():
        return fd.AbsorbingObject(permittivity=1.0, conductivity=1e8)

    g[sl, sl, ground:ground + 1] = copper()
    g[sl, sl, top:top + 1] = copper()
    px, py = patch[0] // 2 + border, patch[1] // 4 + border
    t = np.arange(steps) * g.time_step
    fwhm = steps * g.time_step / 5.0
    sigma = fwhm / 2.355
    wave = np.exp(-0.5 * ((t - 2.0 * fwhm) / sigma) ** 2)
    saps = getattr(fd, "SoftArbitraryPointSource", None) or fd.sources.SoftArbitraryPointSource
    g[px, py, ground + 1] = saps(waveform_array=wave, impedance=0.5)
    g[px:px + 1, py:py + 1, ground + 2:top + 1] = copper()          # feed via, one cell into the top plane
    eps_r = 2.42
    g[border:px, sl, ground + 1:top] = fd.Object(permittivity=eps_r)
    g[px + 1:-border, sl, ground + 1:top] = fd.Object(permittivity=eps_r)
    g[px:px + 1, border:py, ground + 1:top] = fd.Object(permittivity=eps_r)
    g[px:px + 1, py + 1:-border, ground + 1:top] = fd.Object(permittivity=eps_r)
    g[px + 3, py, ground + 2] = fd.BlockDetector(name="near")
    return g


# name -> (builder, steps)
SCENES = {
    "quickstart2d": (quickstart2d, 300),
    "quickstart2d_full": (quickstart2d, 1000),       # BASELINE configs[0] at its own step count
    "pml3d": (pml3d, 60),
    "objects3d": (objects3d, 60),
    "periodic3d": (periodic3d, 80),
    "vacuum_aniso": (vacuum_aniso, 40),
    "slab2d_xz": (slab2d_xz, 120),
    "c4small": (c4small, 50),
    "feed50": (feed50, 80),
    "overlaps3d": (overlaps3d, 60),
    "overlaps3d_stable": (overlaps3d_stable, 300),
    "stacked3d": (stacked3d, 120),
    "patch_antenna": (patch_antenna, 240),
    "ring3d": (ring3d, 70),
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


def spectra(fd, grid):
    """what a user gets out of FrequencyRoutines (fdtd/fourier.py) for a scene with a SoftArbitraryPointSource
    and a one-cell BlockDetector: the port impedance, the FFT of the port record, and the FFT of a detector's
    Ez time trace with and without zero... edge padding."""
    FR = fd.FrequencyRoutines
    port = next(s for s in grid.sources if hasattr(s, "source_voltage"))
    det = next(d for d in grid.detectors if not hasattr(d, "I"))
    trace = np.array([_np(e)[0, 0, 0, 2] for e in det.E], dtype=np.float64)
    out = {}
    for key, (f, v) in {
        "Z": FR(grid, port).impedance(),
        "Zpad": FR(grid, port).impedance(fft_num_bins_in_window=3 * len(trace)),
        "port": FR(grid, port).FFT(),
        "current": FR(grid, port.current_detector).FFT(),
        "trace": FR(grid, trace).FFT(),
        "tracepad": FR(grid, trace).FFT(fft_num_bins_in_window=4 * len(trace)),
    }.items():
        out[key + "_f"], out[key] = np.asarray(_np(f), dtype=np.float64), np.asarray(_np(v))
    return out


def track_all(grid, steps, bins=(1, 3, 7)):
    """fdtd_b200 only: running device DFT of every detector at FFT bins of a `steps`-long record."""
    freqs = np.asarray(bins, dtype=np.float64) / (steps * grid.time_step)
    for det in grid.detectors:
        det.track_frequencies(freqs)
    return freqs


def dump_tracked(grid):
    out = {}
    for n, det in enumerate(grid.detectors):
        if hasattr(det, "I"):
            out[f"det{n}_SI"] = det.spectrum_I
        else:
            out[f"det{n}_SE"] = det.spectrum_E
            out[f"det{n}_SH"] = det.spectrum_H
    return out


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if np.iscomplexobj(a) or np.iscomplexobj(b):
        a, b = a.astype(np.complex128), b.astype(np.complex128)
    else:
        a, b = a.astype(np.float64), b.astype(np.float64)
    den = np.sqrt((np.abs(b) ** 2).sum())
    num = np.sqrt((np.abs(a - b) ** 2).sum())
    return num / den if den > 0 else num


# ---- user-defined plug-ins (the reference's duck-typed protocol, fdtd/grid.py:279-299, 305-325) -------------------
# Written against that protocol only -- _register_grid + update_phi_E/H, update_E/H, detect_E/H on `grid.E` / `grid.H`
# -- so the same classes run on the unmodified reference, on the oracle and on fdtd_b200.

class UserGainObject:
    """a user object: scales E in its box and feeds a little of the curl back"""

    def __init__(self, factor=0.995, feed=0.01, name=None):
        self.factor, self.feed, self.name = factor, feed, name

    def _register_grid(self, grid, x, y, z):
        self.grid, self.loc = grid, (x, y, z)
        grid.objects.append(self)

    def update_E(self, curl_H):
        E = self.grid.E
        E[self.loc] = E[self.loc] * self.factor + self.feed * curl_H[self.loc]

    def update_H(self, curl_E):
        pass


class UserSquareSource:
    """a user source: a square wave on Ey of one cell, and a kick on Hx every seventh step"""

    def __init__(self, half_period=5, name=None):
        self.half_period, self.name = half_period, name

    def _register_grid(self, grid, x, y, z):
        self.grid, self.cell = grid, (x[0], y[0], z[0])
        grid.sources.append(self)

    def update_E(self):
        q = self.grid.time_steps_passed
        x, y, z = self.cell
        self.grid.E[x, y, z, 1] += 1.0 if (q // self.half_period) % 2 == 0 else -1.0

    def update_H(self):
        if self.grid.time_steps_passed % 7 == 0:
            x, y, z = self.cell
            self.grid.H[x, y, z, 0] += 0.25


class UserProbe:
    """a user detector: one component of one cell per half-step, as python floats"""

    def __init__(self, name=None):
        self.name, self.E, self.H = name, [], []

    def _register_grid(self, grid, x, y, z):
        self.grid, self.cell = grid, (x[0], y[0], z[0])
        grid.detectors.append(self)

    def detect_E(self):
        x, y, z = self.cell
        self.E.append(float(self.grid.E[x, y, z, 1]))

    def detect_H(self):
        x, y, z = self.cell
        self.H.append(float(self.grid.H[x, y, z, 2]))

    def detector_values(self):
        return {"E": self.E, "H": self.H}


class UserWall:
    """a user boundary: a perfect-conductor wall on one x-plane (tangential E forced to zero after every E update)"""

    def __init__(self, name=None):
        self.name = name

    def _register_grid(self, grid, x, y, z):
        self.grid, self.ix = grid, x[0]
        grid.boundaries.append(self)

    def update_phi_E(self):
        pass

    def update_phi_H(self):
        pass

    def update_E(self):
        self.grid.E[self.ix, :, :, 1:] = 0.0

    def update_H(self):
        pass

    # the oracle names the same four hooks differently (oracle/yee_oracle.py, _Boundary)
    def advance_psi_E(self, d):
        pass

    def advance_psi_H(self, d):
        pass

    def apply_E(self):
        self.update_E()

    def apply_H(self):
        pass


def user_plugins(fd, n=(20, 16, 14), t=3):
    """built-in and user-defined plug-ins side by side: PMLs, a built-in object and sources, and a user object (inside
    a PML too), a user source, a user detector and a user boundary."""
    g = fd.Grid(shape=n, grid_spacing=77.5e-9)
    g[0:t, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[5:9, 5:9, 4:8] = fd.Object(permittivity=2.2, name="glass")
    g[1:12, 1:7, 8:13] = UserGainObject(name="gain")
    g[n[0] - 3, :, :] = UserWall(name="wall")
    g[10, 8, 7] = fd.PointSource(period=13, name="pt")
    g[12, 9, 6] = UserSquareSource(name="square")
    g[3:17, 8, 7] = fd.LineDetector(name="line")
    g[11, 9, 6] = UserProbe(name="probe")
    return g
