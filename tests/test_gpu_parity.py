"""Parity of the CUDA path (through the C ABI, on a real GPU) against

  * tests/golden: outputs of the unmodified reference (float64 numpy and true-float32 torch),
  * the oracle on seeded scenes at sizes it finishes in seconds,
  * size-independent properties at BASELINE.json's full sizes (exact linearity under a power-of-two
    amplitude, x-chunk invariance, run(n) == n*step(), a quiescent grid stays exactly zero).

Tolerances are north_star's: rel-L2 <= 1e-12 (float64), <= 1e-5 (float32) on final E, H and on every
detector trace; bit-equality is asserted in addition where the reference's operation order is kept.
"""
import glob
import os

import numpy as np
import pytest
import torch

import scenes
from oracle import yee_oracle as yo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = {"float64": 1e-12, "float32": 1e-5}


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def cuda(dtype):
    import fdtd_b200 as fd
    fd.set_backend("cuda." + dtype)
    from fdtd_b200 import _capi
    assert fd.backend.lib is _capi.load()          # the nvcc-built library, nothing else
    return fd


def run_scene(fd, build, steps, **kw):
    g = build(fd, **kw)
    g.run(steps, progress_bar=False)
    return scenes.dump(g)


def run_oracle(build, steps, dtype="float64", **kw):
    yo.set_backend("numpy" if dtype == "float64" else "torch", dtype)
    try:
        g = build(yo, **kw)
        g.run(steps)
        return scenes.dump(g)
    finally:
        yo.set_backend("numpy", "float64")


def compare(got, want, tol, bitwise=False):
    assert set(got) == set(want)
    worst = 0.0
    for k in want:
        assert got[k].shape == want[k].shape, k
        err = scenes.rel_l2(got[k], want[k])
        worst = max(worst, err)
        assert err <= tol, f"{k}: rel-L2 {err:.3e} > {tol}"
        if bitwise and not k.startswith("src"):      # recorded source voltages: host float64 in the reference
            assert np.array_equal(got[k], want[k]), f"{k}: not bit-identical (rel-L2 {err:.3e})"
    return worst


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "*_f*.npz"))),
                         ids=lambda p: os.path.basename(p)[:-4])
def test_cuda_vs_reference_golden(path):
    scene, prec = os.path.basename(path)[:-4].rsplit("_", 1)
    gold = dict(np.load(path))
    steps = int(gold.pop("steps"))
    dtype = "float64" if prec == "f64" else "float32"
    got = run_scene(cuda(dtype), scenes.SCENES[scene][0], steps)
    # bit-identical to the unmodified reference: same operation order, -fmad=false, host tables (exp of a few dozen
    # numbers) from the same numpy / torch the reference backend uses
    # (bitwise: every array but the recorded source voltages, which the reference keeps as host float64)
    worst = compare(got, gold, TOL[dtype], bitwise=True)
    print(f"{scene}/{dtype}: bit-identical to the reference; worst rel-L2 incl. source records {worst:.3e}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n,t", [((13, 11, 12), 4), ((12, 9, 10), 3), ((11, 12, 7), 3), ((9, 40, 36), 5),
                                 ((10, 5, 132), 2), ((40, 72, 260), 6)])
def test_cuda_vs_oracle_shapes(n, t, dtype):
    got = run_scene(cuda(dtype), scenes.pml3d, 25, n=n, t=t)
    want = run_oracle(scenes.pml3d, 25, dtype, n=n, t=t)
    compare(got, want, TOL[dtype], bitwise=True)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_cuda_vs_oracle_objects_medium(dtype):
    """config-2/3 structure at 64x56x48: heterogeneous eps/mu, all object kinds, all source kinds."""
    n = (64, 56, 48)

    def build(fd):
        rs = np.random.RandomState(11)
        eps = 1.0 + rs.rand(*n, 3)
        g = fd.Grid(shape=n, grid_spacing=77.5e-9, permittivity=eps)
        for key in ((slice(0, 8), slice(None), slice(None)), (slice(-8, None), slice(None), slice(None)),
                    (slice(None), slice(0, 8), slice(None)), (slice(None), slice(-8, None), slice(None)),
                    (slice(None), slice(None), slice(0, 8)), (slice(None), slice(None), slice(-8, None))):
            g[key] = fd.PML()
        g[12, :, :] = fd.PlaneSource(period=20, polarization="y")
        g[20:30, 10:40, 12:40] = fd.AbsorbingObject(permittivity=2.5, conductivity=1.5e4)
        g[34:50, 16:48, 8:40] = fd.Object(permittivity=1.0 + rs.rand(16, 32, 32))
        g[52:60, 2:30, 30:46] = fd.Object(permittivity=4.0)       # overlaps three PMLs
        g[16:40, 20:44, 24] = fd.LineSource(period=15, amplitude=2.0)
        g[8:56, 28, 24] = fd.LineDetector()
        g[40:42, 30:31, 20:22] = fd.BlockDetector()
        return g

    got = run_scene(cuda(dtype), build, 40)
    want = run_oracle(build, 40, dtype)
    compare(got, want, TOL[dtype], bitwise=True)


def test_pokes_and_rebake_on_gpu():
    def drive(fd):
        g = scenes.pml3d(fd, n=(20, 18, 16), t=4)
        for n in range(20):
            g.update_E()
            g.E[10, 9, 8, 2] = 0.25 * n
            g.update_H()
            g.time_steps_passed += 1
        g[3:6, 2:5, 2:6] = fd.Object(permittivity=2.2)
        g.run(15, progress_bar=False)
        g.reset()
        g.run(5, progress_bar=False)
        return scenes.dump(g)
    got = drive(cuda("float64"))
    yo.set_backend("numpy", "float64")
    want = drive(yo)
    compare(got, want, 1e-12, bitwise=True)


# ---- full-size properties (BASELINE.json configs 1-3 sizes) -------------------------------------

def _c4(fd, n, amplitude=1.0, with_source=True):
    if not isinstance(n, int):
        nx, ny, nz = n
        g = fd.Grid(shape=(nx, ny, nz), grid_spacing=77.5e-9)
        for key in ((slice(0, 10),), (slice(-10, None),), (slice(None), slice(0, 10)), (slice(None), slice(-10, None)),
                    (slice(None), slice(None), slice(0, 10)), (slice(None), slice(None), slice(-10, None))):
            g[key] = fd.PML()
        g[nx // 2, ny // 2, nz // 2] = fd.PointSource(period=20, amplitude=amplitude)
        g[nx // 2 + 4, ny // 2, 12:nz - 12] = fd.LineDetector()
        return g
    g = fd.Grid(shape=(n, n, n), grid_spacing=77.5e-9)
    g[0:10, :, :] = fd.PML()
    g[-10:, :, :] = fd.PML()
    g[:, 0:10, :] = fd.PML()
    g[:, -10:, :] = fd.PML()
    g[:, :, 0:10] = fd.PML()
    g[:, :, -10:] = fd.PML()
    if with_source:
        g[n // 2, n // 2, n // 2] = fd.PointSource(period=20, amplitude=amplitude)
    g[n // 2 + 4, n // 2, 12:n - 12] = fd.LineDetector()
    return g


@pytest.mark.parametrize("n,dtype,steps", [(256, "float64", 60), (512, "float32", 40)])
def test_full_size_linearity_is_exact(n, dtype, steps):
    """doubling the source amplitude doubles every field value and detector sample exactly
    (scaling by a power of two commutes with every rounding of the update)."""
    fd = cuda(dtype)
    a = _c4(fd, n, 1.0)
    a.run(steps, progress_bar=False)
    Ea = a.E.clone()
    da = np.stack(a.detectors[0].E)
    del a
    b = _c4(fd, n, 2.0)
    b.run(steps, progress_bar=False)
    assert float(Ea.abs().max()) > 0
    assert torch.equal(b.E, 2 * Ea)
    assert np.array_equal(np.stack(b.detectors[0].E), 2 * da)


def test_full_size_chunk_invariance_and_stepping():
    fd = cuda("float32")
    ref = _c4(fd, 256)
    ref.run(30, progress_bar=False)
    other = _c4(fd, 256)
    other._x_chunk = 7
    for _ in range(30):
        other.step()
    assert torch.equal(ref.E, other.E) and torch.equal(ref.H, other.H)
    assert np.array_equal(np.stack(ref.detectors[0].H), np.stack(other.detectors[0].H))


def test_quiescent_grid_stays_zero():
    fd = cuda("float32")
    g = _c4(fd, 256, with_source=False)
    g.run(10, progress_bar=False)
    assert float(g.E.abs().max()) == 0.0 and float(g.H.abs().max()) == 0.0


def test_c4_256_vs_oracle_prefix():
    """the bench workload at 96^3 against the oracle (the oracle needs seconds per step beyond that)."""
    def build(fd):
        return _c4(fd, 96)
    got = run_scene(cuda("float32"), build, 30)
    want = run_oracle(build, 30, "float32")
    compare(got, want, 1e-5, bitwise=True)


def test_graph_replay_equals_plain_launches():
    """run() on a small grid replays CUDA graphs of 32-step chunks (waveform index and ring slot come
    from device scalars); results must equal plain per-step launches bit for bit."""
    fd = cuda("float64")
    outs = []
    for graphs in (True, False):
        g = scenes.objects3d(fd)
        g._use_graphs = graphs
        g.run(70, progress_bar=False)
        g.run(45, progress_bar=False)
        assert g._engine.desc.use_graphs == (1 if graphs else 0)
        outs.append(scenes.dump(g))
    compare(outs[0], outs[1], 0.0, bitwise=True)
    want = run_oracle(scenes.objects3d, 115)
    compare(outs[0], want, 1e-12)


def test_many_sources_fall_back_to_unfused_kernels():
    """more than FDTD_FUSED_MAX sources: the separate source / detector kernels run instead."""
    def build(fd):
        g = scenes.pml3d(fd, n=(20, 18, 16), t=4)
        for n in range(8):
            g[5 + n, 6, 7] = fd.PointSource(period=9 + n, amplitude=0.5 + 0.1 * n)
        return g
    fd = cuda("float64")
    g = build(fd)
    g.run(40, progress_bar=False)
    import ctypes
    assert g._engine.lib.fdtd_post_is_fused(ctypes.byref(g._engine.desc)) == 0
    yo.set_backend("numpy", "float64")
    o = build(yo)
    o.run(40)
    compare(scenes.dump(g), scenes.dump(o), 1e-12, bitwise=True)


def test_float32_error_growth_against_float64_reference():
    """north_star: <= 1e-5 in float32 against the reference's own backends -- which compute in float64 whatever
    their name says.  40^3, six PMLs, continuous point source, to 2000 steps (BASELINE config 1's
    count).  "cuda.float32" is bit-identical to the reference's TRUE float32 run and inherits its drift (the rounded
    coefficients shift the phase velocity: ~1e-5 per 1000 steps); "cuda.float32x" (float32 state, float64 arithmetic
    and coefficients) holds the bar at every step count.  The curve is printed."""
    def build(fd):
        return scenes.pml3d(fd, n=(40, 40, 40), t=8)
    yo.set_backend("torch", "float64")
    try:
        ref = build(yo)
        grids = {"float32": build(cuda("float32")), "float32x": build(cuda("float32x"))}
        done, curve = 0, {}
        for steps in (100, 200, 500, 1000, 2000):
            ref.run(steps - done)
            for mode, g in grids.items():
                cuda(mode)
                g.run(steps - done, progress_bar=False)
                e = scenes.rel_l2(g.E.double().cpu().numpy(), ref.E.numpy())
                h = scenes.rel_l2(g.H.double().cpu().numpy(), ref.H.numpy())
                curve[(mode, steps)] = (e, h)
            done = steps
        want = np.stack([np.asarray(v) for v in ref.detectors[0].E])
        det = {m: scenes.rel_l2(np.stack(g.detectors[0].E).astype(np.float64), want) for m, g in grids.items()}
    finally:
        yo.set_backend("numpy", "float64")
    print("rel-L2 vs the float64 reference, (E, H) by step count:")
    for mode in grids:
        print(f"  cuda.{mode}: " + ", ".join(f"{n}: ({curve[(mode, n)][0]:.2e}, {curve[(mode, n)][1]:.2e})"
                                             for n in (100, 200, 500, 1000, 2000)) + f"; detector {det[mode]:.2e}")
    for steps in (100, 200, 500):                    # the step counts of BASELINE's float32 configs (2-4: <= 500)
        assert max(curve[("float32", steps)]) <= 1e-5, curve
    for steps in (100, 200, 500, 1000, 2000):
        assert max(curve[("float32x", steps)]) <= 1e-5, curve
    assert det["float32x"] <= 1e-5


@pytest.mark.parametrize("scene", ["pml3d", "objects3d", "periodic3d", "overlaps3d_stable", "ring3d", "quickstart2d",
                                   "patch_antenna", "c4small"])
def test_float32x_vs_float64_reference(scene):
    """cuda.float32x against the reference's float64 outputs (tests/golden): <= 1e-5 on E, H and every trace."""
    gold = dict(np.load(os.path.join(GOLD, f"{scene}_f64.npz")))
    steps = int(gold.pop("steps"))
    got = run_scene(cuda("float32x"), scenes.SCENES[scene][0], steps)
    assert set(got) == set(gold)
    for k in gold:
        if k.startswith("src"):
            continue
        assert got[k].dtype == np.float32
        err = scenes.rel_l2(got[k].astype(np.float64), gold[k])
        assert err <= 1e-5, f"{scene} {k}: rel-L2 {err:.3e}"


def test_config1_at_its_full_step_count():
    """BASELINE configs[0] exactly as stated: 161x97x1 quick-start grid, 1000 steps, float64 -- against the
    unmodified reference (tests/golden/quickstart2d_full_f64.npz), bit for bit, through the graph-replay path."""
    gold = dict(np.load(os.path.join(GOLD, "quickstart2d_full_f64.npz")))
    steps = int(gold.pop("steps"))
    assert steps == 1000
    got = run_scene(cuda("float64"), scenes.quickstart2d, steps)
    compare(got, gold, 1e-12, bitwise=True)


def test_config2_shape_at_its_full_step_count():
    """BASELINE configs[1]'s structure (float64, six 10-cell PMLs, PointSource + BlockDetector) at 64^3 for its full
    2000 steps against the oracle (torch-CPU float64, all host threads): rel-L2 <= 1e-12 on E, H and the traces."""
    def build(fd):
        g = fd.Grid(shape=(64, 64, 64), grid_spacing=77.5e-9)
        for key in ((slice(0, 10),), (slice(-10, None),), (slice(None), slice(0, 10)), (slice(None), slice(-10, None)),
                    (slice(None), slice(None), slice(0, 10)), (slice(None), slice(None), slice(-10, None))):
            g[key] = fd.PML()
        g[32, 32, 32] = fd.PointSource(period=20)
        g[36:38, 32:34, 32:34] = fd.BlockDetector()
        return g
    got = run_scene(cuda("float64"), build, 2000)
    yo.set_backend("torch", "float64")
    try:
        o = build(yo)
        o.run(2000)
        want = scenes.dump(o)
    finally:
        yo.set_backend("numpy", "float64")
    worst = compare(got, want, 1e-12)
    print(f"config-2 shape, 64^3, 2000 steps: worst rel-L2 {worst:.3e}")


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("seed", range(100, 116))
def test_random_scene_on_gpu(seed, dtype):
    """seeded random registrations (tests/fuzz_scenes.py): CUDA vs oracle, bit for bit."""
    from fuzz_scenes import random_scene
    build, steps = random_scene(seed)
    g = build(cuda(dtype))
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
    compare(got, want, TOL[dtype], bitwise=True)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("seed", range(300, 312))
def test_random_fused_scene_on_gpu(seed, dtype, monkeypatch):
    """the single-pass E+H kernel against the two half-steps on random eligible scenes (tests/fuzz_scenes.py: any
    subset of PML faces and thicknesses in any order, sources and detectors anywhere, z extents around the tile length,
    random x-chunking, the step as one launch or as the two launches of a slab): bit-identical on the device."""
    from fuzz_scenes import random_fused_scene
    build, steps, x_chunk, split = random_fused_scene(seed, large=True)
    fd = cuda(dtype)
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


@pytest.mark.parametrize("dtype,n,t", [("float32", (72, 64, 192), 6), ("float64", (40, 52, 100), 5),
                                       ("float32", (33, 41, 148), 3)])
def test_temporally_fused_steps_equal_two_half_steps(dtype, n, t):
    """run() with the single-pass E+H kernel (grid._fuse_eh = 1: one launch per step over the whole grid, ping-pong
    field and psi_E buffers, TMA-staged inputs, shared-memory / shuffle exchange of E_new) must reproduce the two-half-
    step path bit for bit, for even and odd step counts, sources in the interior and in the slabs, detectors
    everywhere."""
    fd = cuda(dtype)

    def build():
        g = fd.Grid(shape=n, grid_spacing=77.5e-9, permittivity=1.3, permeability=1.1)
        g[0:t, :, :] = fd.PML()
        g[-t:, :, :] = fd.PML()
        g[:, 0:t, :] = fd.PML()
        g[:, -t:, :] = fd.PML()
        g[:, :, 0:t + 1] = fd.PML()
        g[n[0] // 2, n[1] // 2, n[2] // 2] = fd.PointSource(period=17, name="centre")
        g[2, n[1] // 2, 3] = fd.PointSource(period=11, amplitude=0.4, name="in_pml")
        g[t + 2:n[0] - t - 2, t + 3:n[1] - t - 3, n[2] // 3] = fd.LineSource(period=23, name="line")
        g[1:n[0] - 1, n[1] // 2 + 1, n[2] // 2 + 2] = fd.LineDetector(name="across")
        g[n[0] // 2:n[0] // 2 + 1, 1:3, n[2] - 3:n[2] - 2] = fd.BlockDetector(name="corner")
        return g

    outs = []
    for fuse in (0, 1):                         # two half-steps / single-pass kernel
        g = build()
        g._fuse_eh = fuse
        g.run(31, progress_bar=False)
        g.step()
        g.run(10, progress_bar=False)
        assert bool(g._engine.lib.fdtd_fuse_eh_active(g._engine.desc)) == bool(fuse)
        outs.append(scenes.dump(g))
    assert float(np.abs(outs[0]["E"]).max()) > 0
    compare(outs[1], outs[0], 0.0, bitwise=True)


@pytest.mark.gpu
def test_frequency_routines_vs_reference():
    """FrequencyRoutines (fdtd/fourier.py) on the reduced patch-antenna run, transforms on the device (cuFFT),
    against the reference's own outputs (tests/golden/spectra_patch_antenna.npz)."""
    gold = dict(np.load(os.path.join(GOLD, "spectra_patch_antenna.npz")))
    steps = int(gold.pop("steps"))
    fd = cuda("float64")
    g = scenes.patch_antenna(fd)
    g.run(steps, progress_bar=False)
    fd.FrequencyRoutines.verbose = False
    got = scenes.spectra(fd, g)
    assert set(got) == set(gold)
    for k in gold:
        assert got[k].shape == gold[k].shape, k
        if k.endswith("_f"):
            assert np.array_equal(got[k], gold[k]), k
        else:
            assert scenes.rel_l2(got[k], gold[k]) <= 1e-10, f"{k}: {scenes.rel_l2(got[k], gold[k]):.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_running_dft_equals_fft_of_the_record(dtype, monkeypatch):
    """detector.track_frequencies (fdtd_dft_accumulate on the device, several ring flushes) against numpy's FFT
    of the same record and of the reference's golden trace."""
    import fdtd_b200.engine as engine
    monkeypatch.setattr(engine, "RING_BYTES", 1)            # ring capacity 16
    gold = np.load(os.path.join(GOLD, f"pml3d_{'f64' if dtype == 'float64' else 'f32'}.npz"))
    steps = int(gold["steps"])
    fd = cuda(dtype)
    g = scenes.pml3d(fd)
    bins = (1, 3, 7)
    scenes.track_all(g, steps, bins)
    g.run(steps, progress_bar=False)
    got = scenes.dump_tracked(g)
    for n, det in enumerate(g.detectors):
        for f in "EH":
            ours = np.fft.fft(np.asarray(getattr(det, f), dtype=np.float64), axis=0)[list(bins)]
            ref = np.fft.fft(gold[f"det{n}_{f}"].astype(np.float64), axis=0)[list(bins)]
            assert scenes.rel_l2(got[f"det{n}_S{f}"], ours) <= 1e-12
            assert scenes.rel_l2(got[f"det{n}_S{f}"], ref) <= (1e-12 if dtype == "float64" else 1e-6)


def test_fused_steps_are_the_default_on_large_homogeneous_grids():
    """automatic mode (grid._fuse_eh = 2, the default): large grids run pairs of single-pass steps when the second buffers fit; the result equals the two-half-step path bit for bit at that
    size (float32)."""
    fd = cuda("float32")
    n = 768                              # 4.5e8 cells: 21 GiB of fields + second buffers
    if torch.cuda.mem_get_info()[0] < 60 << 30:
        pytest.skip("not enough free device memory")
    outs = []
    for fuse in (2, 0):
        g = _c4(fd, n)
        g._fuse_eh = fuse
        g.run(11, progress_bar=False)
        assert g._engine.lib.fdtd_fuse_eh_active(g._engine.desc) == (1 if fuse else 0)
        outs.append((g.E.clone(), g.H.clone(), np.stack(g.detectors[0].E)))
        del g
    assert float(outs[0][0].abs().max()) > 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][2], outs[1][2])


@pytest.mark.parametrize("dtype,shape,fused", [("float32", (64, 1024, 1024), True), ("float32", (64, 576, 576), True),
                                               ("float32", (64, 512, 512), False),       # plane below 1.07 MiB
                                               ("float32", (64, 1024, 520), False),      # 4.19 z tiles: fill 0.84
                                               ("float32", (48, 1024, 1024), False),     # march too short
                                               ("float64", (64, 384, 384), True), ("float64", (64, 256, 256), False)])
def test_automatic_mode_of_the_fused_steps_by_grid_shape(dtype, shape, fused):
    """grid._fuse_eh = 2 (the default): the host layer (which allocates the second buffers) and the library (which
    launches) apply the same test -- plane size, z-tile fill, march length (profiles/r2_s14/fused_sizes.log) -- and the
    fused steps give the two half-steps' result bit for bit wherever they are chosen."""
    fd = cuda(dtype)
    outs = []
    for fuse in (2, 0):
        g = _c4(fd, shape)
        g._fuse_eh = fuse
        g.run(7, progress_bar=False)
        active = g._engine.lib.fdtd_fuse_eh_active(g._engine.desc) == 1
        assert active == (fused and fuse == 2)
        assert (g._E2 is not None) == active          # no second buffers where they are not used
        outs.append((g.E.clone(), g.H.clone()))
        del g
    assert float(outs[0][0].abs().max()) > 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_energy_slice_and_visualize_on_the_device(monkeypatch):
    """SURVEY 8f rank 3 on the real device: `energy_slice` squares and sums ONE plane of the SoA storage on the GPU
    (the reference materialises E^2 + H^2 of the whole grid, fdtd/visualization.py:96-104) and `Grid.visualize`
    draws it (matplotlib is a recording stand-in: this image has none)."""
    import sys
    import types
    from test_visualization import _Recorder
    from fdtd_b200.visualization import energy_slice
    fd = cuda("float32")
    g = scenes.objects3d(fd)
    g.run(25, progress_bar=False)
    E, H = g.E.double().cpu().numpy(), g.H.double().cpu().numpy()
    energy = (E.astype(np.float32) ** 2 + H.astype(np.float32) ** 2)
    for kw, want in (({"x": 7}, energy[7].sum(-1)), ({"y": 9}, energy[:, 9].sum(-1).T), ({"z": -3}, energy[:, :, -3].sum(-1))):
        got = energy_slice(g, **kw)
        assert got.shape == want.shape and got.dtype == np.float32
        assert scenes.rel_l2(got, want) <= 1e-6, kw                  # (the sum over three components may associate differently)
    assert energy_slice(g, z=4).max() > 0
    log = []
    plt, ptc, colors = (_Recorder(n, log) for n in ("matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors"))
    root = types.ModuleType("matplotlib")
    root.pyplot, root.patches, root.colors = plt, ptc, colors
    for name, mod in (("matplotlib", root), ("matplotlib.pyplot", plt), ("matplotlib.patches", ptc),
                      ("matplotlib.colors", colors)):
        monkeypatch.setitem(sys.modules, name, mod)
    fig = g.visualize(z=6)
    calls = [c[0] for c in log]
    assert fig is not None and "imshow" in calls
    shown = next(c for c in log if c[0] == "imshow")[1][0]
    assert shown.shape == (g.Nx, g.Ny) and float(shown.max()) > 0
