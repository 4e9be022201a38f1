"""Seeded random scenes: arbitrary extents (odd / even / multiple-of-4 z, flat axes), any subset of PML
faces with different thicknesses, periodic axes, scalar / per-cell / anisotropic grid materials, every object
kind (possibly overlapping PMLs and each other), every source and detector kind, in random registration
order where order matters.  Used against the oracle by the emu (CPU) and CUDA (GPU) parity tests."""
import numpy as np


def random_scene(seed):
    rs = np.random.RandomState(seed)

    def build(fd):
        r = np.random.RandomState(seed)        # same draws for every implementation
        dims = [int(r.randint(6, 19)) for _ in range(3)]
        flat = r.randint(0, 6)
        if flat < 3 and r.rand() < 0.35:
            dims[flat] = 1
        if r.rand() < 0.5:
            dims[2] = int(4 * r.randint(2, 6))     # vector-width friendly z more often
        n = tuple(dims)
        kw = {}
        mat = r.randint(0, 4)
        if mat == 1:
            kw["permittivity"] = 1.0 + r.rand(*n, 3)
        elif mat == 2:
            kw["permittivity"] = float(1.0 + r.rand())
            kw["permeability"] = 1.0 + 0.5 * r.rand(*n)
        elif mat == 3:
            kw["permittivity"] = 1.0 + r.rand(*n, 1)
            kw["permeability"] = float(1.0 + 0.2 * r.rand())
        g = fd.Grid(shape=n, grid_spacing=float(50e-9 * (1 + r.rand())), **kw)

        # boundaries: per axis either nothing, PML low / high (any thickness), or periodic
        boundary_ops = []
        for axis in range(3):
            if n[axis] < 4:
                continue
            mode = r.randint(0, 4)
            if mode == 3:
                boundary_ops.append(("periodic", axis, None))
                continue
            for side in ("low", "high"):
                if mode == 0 or r.rand() < 0.75:
                    t = int(r.randint(1, max(2, n[axis] // 3 + 1)))
                    boundary_ops.append(("pml", axis, (side, t)))
        r.shuffle(boundary_ops)
        for kind, axis, arg in boundary_ops:
            key = [slice(None)] * 3
            if kind == "periodic":
                key[axis] = 0 if r.rand() < 0.5 else -1
                g[tuple(key)] = fd.PeriodicBoundary()
            else:
                side, t = arg
                key[axis] = slice(0, t) if side == "low" else slice(-t, None)
                g[tuple(key)] = fd.PML(a=float(10 ** r.uniform(-9, -6)))

        def box(min_extent=1):
            out = []
            for axis in range(3):
                a = int(r.randint(0, n[axis]))
                b = int(r.randint(a + 1, n[axis] + 1)) if n[axis] > 1 else 1
                if b - a < min_extent and n[axis] >= min_extent:
                    a, b = 0, min_extent
                out.append(slice(a, b))
            return tuple(out)

        # any number of objects of any kinds may share a cell (the reference updates each in registration order)
        for _ in range(r.randint(0, 7)):
            b = box()
            shape = tuple(s.stop - s.start for s in b)
            kind = r.randint(0, 3)
            if kind == 0:
                eps = r.choice([float(1 + 2 * r.rand()), None])
                eps = eps if eps is not None else 1.0 + r.rand(*shape)
                g[b] = fd.Object(permittivity=eps)
            elif kind == 1:
                g[b] = fd.AnisotropicObject(permittivity=1.0 + r.rand(*shape, 3))
            else:
                g[b] = fd.AbsorbingObject(permittivity=float(1 + r.rand()),
                                          conductivity=float(10 ** r.uniform(2, 4.5)))

        def cell():
            return tuple(int(r.randint(0, v)) for v in n)

        for _ in range(r.randint(1, 4)):
            kind = r.randint(0, 3)
            if kind == 0:
                g[cell()] = fd.PointSource(period=int(r.randint(5, 30)), amplitude=float(r.rand() + 0.2),
                                           phase_shift=float(r.rand()), pulse=bool(r.rand() < 0.3),
                                           cycle=int(r.randint(2, 6)), hanning_dt=float(r.uniform(1, 6)))
            elif kind == 1:
                b = box(min_extent=2)
                if max(s.stop - s.start for s in b) >= 2:
                    g[b] = fd.LineSource(period=int(r.randint(5, 30)), amplitude=float(r.rand() + 0.2),
                                         pulse=bool(r.rand() < 0.3))
            else:
                axes = [a for a in range(3) if n[a] > 1]
                if len(axes) >= 2:
                    normal = int(r.choice(range(3)))
                    others = [a for a in range(3) if a != normal]
                    if all(n[a] > 1 for a in others):
                        key = [slice(None)] * 3
                        key[normal] = int(r.randint(0, n[normal]))
                        pol = "xyz"[int(r.choice(others))]
                        g[tuple(key)] = fd.PlaneSource(period=int(r.randint(8, 30)), polarization=pol,
                                                       amplitude=float(r.rand() + 0.5))
        for _ in range(r.randint(1, 3)):
            if r.rand() < 0.5:
                b = box()
                if max(s.stop - s.start for s in b) >= 1:
                    g[b] = fd.LineDetector()
            else:
                c = cell()
                key = tuple(slice(v, min(v + int(r.randint(0, 2)), m - 1)) for v, m in zip(c, n))
                g[key] = fd.BlockDetector()
        if r.rand() < 0.4:
            g[cell()] = fd.CurrentDetector()
        return g

    steps = int(rs.randint(12, 30))
    return build, steps


def random_fused_scene(seed, large=False):
    """Seeded random scenes the single-pass E+H kernel is eligible for: homogeneous background (scalar materials),
    any subset of the six PML faces with their own thicknesses in random registration order, point / line sources on
    E anywhere (interior, slabs, faces), line / block detectors; z extents around the kernel's tile length (31 vectors
    of 4 float32 / 2 float64 cells), so that slabs start at, before and after tile and halo-lane boundaries.
    `large`: extents the library accepts for the kernel on a real GPU (at least 8 x 8 x 32 vectors).
    Returns (build, steps, x_chunk, split): the chunking and whether the step runs as the two launches of a slab."""
    r0 = np.random.RandomState(1000 + seed)
    steps = int(r0.randint(5, 12))
    x_chunk = int(r0.choice([0, 0, 1, 2, 3, 5, 7]))
    split = bool(r0.randint(0, 2))

    def build(fd):
        r = np.random.RandomState(seed)
        nz = int(r.choice([int(r.randint(8, 40)), 62, 64, 66, 124, 126, 128, 132, int(r.randint(120, 150))]))
        nz -= nz % 4                                        # (the fused kernel wants whole vectors along z)
        n = (int(r.randint(4, 24)), int(r.randint(3, 20)), max(8, nz))
        if large:
            n = (int(r.randint(8, 60)), int(r.randint(8, 40)), int(r.choice([128, 132, 136, 160, 244, 248, 252, 256, 300])))
        g = fd.Grid(shape=n, grid_spacing=float(60e-9 * (1 + r.rand())), permittivity=float(1 + r.rand()),
                    permeability=float(1 + 0.3 * r.rand()))
        faces = [(a, side) for a in range(3) for side in (0, 1) if r.rand() < 0.7]
        r.shuffle(faces)
        for a, side in faces:
            t = int(r.randint(1, max(2, min(n[a] // 2 - 1, 18 if a == 2 else 6))))
            if t >= n[a] // 2:
                continue
            key = [slice(None)] * 3
            key[a] = slice(0, t) if side == 0 else slice(-t, None)
            g[tuple(key)] = fd.PML()
        for k in range(int(r.randint(1, 5))):
            p = [int(r.randint(0, n[a])) for a in range(3)]
            g[p[0], p[1], p[2]] = fd.PointSource(period=int(r.randint(5, 30)), amplitude=float(0.2 + r.rand()),
                                                 name=f"p{k}")
        if n[0] > 6 and r.rand() < 0.6:
            y, z = int(r.randint(0, n[1])), int(r.randint(0, n[2]))
            g[1:n[0] - 1, y, z] = fd.LineSource(period=int(r.randint(7, 25)), name="line")
        y, z = int(r.randint(0, n[1])), int(r.randint(0, n[2]))
        g[0:n[0], y, z] = fd.LineDetector(name="across")
        x, y = int(r.randint(0, n[0] - 1)), int(r.randint(0, n[1] - 2))      # (BlockDetector ranges include `stop`)
        g[x:x + 1, y:y + 1, n[2] - 3:n[2] - 2] = fd.BlockDetector(name="block")
        return g

    return build, steps, x_chunk, split
