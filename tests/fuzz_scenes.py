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
