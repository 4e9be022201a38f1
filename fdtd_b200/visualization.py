"""Plotting helpers of the reference (fdtd/visualization.py:28-480): `Grid.visualize`, `dB_map_2D`,
`plot_detection`.

Each function is split into a numerical part that needs no plotting library -- and runs where the data is --
and a thin matplotlib part (imported lazily; this module imports without matplotlib):

  * `energy_slice(grid, x|y|z)`: E^2 + H^2 summed over components for ONE plane, computed on the device from
    the strided SoA storage and copied to the host as a 2-D array.  The reference squares and sums the whole
    grid first (fdtd/visualization.py:123); at 1024^3 that is 50 GB of temporaries for a 4 MB picture.  On an
    x-sharded grid only the plane is gathered (x-projection: the owning rank broadcasts it).
  * `scene_outline(grid, x|y|z)`: the sources, detectors, boundaries and objects as drawing primitives in
    the reference's plot coordinates.
  * `peak_to_peak_dB(block_det, axis)`: the decibel map of `dB_map_2D` (fdtd/visualization.py:370-387).
  * `envelope_arrivals(detector_dict, ...)`: Hilbert envelopes and arrival steps of `plot_detection`
    (fdtd/visualization.py:410-440).
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from .backend import backend as bd


def _pyplot():
    try:
        import matplotlib.pyplot as plt
        import matplotlib.patches as ptc
        from matplotlib.colors import LogNorm
    except ImportError as exc:                         # pragma: no cover - depends on the environment
        raise ImportError("fdtd_b200.visualization draws with matplotlib, which is not installed; the numerical "
                          "parts (energy_slice, scene_outline, peak_to_peak_dB, envelope_arrivals) work "
                          "without it") from exc
    return plt, ptc, LogNorm


def _projection(x, y, z):
    """validate the plane selection like the reference (fdtd/visualization.py:88-113) -> (axis, index)."""
    given = [(n, v) for n, v in (("x", x), ("y", y), ("z", z)) if v is not None]
    if not given:
        raise ValueError("at least one projection plane (x, y or z) should be supplied to visualize the grid!")
    name, value = given[0]
    if not isinstance(value, int):
        raise ValueError(f"the `{name}`-location supplied should be a single integer")
    if len(given) > 1:
        others = {"x": ("y", "z"), "y": ("z", "x"), "z": ("x", "y")}[name]
        raise ValueError(f"if an `{name}`-location is supplied, one should not supply a `{others[0]}` or a "
                         f"`{others[1]}`-location!")
    return "xyz".index(name), value


# (row axis, column axis) of the picture for a plane normal to `axis`: x -> (y, z); y -> (z, x); z -> (x, y)
_ROWS_COLS = {0: (1, 2), 1: (2, 0), 2: (0, 1)}


def energy_slice(grid, x=None, y=None, z=None):
    """sum over components of E^2 + H^2 on one plane, as the 2-D host array the reference shows
    (fdtd/visualization.py:123-146: [y, z] for an x-plane, [z, x] for a y-plane, [x, y] for a z-plane)."""
    axis, index = _projection(x, y, z)
    n = (grid.Nx, grid.Ny, grid.Nz)
    rows, cols = _ROWS_COLS[axis]
    if not (n[rows] > 1 and n[cols] > 1):
        raise AssertionError("the projection plane must span more than one cell in both directions")
    if index < 0:
        index += n[axis]
    if not 0 <= index < n[axis]:
        raise IndexError(f"plane index {index} outside the grid")
    if grid._engine is not None:
        grid._engine.quiesce()
    part = grid._part
    E, H = grid._E[:, 1:-1], grid._H[:, 1:-1]            # (3, nx_local, Ny, Nz) SoA storage
    if axis == 0:
        owner = part.owner(index) if part.sharded else 0
        if not part.sharded or part.rank == owner:
            i = index - part.x0
            plane = (E[:, i] ** 2 + H[:, i] ** 2).sum(0)                     # (Ny, Nz)
        else:
            plane = torch.empty((grid.Ny, grid.Nz), dtype=grid._sdtype, device=E.device)
        if part.sharded:
            plane = plane.contiguous()
            dist.broadcast(plane, src=owner)
    else:
        sel = (slice(None), slice(None), index) if axis == 1 else (slice(None), slice(None), slice(None), index)
        local = (E[sel] ** 2 + H[sel] ** 2).sum(0)                           # (nx_local, Nz) or (nx_local, Ny)
        if part.sharded:
            from .sharding import all_gather_slabs
            local = all_gather_slabs(part, local.contiguous(), 0)
        plane = local.T if axis == 1 else local                              # y-plane is shown as [z, x]
    return plane.to("cpu").numpy()


def _span(v):
    """first and last index of a point list / the bounds of a slice / a single index."""
    if isinstance(v, slice):
        return v.start, v.stop
    if isinstance(v, (list, tuple, np.ndarray)):
        return int(v[0]), int(v[-1])
    return int(v), int(v)


def scene_outline(grid, x=None, y=None, z=None):
    """what `visualize` draws on top of the energy picture, as primitives in plot coordinates (horizontal =
    column axis, vertical = row axis of `energy_slice`):
        {"kind": "line" | "marker" | "rect", "role": "source" | "detector" | "periodic" | "pml" | "object", ...}
    lines carry "h", "v" coordinate lists, markers "h", "v" and the (row, col) "cell" whose energy is blanked,
    rectangles "xy", "width", "height" (fdtd/visualization.py:148-310)."""
    from .boundaries import PML, PeriodicBoundary
    from .sources import LineSource, PlaneSource, PointSource, SoftArbitraryPointSource
    axis, _ = _projection(x, y, z)
    rows, cols = _ROWS_COLS[axis]
    n = (grid.Nx, grid.Ny, grid.Nz)
    R, Cn = n[rows], n[cols]
    out = []

    def coords(thing):
        v = (thing.x, thing.y, thing.z)
        return v[rows], v[cols]

    for src in grid.sources:
        r, c = coords(src)
        if isinstance(src, LineSource):
            (r0, r1), (c0, c1) = _span(r), _span(c)
            out.append({"kind": "line", "role": "source", "h": [c0, c1], "v": [r0, r1]})
        elif isinstance(src, (PointSource, SoftArbitraryPointSource)):
            out.append({"kind": "marker", "role": "source", "h": c - 0.5, "v": r - 0.5, "cell": (r, c)})
        elif isinstance(src, PlaneSource):
            # a one-cell extent (the plane's normal, or a flat axis) is drawn with zero size
            (r0, r1), (c0, c1) = [(s.start, s.stop if s.stop > s.start + 1 else s.start) for s in (r, c)]
            out.append({"kind": "rect", "role": "source", "xy": (c0 - 0.5, r0 - 0.5), "width": c1 - c0,
                        "height": r1 - r0})
    for det in grid.detectors:
        r, c = coords(det)
        (r0, r1), (c0, c1) = _span(r), _span(c)
        if type(det).__name__ == "BlockDetector":
            out.append({"kind": "line", "role": "detector", "h": [c0, c1, c1, c0, c0], "v": [r0, r0, r1, r1, r0]})
        else:
            out.append({"kind": "line", "role": "detector", "h": [c0, c1], "v": [r0, r1]})
    nan = float("nan")
    for b in grid.boundaries:
        if isinstance(b, PeriodicBoundary):
            if b.axis == rows:       # the two horizontal edges
                out.append({"kind": "line", "role": "periodic", "h": [-0.5, Cn - 0.5, nan, -0.5, Cn - 0.5],
                            "v": [-0.5, -0.5, nan, R - 0.5, R - 0.5]})
            elif b.axis == cols:     # the two vertical edges
                out.append({"kind": "line", "role": "periodic", "h": [-0.5, -0.5, nan, Cn - 0.5, Cn - 0.5],
                            "v": [-0.5, R - 0.5, nan, -0.5, R - 0.5]})
        elif isinstance(b, PML):
            t, low = b.thickness, b.side == "low"
            if b.axis == cols:
                out.append({"kind": "rect", "role": "pml", "xy": (-0.5 if low else Cn - 0.5 - t, -0.5),
                            "width": t, "height": R})
            elif b.axis == rows:
                out.append({"kind": "rect", "role": "pml", "xy": (-0.5, -0.5 if low else R - t - 0.5),
                            "width": Cn, "height": t})
    for obj in grid.objects:
        r, c = coords(obj)
        out.append({"kind": "rect", "role": "object", "xy": (min(c.start, c.stop) - 0.5, min(r.start, r.stop) - 0.5),
                    "width": abs(c.stop - c.start), "height": abs(r.stop - r.start)})
    return out


def visualize(grid, x=None, y=None, z=None, cmap="Blues", pbcolor="C3", pmlcolor=(0, 0, 0, 0.1),
              objcolor=(1, 0, 0, 0.1), srccolor="C0", detcolor="C2", norm="linear", animate=False, index=None,
              save=False, folder=None, show=False, style=None):
    """Show one plane of the grid: the field energy with sources, detectors, boundaries and objects drawn over
    it.  Arguments as in the reference (fdtd/visualization.py:28-66); returns the matplotlib figure."""
    if norm not in ("linear", "lin", "log"):
        raise ValueError("Color map normalization should be 'linear' or 'log'.")
    axis, _ = _projection(x, y, z)
    plt, ptc, LogNorm = _pyplot()
    if style is not None:
        plt.style.use(style)
    if animate:
        plt.pause(0.02)
        plt.clf()
        plt.ion()
    color = {"source": srccolor, "detector": detcolor, "periodic": pbcolor, "pml": pmlcolor, "object": objcolor}
    for label, role, lw in (("Objects", "object", 7), ("PML", "pml", 7), ("Periodic Boundaries", "periodic", 3),
                            ("Sources", "source", 3), ("Detectors", "detector", 3)):
        plt.plot([], lw=lw, color=color[role], label=label)
    energy = energy_slice(grid, x, y, z)
    for item in scene_outline(grid, x, y, z):
        c = color[item["role"]]
        if item["kind"] == "line":
            plt.plot(item["h"], item["v"], lw=3, color=c)
        elif item["kind"] == "marker":
            plt.plot(item["h"], item["v"], lw=3, marker="o", color=c)
            energy[item["cell"]] = 0              # the source cell would dominate the colour scale
        else:
            plt.gca().add_patch(ptc.Rectangle(xy=item["xy"], width=item["width"], height=item["height"],
                                              linewidth=0, edgecolor="none", facecolor=c))
    cmap_norm = LogNorm(vmin=1e-4, vmax=energy.max() + 1e-4) if norm == "log" else None
    plt.imshow(np.abs(energy), cmap=cmap, interpolation="sinc", norm=cmap_norm)
    rows, cols = _ROWS_COLS[axis]
    plt.ylabel("xyz"[rows])
    plt.xlabel("xyz"[cols])
    plt.ylim(energy.shape[0], -1)
    plt.xlim(-1, energy.shape[1])
    plt.figlegend()
    plt.tight_layout()
    if save:
        plt.savefig(os.path.join(folder, f"file{str(index).zfill(4)}.png"))
    if show:
        plt.show()
    return plt.gcf()


def peak_to_peak_dB(block_det, choose_axis=2):
    """10 log10 of the peak-to-peak swing of one field component over time, per (x, y) cell of the first z level
    of a BlockDetector record (time, nx, ny, nz, 3), relative to the smallest swing (fdtd/visualization.py:370-387)."""
    if block_det is None:
        raise ValueError("Function 'dBmap' requires a detector_readings object as parameter.")
    block_det = bd.numpy(block_det) if torch.is_tensor(block_det) else np.asarray(block_det)
    if block_det.ndim != 5:
        raise ValueError("Function 'dBmap' requires object of readings recorded by 'fdtd.BlockDetector'.")
    trace = block_det[:, :, :, 0, choose_axis]
    swing = trace.max(axis=0) - trace.min(axis=0)
    return 10 * np.log10(swing / swing.min())


def dB_map_2D(block_det=None, choose_axis=2, interpolation="spline16", show=True, style=None):
    """decibel map of a BlockDetector record (continuous sources), fdtd/visualization.py:339-395."""
    a = peak_to_peak_dB(block_det, choose_axis)
    plt, _, _ = _pyplot()
    if style is not None:
        plt.style.use(style)
    plt.ioff()
    plt.close()
    plt.title("dB map of Electrical waves in detector region")
    plt.imshow(a, cmap="inferno", interpolation=interpolation)
    cbar = plt.colorbar()
    cbar.ax.set_ylabel("dB scale", rotation=270)
    if show:
        plt.show()
    return plt.gcf()


def envelope_arrivals(detector_dict, specific_plot=None, verbose=True):
    """For every LineDetector record of a `Grid.save_data()` dictionary ("<name> (E)" / "<name> (H)", arrays
    (time, points, 3)): the Hilbert envelope of the first point's components and the step at which it peaks.
    Returns {"E" | "H": {component index: [(record name, envelope, arrival step), ...]}}
    (fdtd/visualization.py:410-440)."""
    from scipy.signal import hilbert
    if detector_dict is None:
        raise Exception("Function plotDetection() requires a dictionary of detector readings as 'detector_dict' "
                        "parameter.")
    out = {}
    for name, record in detector_dict.items():
        record = np.asarray(record)
        if record.ndim != 3:
            if verbose:
                print("Detector '{}' not LineDetector; dumped.".format(name))
            continue
        field = name[-2]
        if specific_plot is not None and field != specific_plot[0]:
            continue
        for comp in range(record.shape[2]):
            if specific_plot is not None and "xyz".index(specific_plot[1]) != comp:
                continue
            env = np.abs(hilbert(record[:, 0, comp]))
            out.setdefault(field, {}).setdefault(comp, []).append((name, env, int(np.argmax(env))))
    return out


def plot_detection(detector_dict=None, specific_plot=None, show=True, style=None):
    """intensity envelopes of LineDetector records over time and the time-of-arrival plot of a pulse
    (fdtd/visualization.py:398-480)."""
    data = envelope_arrivals(detector_dict, specific_plot)
    plt, _, _ = _pyplot()
    if style is not None:
        plt.style.use(style)
    plt.ioff()
    plt.close()
    single = specific_plot is not None
    side = 1 if single else 2
    for fig, field in enumerate("EH"):
        if field not in data:
            continue
        plt.figure(fig, figsize=(15, 15))
        for comp, items in data[field].items():
            plt.subplot(side, side, 1 if single else comp + 1)
            for name, env, _ in items:
                plt.plot(env, label=name)
            plt.title(field + "(" + "xyz"[comp] + ")")
            plt.xlabel("Time steps")
            plt.ylabel("Magnitude")
        plt.suptitle("Intensity profile")
    plt.legend()
    plt.show()
    for field, comps in data.items():
        plt.figure(figsize=(15, 15))
        for comp, items in comps.items():
            plt.plot([arrival for _, _, arrival in items], [name for name, _, _ in items], label="xyz"[comp])
        plt.title(field)
        plt.xlabel("Time of arrival (time steps)")
        plt.legend()
        plt.suptitle("Time-of-arrival plot")
    if show:
        plt.show()
    return plt.gcf()
