"""Small scenes for compute-sanitizer (development tool): every kernel family on a few dozen steps.

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_scenes.py [fused]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fdtd_b200 as fd  # noqa: E402
import scenes  # noqa: E402


def main():
    fused = "fused" in sys.argv[1:] or "only-fused" in sys.argv[1:]
    for dtype in () if "only-fused" in sys.argv[1:] else ("float64", "float32", "float32x"):
        fd.set_backend("cuda." + dtype)
        for name in ("pml3d", "objects3d", "periodic3d", "stacked3d", "feed50", "quickstart2d"):
            g = scenes.SCENES[name][0](fd)
            g._use_graphs = False
            g.run(12, progress_bar=False)
            g._fuse_post = False                   # the separate source / detector / object-layer kernels too
            g._registration_count += 1
            for _ in range(4):
                g.step()
            torch.cuda.synchronize()
            print(dtype, name, float(g.E.abs().max()), flush=True)
    if fused:
        fd.set_backend("cuda.float32")
        n, t = (40, 36, 136), 4
        g = fd.Grid(shape=n, grid_spacing=77.5e-9)
        for key in ((slice(0, t),), (slice(-t, None),), (slice(None), slice(0, t)), (slice(None), slice(-t, None)),
                    (slice(None), slice(None), slice(0, t)), (slice(None), slice(None), slice(-t, None))):
            g[key] = fd.PML()
        g[20, 18, 70] = fd.PointSource(period=15)
        g[2:38, 18, 60] = fd.LineDetector()
        g._fuse_eh = 1
        g.run(10, progress_bar=False)
        torch.cuda.synchronize()
        assert g._engine.lib.fdtd_fuse_eh_active(g._engine.desc) == 1
        print("fused E+H", float(g.E.abs().max()), flush=True)
    print("done")


if __name__ == "__main__":
    main()
