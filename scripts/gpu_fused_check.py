"""Quick hardware check of the temporally fused E+H kernel variants (development tool):
    python scripts/gpu_fused_check.py [variant ...]      # default: 1 2 3
prints PASS / FAIL per variant (bit-equality with the two-half-step path) and, with --time, ms per step at 512^3."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fdtd_b200 as fd  # noqa: E402


def build(n, t):
    g = fd.Grid(shape=n, grid_spacing=77.5e-9, permittivity=1.3, permeability=1.1)
    g[0:t, :, :] = fd.PML(); g[-t:, :, :] = fd.PML()
    g[:, 0:t, :] = fd.PML(); g[:, -t:, :] = fd.PML()
    g[:, :, 0:t + 1] = fd.PML()
    g[:, :, -t:] = fd.PML()
    g[0, 3, 5] = fd.PointSource(period=9, amplitude=0.3)          # on a face, inside a PML
    g[n[0] // 2, n[1] // 2, n[2] // 2] = fd.PointSource(period=17)
    g[t + 2:n[0] - t - 2, t + 3:n[1] - t - 3, n[2] // 3] = fd.LineSource(period=23)
    g[1:n[0] - 1, n[1] // 2 + 1, n[2] // 2 + 2] = fd.LineDetector()
    return g


def main():
    import torch
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    variants = [int(a) for a in args] or [1]
    fd.set_backend("cuda.float32")
    n, t = (72, 64, 192), 6
    ref = build(n, t)
    ref.run(21, progress_bar=False)
    E0, H0 = ref.E.clone(), ref.H.clone()
    d0 = np.array(ref.detectors[0].E)
    for v in variants:
        g = build(n, t)
        g._fuse_eh = v
        try:
            g.run(21, progress_bar=False)
            active = bool(g._engine.lib.fdtd_fuse_eh_active(g._engine.desc))
            ok = active and torch.equal(g.E, E0) and torch.equal(g.H, H0) and np.array_equal(np.array(g.detectors[0].E), d0)
            print(f"variant {v}: {'PASS' if ok else 'FAIL'} (active={active})", flush=True)
        except Exception as exc:     # noqa: BLE001
            print(f"variant {v}: ERROR {exc}", flush=True)
    if "--time" in sys.argv:
        for v in [0] + variants:
            g = build((512, 512, 512), 10)
            g._fuse_eh = v
            g.run(4, progress_bar=False)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.run(40, progress_bar=False); b.record(); torch.cuda.synchronize()
            print(f"variant {v}: {a.elapsed_time(b) / 40:.3f} ms per step at 512^3 f32", flush=True)
            del g


if __name__ == "__main__":
    main()
