"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

The reference is imported from where it lies under the alias `fdtd_ref` (all its
intra-package imports are relative, so the alias is safe) with an empty
matplotlib stub: `import fdtd` hard-imports matplotlib (fdtd/grid.py:88 ->
fdtd/visualization.py:12) which this image does not have, and nothing on the
stepping path uses it.  Nothing from the reference is copied into the repo: only
its numerical outputs (final E, H and every detector trace) are stored.

Two precisions per scene where marked:
  *_f64 : fdtd_ref.set_backend("numpy")            -- the reference default
  *_f32 : torch.set_default_dtype(float32) + fdtd_ref.set_backend("torch.float32")
          -- the only way to make the reference really compute in float32
          (fdtd/backend.py:43-49,79-89,281-284: every named backend is float64
          otherwise)
The two curl golden vectors of the reference's own tests are also stored
(tests/test_grid.py:47-164) by evaluating its curl_E / curl_H on the same seeds.
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("FDTD_REFERENCE", "/root/reference")


def load_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib.colors"].LogNorm = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    spec = importlib.util.spec_from_file_location(
        "fdtd_ref", os.path.join(REF, "fdtd", "__init__.py"),
        submodule_search_locations=[os.path.join(REF, "fdtd")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["fdtd_ref"] = mod
    spec.loader.exec_module(mod)
    return mod


F32_SCENES = ("pml3d", "objects3d", "periodic3d", "c4small", "feed50", "overlaps3d", "patch_antenna", "ring3d",
              "overlaps3d_stable", "stacked3d")
SPECTRA_SCENES = ("patch_antenna",)


def main():
    import torch
    import scenes
    ref = load_reference()

    # curl golden vectors, same seeds as the reference's tests
    from fdtd_ref.grid import curl_E, curl_H
    ref.set_backend("numpy")
    E = np.random.RandomState(0).randn(3, 3, 3, 3)
    H = np.random.RandomState(1).randn(3, 3, 3, 3)
    np.savez_compressed(os.path.join(HERE, "curls.npz"), E=E, H=H,
                        curl_E=curl_E(E), curl_H=curl_H(H))

    only = sys.argv[1:]
    for name, (build, steps) in scenes.SCENES.items():
        if only and name not in only:
            continue
        torch.set_default_dtype(torch.float64)
        ref.set_backend("numpy")
        g = build(ref)
        g.run(steps, progress_bar=False)
        out = scenes.dump(g)
        assert out["E"].dtype == np.float64
        np.savez_compressed(os.path.join(HERE, f"{name}_f64.npz"), steps=steps, **out)
        print(name, "f64", {k: v.shape for k, v in out.items()})
        if name in SPECTRA_SCENES:          # FrequencyRoutines on the same run (fdtd/fourier.py)
            spec = scenes.spectra(ref, g)
            np.savez_compressed(os.path.join(HERE, f"spectra_{name}.npz"), steps=steps, **spec)
            print(name, "spectra", {k: v.shape for k, v in spec.items()})
        if name in F32_SCENES:
            torch.set_default_dtype(torch.float32)
            ref.set_backend("torch.float32")
            g = build(ref)
            g.run(steps, progress_bar=False)
            out = scenes.dump(g)
            assert out["E"].dtype == np.float32, out["E"].dtype
            np.savez_compressed(os.path.join(HERE, f"{name}_f32.npz"), steps=steps, **out)
            print(name, "f32")
            torch.set_default_dtype(torch.float64)
    ref.set_backend("numpy")


if __name__ == "__main__":
    main()
