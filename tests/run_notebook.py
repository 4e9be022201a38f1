"""Execute the code cells of one of the REFERENCE's example notebooks against fdtd_b200 (TEST INFRASTRUCTURE):

    python tests/run_notebook.py /root/reference/examples/00-quick-start.ipynb [...]

`fdtd` is aliased to fdtd_b200 by tests/refshim.py; matplotlib / IPython / line_profiler are forgiving stand-ins (every attribute is
a callable that returns itself), notebook magics are dropped.  Prints "OK <notebook>" per notebook, or the failing
cell, and exits non-zero on the first failure."""
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refshim  # noqa: E402,F401   (aliases fdtd -> fdtd_b200)


class _Anything:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter([self, self])


for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "IPython", "IPython.display",
           "line_profiler"):
    _mod = types.ModuleType(_m)
    _mod.__getattr__ = lambda name, _a=_Anything(): _a
    sys.modules[_m] = _mod


def run(path):
    nb = json.load(open(path))
    scope = {"__name__": "__main__"}
    t0 = time.time()
    for n, cell in enumerate(nb["cells"]):
        if cell["cell_type"] != "code":
            continue
        src = "\n".join(line for line in "".join(cell["source"]).splitlines() if not line.strip().startswith(("%", "!")))
        try:
            exec(compile(src, f"{os.path.basename(path)}:cell{n}", "exec"), scope)
        except Exception as exc:     # noqa: BLE001
            print(f"FAILED {path} cell {n}: {type(exc).__name__}: {exc}\n{src[:400]}")
            return False
    print(f"OK {path} {time.time() - t0:.1f}s")
    return True


if __name__ == "__main__":
    sys.exit(0 if all(run(p) for p in sys.argv[1:]) else 1)
