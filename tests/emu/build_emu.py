"""Build the serial-interpreter (CPU) variant of the kernels for the CPU test-suite.  TESTS ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "fdtd_b200", "csrc", "fdtd_b200.cu")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libfdtd_b200_emu.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "fdtd_b200", "csrc", "yee_kernels.cuh"),
            os.path.join(ROOT, "fdtd_b200", "csrc", "yee_fused_eh.cuh"),
            os.path.join(ROOT, "include", "fdtd_b200.h"), os.path.join(HERE, "cuda_emu.h")]
    extra = os.environ.get("FDTD_EMU_DEFS", "").split()      # e.g. "-DFDTD_FUSED_RT_ROWS=2" (kernel variants)
    force = force or bool(extra)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps)):
        return OUT
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-DFDTD_EMU",
           "-fPIC", "-shared", *extra, "-I", HERE, "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "fdtd_b200", "csrc"), SRC, "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
