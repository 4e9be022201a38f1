"""The C-ABI library: loads, exports every symbol include/fdtd_b200.h declares, and the ctypes
mirror of fdtd_desc has the library's layout.  No compute calls (CPU-only container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fdtd_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdtd_[a-zA-Z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as entry
    entry.build()
    return entry.OUT


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ("fdtd_e_halfstep", "fdtd_h_halfstep", "fdtd_post_E", "fdtd_post_H", "fdtd_update_E",
                 "fdtd_update_H", "fdtd_run", "fdtd_validate", "fdtd_tile_shape", "fdtd_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_binding_matches_library(lib_path):
    from fdtd_b200 import _capi
    lib = _capi.bind(lib_path)          # checks ABI version and sizeof(fdtd_desc)
    assert set(_capi.EXPORTS) == set(declared_functions())
    ty, tz = ctypes.c_int32(), ctypes.c_int32()
    assert lib.fdtd_tile_shape(_capi.F32, 1024, 1024, ctypes.byref(ty), ctypes.byref(tz)) == 0
    assert (ty.value, tz.value) == (8, 128)
    assert lib.fdtd_tile_shape(_capi.F64, 97, 1, ctypes.byref(ty), ctypes.byref(tz)) == 0
    assert (ty.value, tz.value) == (128, 1)
    assert lib.fdtd_tile_shape(7, 4, 4, ctypes.byref(ty), ctypes.byref(tz)) < 0
    assert b"bad argument" in lib.fdtd_last_error()


def test_validate_rejects_bad_descriptors(lib_path):
    from fdtd_b200 import _capi
    lib = _capi.bind(lib_path)
    d = _capi.Desc()
    assert lib.fdtd_validate(ctypes.byref(d)) == -1          # ABI version 0
    d.abi_version = _capi.ABI_VERSION
    d.dtype = 5
    assert lib.fdtd_validate(ctypes.byref(d)) == -1
    assert b"dtype" in lib.fdtd_last_error()
    d.dtype = _capi.F32
    d.Nx = d.Ny = d.Nz = 4
    d.Nx_global = 4
    d.plane = 15
    assert lib.fdtd_validate(ctypes.byref(d)) == -1
    assert b"plane" in lib.fdtd_last_error()
    d.plane = 16
    assert lib.fdtd_validate(ctypes.byref(d)) == -1          # null field pointers
    assert b"null field" in lib.fdtd_last_error()


def test_no_cpu_fallback_without_cuda():
    """the product refuses to run without a GPU instead of falling back (CPU container only)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fdtd_b200 as fd
    with pytest.raises(RuntimeError):
        fd.set_backend("cuda.float32")
    with pytest.raises(ValueError):
        fd.set_backend("numpy")
    with pytest.raises(ValueError):
        fd.set_backend("torch.float64")


def test_header_is_plain_c_and_layouts_agree(lib_path, tmp_path):
    """include/fdtd_b200.h compiles as C99 (the boundary is a C ABI, not a C++ one), and a C program linked
    against the library sees the same sizeof(fdtd_desc) and field offsets as the library and the ctypes mirror."""
    import shutil
    import subprocess
    from fdtd_b200 import _capi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "fdtd_b200.h"\n'
        "int main(void) {\n"
        '  printf("%d %zu %zu %zu %zu %zu %zu\\n", FDTD_ABI_VERSION, sizeof(fdtd_desc), offsetof(fdtd_desc, slabs),\n'
        "         offsetof(fdtd_desc, sources), offsetof(fdtd_desc, detectors), offsetof(fdtd_desc, absorb2),\n"
        "         offsetof(fdtd_desc, x_wrap));\n"
        '  printf("%d %lld\\n", (int)fdtd_abi_version(), (long long)fdtd_sizeof_desc());\n'
        "  return fdtd_validate(NULL) == FDTD_ERR_ARG ? 0 : 1;\n}\n")
    exe = tmp_path / "probe"
    libdir = os.path.dirname(lib_path)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:" + os.path.basename(lib_path),
                    "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    abi, size, o_slabs, o_src, o_det, o_abs2, o_wrap, lib_abi, lib_size = (int(v) for v in out)
    assert abi == lib_abi == _capi.ABI_VERSION
    assert size == lib_size == ctypes.sizeof(_capi.Desc)
    D = _capi.Desc
    assert (o_slabs, o_src, o_det, o_abs2, o_wrap) == (D.slabs.offset, D.sources.offset, D.detectors.offset,
                                                       D.absorb2.offset, D.x_wrap.offset)


def test_argument_checks_of_the_auxiliary_entry_points(lib_path):
    """entry points that validate before they launch can be exercised without a GPU."""
    from fdtd_b200 import _capi
    lib = _capi.bind(lib_path)
    null = ctypes.c_void_p(None)
    assert lib.fdtd_dft_accumulate(_capi.F32, null, 0, 16, null, 4, null, null) == 0          # nothing to do
    assert lib.fdtd_dft_accumulate(_capi.F32, null, 8, 16, null, 4, null, null) == -1         # null pointers
    assert b"null pointer" in lib.fdtd_last_error()
    assert lib.fdtd_dft_accumulate(9, null, 8, 16, null, 4, null, null) == -1
    assert lib.fdtd_dft_accumulate(_capi.F64, null, -1, 16, null, 4, null, null) == -1
    assert lib.fdtd_halo_signal(null, 1, null) == -1 and b"null flag" in lib.fdtd_last_error()
    assert lib.fdtd_ipc_export(null, null, None) == -1
    d = _capi.Desc()
    assert lib.fdtd_post_part(ctypes.byref(d), 0, 0, 0, 0, null) == -1                         # invalid descriptor
    assert lib.fdtd_fuse_eh_active(ctypes.byref(d)) == -1
