"""ctypes binding of include/fdtd_b200.h (the C ABI of the CUDA engine).

The product opens exactly one library, fdtd_b200/libfdtd_b200.so (nvcc, sm_100a), and fails
loudly when it is missing: there is no CPU or PyTorch fallback for the hot path.
"""
import ctypes as C
import os

ABI_VERSION = 15
MAX_SLABS, MAX_POST, FUSED_MAX = 6, 16, 6
F32, F64, F32X = 0, 1, 2
CLS_VARY_E, CLS_VARY_H, CLS_ABSORB, CLS_OBJECT, CLS_ANISO, CLS_OVERLAP, CLS_ABSORB2 = 1, 2, 4, 8, 16, 32, 64
POST_PERIODIC, POST_PML_ADD = 0, 1
SRC_POINTS, SRC_BOX, SRC_FEEDBACK = 0, 1, 2
DET_FIELD, DET_CURRENT = 0, 1
PHASE_BEFORE, PHASE_AFTER, PHASE_SOURCES, PHASE_DETECTORS, PHASE_OBJECTS, PHASE_ALL = 1, 2, 4, 8, 16, 31

_vp = C.c_void_p


class Slab(C.Structure):
    _fields_ = [("axis", C.c_int32), ("lo", C.c_int32), ("thickness", C.c_int32), ("fused", C.c_int32),
                ("x0", C.c_int32), ("x1", C.c_int32), ("psi_count", C.c_int64),
                ("psi_E", _vp), ("psi_H", _vp), ("bE", _vp), ("cE", _vp), ("bH", _vp), ("cH", _vp)]


class Source(C.Structure):
    _fields_ = [("kind", C.c_int32), ("field", C.c_int32), ("comp", C.c_int32), ("n", C.c_int32),
                ("idx", _vp), ("profile", _vp), ("amplitude", C.c_double), ("box", C.c_int32 * 6),
                ("wave", _vp), ("wave_q0", C.c_int64), ("wave_len", C.c_int64), ("bbox", C.c_int32 * 6),
                ("impedance", C.c_double), ("spacing", C.c_double), ("feedback", _vp), ("record", _vp),
                ("record_capacity", C.c_int64)]


class Detector(C.Structure):
    _fields_ = [("n", C.c_int32), ("kind", C.c_int32), ("idx", _vp), ("pos", _vp), ("ring_E", _vp),
                ("ring_H", _vp), ("capacity", C.c_int64), ("bbox", C.c_int32 * 6), ("last", _vp),
                ("spacing", C.c_double)]


class DeepObject(C.Structure):
    _fields_ = [("kind", C.c_int32), ("box", C.c_int32 * 6), ("inv", _vp * 3), ("absorb", _vp * 3), ("mask", _vp)]


OBJ_PLAIN, OBJ_ANISO, OBJ_ABSORB = 0, 1, 2


class Desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dtype", C.c_int32),
                ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
                ("x_offset", C.c_int32), ("Nx_global", C.c_int32), ("pad0_", C.c_int32),
                ("plane", C.c_int64),
                ("E", _vp * 3), ("H", _vp * 3),
                ("courant", C.c_double), ("bg_inv_eps", C.c_double * 3), ("bg_inv_mu", C.c_double * 3),
                ("inv_eps", _vp * 3), ("inv_eps2", _vp * 3), ("inv_eps_grid", _vp * 3), ("absorb", _vp * 3), ("absorb2", _vp * 3),
                ("inv_mu", _vp * 3),
                ("tile_class", _vp), ("plane_class", _vp), ("tile_y", C.c_int32), ("tile_z", C.c_int32),
                ("n_slabs", C.c_int32), ("n_post", C.c_int32),
                ("slabs", Slab * MAX_SLABS),
                ("post_kind", C.c_int32 * MAX_POST), ("post_arg", C.c_int32 * MAX_POST),
                ("n_sources", C.c_int32), ("n_detectors", C.c_int32),
                ("sources", C.POINTER(Source)), ("detectors", C.POINTER(Detector)),
                ("x_chunk", C.c_int32), ("use_graphs", C.c_int32), ("dyn", _vp),
                ("fuse_eh", C.c_int32), ("pad2_", C.c_int32), ("E2", _vp * 3), ("H2", _vp * 3),
                ("fuse_post", C.c_int32), ("pad3_", C.c_int32),
                ("psi_E2", _vp * MAX_SLABS), ("n_deep", C.c_int32), ("h_wrap_ghost", C.c_int32),
                ("deep", C.POINTER(DeepObject)), ("x_wrap", C.c_int32), ("pad4_", C.c_int32)]


class Halo(C.Structure):
    """mirror of fdtd_halo: one rank's peer pointers, flags and push counts of the peer-to-peer halo exchange"""
    _fields_ = [("has_left", C.c_int32), ("has_right", C.c_int32),
                ("left_ghost_y", _vp), ("left_ghost_z", _vp), ("right_ghost_y", _vp), ("right_ghost_z", _vp),
                ("left_ghost_y2", _vp), ("left_ghost_z2", _vp), ("right_ghost_y2", _vp), ("right_ghost_z2", _vp),
                ("left_flag", _vp), ("right_flag", _vp), ("flags", _vp), ("error", _vp),
                ("count", C.c_int64 * 2), ("push_fused", C.c_int32 * 2), ("side_stream", _vp),
                ("timeout_ns", C.c_int64)]


EXPORTS = {
    # name: (restype, argtypes)
    "fdtd_abi_version": (C.c_int32, []),
    "fdtd_sizeof_desc": (C.c_int64, []),
    "fdtd_last_error": (C.c_char_p, []),
    "fdtd_launch_count": (C.c_int64, []),
    "fdtd_tile_shape": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fdtd_validate": (C.c_int, [C.POINTER(Desc)]),
    "fdtd_e_halfstep": (C.c_int, [C.POINTER(Desc), C.c_int32, C.c_int32, C.c_int64, C.c_int64, _vp]),
    "fdtd_h_halfstep": (C.c_int, [C.POINTER(Desc), C.c_int32, C.c_int32, C.c_int64, C.c_int64, _vp]),
    "fdtd_post_is_fused": (C.c_int, [C.POINTER(Desc)]),
    "fdtd_post_E": (C.c_int, [C.POINTER(Desc), C.c_int64, C.c_int64, _vp]),
    "fdtd_post_H": (C.c_int, [C.POINTER(Desc), C.c_int64, C.c_int64, _vp]),
    "fdtd_update_E": (C.c_int, [C.POINTER(Desc), C.c_int64, C.c_int64, _vp]),
    "fdtd_update_H": (C.c_int, [C.POINTER(Desc), C.c_int64, C.c_int64, _vp]),
    "fdtd_run": (C.c_int, [C.POINTER(Desc), C.c_int64, C.c_int64, C.c_int64, _vp]),
    "fdtd_fuse_eh_active": (C.c_int, [C.POINTER(Desc)]),
    "fdtd_ipc_export": (C.c_int, [_vp, _vp, C.POINTER(C.c_int64)]),
    "fdtd_ipc_import": (C.c_int, [_vp, C.c_int64, C.POINTER(_vp)]),
    "fdtd_halfstep_push": (C.c_int, [C.POINTER(Desc), C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                     _vp, _vp, _vp]),
    "fdtd_halo_push": (C.c_int, [C.POINTER(Desc), C.c_int32, _vp, _vp, _vp]),
    "fdtd_halo_signal": (C.c_int, [_vp, C.c_int64, _vp]),
    "fdtd_halo_wait": (C.c_int, [_vp, C.c_int64, _vp, C.c_int64, _vp]),
    "fdtd_sharded_halfstep": (C.c_int, [C.POINTER(Desc), C.POINTER(Halo), C.c_int32, C.c_int64, C.c_int64, _vp]),
    "fdtd_run_sharded": (C.c_int, [C.POINTER(Desc), C.POINTER(Halo), C.c_int64, C.c_int64, C.c_int64, _vp]),
    "fdtd_fuse_eh_sharded_active": (C.c_int, [C.POINTER(Desc), C.POINTER(Halo)]),
    "fdtd_halo_refresh": (C.c_int, [C.POINTER(Desc), C.POINTER(Halo), _vp]),
    "fdtd_sizeof_halo": (C.c_int64, []),
    "fdtd_post_phases": (C.c_int, [C.POINTER(Desc), C.c_int32, C.c_uint32, C.c_int64, C.c_int64, _vp]),
    "fdtd_post_part": (C.c_int, [C.POINTER(Desc), C.c_int32, C.c_int32, C.c_int64, C.c_int64, _vp]),
    "fdtd_dft_accumulate": (C.c_int, [C.c_int32, _vp, C.c_int64, C.c_int64, _vp, C.c_int32, _vp, _vp]),
}

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfdtd_b200.so")


def bind(path):
    """open a build of the C ABI and attach prototypes; checks ABI version and struct layout."""
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype, fn.argtypes = res, args
    if lib.fdtd_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI {lib.fdtd_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    if lib.fdtd_sizeof_desc() != C.sizeof(Desc):
        raise RuntimeError(f"{path}: sizeof(fdtd_desc) {lib.fdtd_sizeof_desc()} != binding {C.sizeof(Desc)}")
    if lib.fdtd_sizeof_halo() != C.sizeof(Halo):
        raise RuntimeError(f"{path}: sizeof(fdtd_halo) {lib.fdtd_sizeof_halo()} != binding {C.sizeof(Halo)}")
    return lib


_lib = None


def load():
    """the CUDA library; RuntimeError (never a fallback) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built "
                "(python -c 'import __graft_entry__ as g; g.build()'). fdtd_b200 has no CPU fallback.")
        _lib = bind(LIB_PATH)
    return _lib


class EngineError(RuntimeError):
    pass


def check(lib, rc):
    if rc != 0:
        raise EngineError(f"fdtd_b200 C-ABI error {rc}: {lib.fdtd_last_error().decode()}")
