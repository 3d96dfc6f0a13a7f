"""Loader for libmocca_b200.so (the C ABI in include/mocca_b200.h).  There is NO CPU fallback: importing the
package works anywhere, but creating an env fails loudly unless the CUDA library is built and a B200 is present."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB200_LIB") or os.path.join(_HERE, "libmocca_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", f) for f in
           ("mb200.cu", "mb_core.cuh", "mb_env.cuh", "mb_tables.h", "generated/walker3d_model.h",
            "generated/monkey3d_model.h", "generated/cassie_model.h", "generated/child3d_model.h",
            "generated/mike_model.h", "generated/walker2d_model.h", "generated/crab2d_model.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

SYMBOLS = [
    "mb200_default_physics", "mb200_default_physics_for", "mb200_create", "mb200_destroy", "mb200_dims", "mb200_seed", "mb200_reset",
    "mb200_step", "mb200_step_host", "mb200_get_state", "mb200_set_state", "mb200_get_record", "mb200_set_record",
    "mb200_rng_words", "mb200_get_rng", "mb200_set_rng", "mb200_step_physics", "mb200_mass_matrix", "mb200_inverse_dynamics", "mb200_set_param",
    "mb200_set_param_array", "mb200_record_stride", "mb200_stats",
    "mb200_launch_count", "mb200_measure_fp32_peak", "mb200_last_error",
]


class Physics(C.Structure):
    """mb200_physics (include/mocca_b200.h)."""
    _fields_ = [("dt", C.c_float), ("substeps", C.c_int), ("iterations", C.c_int), ("gravity", C.c_float),
                ("erp_contact", C.c_float), ("erp_joint", C.c_float), ("linear_slop", C.c_float),
                ("lin_damping", C.c_float), ("ang_damping", C.c_float), ("max_coord_vel", C.c_float),
                ("limit_max_impulse", C.c_float), ("split_threshold", C.c_float), ("residual_threshold", C.c_float),
                ("ground_friction", C.c_float), ("has_ground", C.c_int), ("self_collision", C.c_int)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in SOURCES)
    if force or stale:
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
        subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmocca_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
                "there is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, ip = C.c_void_p, C.c_int
        L.mb200_last_error.restype = C.c_char_p
        L.mb200_launch_count.restype = C.c_longlong
        L.mb200_launch_count.argtypes = [vp]
        L.mb200_default_physics.argtypes = [C.POINTER(Physics)]
        L.mb200_default_physics_for.argtypes = [C.c_char_p, C.POINTER(Physics)]
        L.mb200_create.argtypes = [C.c_char_p, ip, ip, C.POINTER(Physics), C.POINTER(vp)]
        L.mb200_destroy.argtypes = [vp]
        L.mb200_destroy.restype = None
        L.mb200_dims.argtypes = [vp] + [C.POINTER(ip)] * 5
        L.mb200_seed.argtypes = [vp, vp, ip]
        L.mb200_reset.argtypes = [vp, vp, vp, vp]
        L.mb200_step.argtypes = [vp] * 8
        L.mb200_step_host.argtypes = [vp] * 7
        for f in ("mb200_get_state", "mb200_set_state", "mb200_get_record", "mb200_set_record", "mb200_mass_matrix"):
            getattr(L, f).argtypes = [vp, vp, vp]
        L.mb200_step_physics.argtypes = [vp, vp, vp, vp, vp]
        L.mb200_rng_words.argtypes = [vp]
        L.mb200_get_rng.argtypes = [vp, vp]
        L.mb200_set_rng.argtypes = [vp, vp]
        L.mb200_inverse_dynamics.argtypes = [vp, vp, vp, vp]
        L.mb200_set_param.argtypes = [vp, C.c_char_p, C.c_float]
        L.mb200_set_param_array.argtypes = [vp, C.c_char_p, vp, ip]
        L.mb200_record_stride.argtypes = [vp]
        L.mb200_stats.argtypes = [vp, C.POINTER(C.c_double), ip]
        L.mb200_measure_fp32_peak.argtypes = [ip, C.POINTER(C.c_double)]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise RuntimeError("libmocca_b200: " + lib().mb200_last_error().decode())
