"""ctypes binding of libmgb200.so (include/mgb200.h).  No CPU fallback: if the library
is missing or CUDA is unavailable, calls raise."""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MGB200_LIB: development hook for A/B timing of two builds of the same sources (tools/gpu_*.sh); never a CPU path
LIB_PATH = os.environ.get("MGB200_LIB") or os.path.join(_HERE, "lib", "libmgb200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "mgb200.h")

MG_OK, MG_ERR_ARG, MG_ERR_CUDA, MG_ERR_STATE, MG_ERR_COMM, MG_ERR_ALLOC = range(6)
MG_F64, MG_F32 = 0, 1
MG_SMOOTH_JACOBI, MG_SMOOTH_RBGS = 0, 1
MG_GRAPH, MG_FUSED, MG_COARSE_TAIL = 1, 2, 4
MG_COARSE_SWEEPS, MG_COARSE_EXACT = 0, 1
MG_COMM_ID_BYTES = 128
(MG_INFO_PITCH, MG_INFO_ROWS_STORED, MG_INFO_ROW_BEGIN, MG_INFO_ROW_END, MG_INFO_LAUNCHES,
 MG_INFO_DISTRIBUTED, MG_INFO_BYTES_ALLOCATED, MG_INFO_GRAPH_LAUNCHES, MG_INFO_AGGLOMERATE_LEVEL,
 MG_INFO_STORED_ROW_BEGIN, MG_INFO_STORED_ROW_END) = range(11)
(MG_OP_SMOOTH1, MG_OP_RESIDUAL, MG_OP_RESTRICT, MG_OP_PROLONG, MG_OP_PRE_FUSED, MG_OP_POST_FUSED,
 MG_OP_RESIDUAL_NORM, MG_OP_SMOOTH2, MG_OP_SMOOTH3, MG_OP_SMOOTH4, MG_OP_POSTPRE_FUSED) = range(11)


class MgConfig(ctypes.Structure):
    _fields_ = [
        ("finest_level", ctypes.c_int), ("coarsest_level", ctypes.c_int), ("dtype", ctypes.c_int),
        ("smoother", ctypes.c_int), ("omega", ctypes.c_double), ("restrict_weight", ctypes.c_double),
        ("device", ctypes.c_int), ("flags", ctypes.c_int), ("rank", ctypes.c_int), ("world", ctypes.c_int),
        ("agglomerate_level", ctypes.c_int), ("comm_id", ctypes.c_void_p),
        ("coarse_solver", ctypes.c_int),
    ]


class MgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmgb200 error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile libmgb200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j", "4"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libmgb200.so failed")
    return LIB_PATH


def declared_symbols() -> list:
    """Every function include/mgb200.h declares (used by the CPU export test)."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> ctypes.CDLL:
    """Load the library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    sig = {
        "mg_config_default": (None, [ctypes.POINTER(MgConfig)]),
        "mg_create": (ci, [ctypes.POINTER(vp), ctypes.POINTER(MgConfig)]),
        "mg_destroy": (ci, [vp]),
        "mg_last_error": (ctypes.c_char_p, [vp]),
        "mg_sync": (ci, [vp]),
        "mg_comm_id": (ci, [vp]),
        "mg_level_side": (ci, [ci]),
        "mg_level_of_size": (ci, [ctypes.c_size_t]),
        "mg_slab_rows": (ci, [ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]),
        "mg_get_info": (ci, [vp, ci, ci, ctypes.POINTER(ctypes.c_int64)]),
        "mg_force_constant": (ci, [vp, cd]),
        "mg_force_synthetic": (ci, [vp, ctypes.c_uint64]),
        "mg_checksum": (ci, [vp, ci, ci, ctypes.POINTER(ctypes.c_uint64)]),
        "mg_set_rhs_host": (ci, [vp, ci, vp]),
        "mg_set_u_host": (ci, [vp, ci, vp]),
        "mg_get_u_host": (ci, [vp, ci, vp]),
        "mg_get_rhs_host": (ci, [vp, ci, vp]),
        "mg_get_r_host": (ci, [vp, ci, vp]),
        "mg_zero_u": (ci, [vp, ci]),
        "mg_smooth": (ci, [vp, ci, ci]),
        "mg_residual": (ci, [vp, ci, ctypes.POINTER(cd)]),
        "mg_restrict": (ci, [vp, ci]),
        "mg_restrict_rhs": (ci, [vp, ci]),
        "mg_prolong_correct": (ci, [vp, ci]),
        "mg_prolong_set": (ci, [vp, ci]),
        "mg_cycle": (ci, [vp, ci, ci, ci, ci]),
        "mg_cycles": (ci, [vp, ci, ci, ci, ci, ci]),
        "mg_fmg": (ci, [vp, ci, ci, ci]),
        "mg_solve": (ci, [vp, cd, ci, ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(cd), vp]),
        "mg_host_jacobirelaxation": (ci, [vp, ci, vp, vp, ci]),
        "mg_host_restriction2d": (ci, [vp, ci, vp, vp]),
        "mg_host_interpolation2d": (ci, [vp, ci, vp, vp]),
        "mg_host_vcyclemultigrid": (ci, [vp, ci, vp, vp, ci, ci, ci]),
        "mg_host_fullmultigrid": (ci, [vp, vp, vp, ci, ci, ci]),
        "mg_plan_vcycle": (ci, [ci, ci, ci, ci, ci, ci, ci, ci, ctypes.POINTER(ci), ci, ctypes.POINTER(ci)]),
        "mg_time_op": (ci, [vp, ci, ci, ci, ctypes.POINTER(ctypes.c_float)]),
        "mg_time_cycle": (ci, [vp, ci, ci, ci, ci, ci, ctypes.POINTER(ctypes.c_float)]),
        "mg_time_phases": (ci, [vp, ci, ci, ci, ci, ci, ctypes.POINTER(cd), ctypes.POINTER(ci)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L
