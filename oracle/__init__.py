"""oracle/ — CPU oracle for the multigrid Poisson hot path.  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``multigrid_nikhil_c-_b200``) never does; it fails loudly without its CUDA library.

Function names follow the reference (/root/reference/Poissons_SYCL.cpp, "P:line"):
jacobirelaxation P:125, restriction2d P:531, interpolation2d P:337,
vcyclemultigrid P:575, fullmultigrid P:629, globalforcefunction P:283.
Vectors are numpy arrays in the reference layout: interior-only, row-major n*n.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build(native: bool = False, quiet: bool = True) -> str:
    """Compile the C oracle (gcc) and return the path of the shared library."""
    target = "native" if native else "all"
    subprocess.run(["make", "-C", _HERE, target], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)
    return os.path.join(_BUILD, "libmg_oracle_native.so" if native else "libmg_oracle.so")


class _Params(ctypes.Structure):
    _fields_ = [("coarsest_level", ctypes.c_int), ("nu1", ctypes.c_int), ("nu2", ctypes.c_int),
                ("gamma", ctypes.c_int), ("smoother", ctypes.c_int), ("omega", ctypes.c_double),
                ("restrict_weight", ctypes.c_double), ("nthreads", ctypes.c_int), ("coarse_exact", ctypes.c_int)]


@dataclass
class Params:
    coarsest_level: int = 1
    nu1: int = 2
    nu2: int = 2
    gamma: int = 1
    smoother: int = 0          # 0 weighted Jacobi, 1 RB-GS
    omega: float = 2.0 / 3.0   # P:127
    restrict_weight: float = 0.25
    nthreads: int = 1
    coarse_exact: int = 0      # 1: exact solve on the coarsest level (M:63-72) instead of nu1+nu2 sweeps (P:583-587)

    def c(self) -> _Params:
        return _Params(self.coarsest_level, self.nu1, self.nu2, self.gamma, self.smoother,
                       self.omega, self.restrict_weight, self.nthreads, self.coarse_exact)


_SUF = {np.dtype(np.float64): "_f64", np.dtype(np.float32): "_f32"}
_CT = {np.dtype(np.float64): ctypes.c_double, np.dtype(np.float32): ctypes.c_float}


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _side(vec: np.ndarray) -> int:
    n = int(round(np.sqrt(vec.size)))
    if n * n != vec.size:
        raise ValueError(f"vector length {vec.size} is not a square")
    return n


def level_of(n: int) -> int:
    """Level from interior side length, the reference's int(log2(sqrt(size)+1)) (P:583)."""
    level = int(np.log2(n + 1))
    if (1 << level) - 1 != n:
        raise ValueError(f"interior side {n} is not 2^L-1")
    return level


class Oracle:
    """ctypes view of libmg_oracle.so with reference-named methods."""

    def __init__(self, native: bool = False, path: str | None = None):
        if path is None:
            path = os.path.join(_BUILD, "libmg_oracle_native.so" if native else "libmg_oracle.so")
            if not os.path.exists(path):
                path = build(native=native)
        self.path = path
        self.lib = ctypes.CDLL(path)
        self.lib.mgo_max_threads.restype = ctypes.c_int
        for suf in _SUF.values():
            getattr(self.lib, "mgo_sumsq" + suf).restype = ctypes.c_double
            getattr(self.lib, "mgo_solve" + suf).restype = ctypes.c_int
            getattr(self.lib, "mgo_csr_build" + suf).restype = ctypes.c_void_p

    def max_threads(self) -> int:
        return int(self.lib.mgo_max_threads())

    def _f(self, name: str, arr: np.ndarray):
        return getattr(self.lib, name + _SUF[arr.dtype])

    @staticmethod
    def _chk(*arrs: np.ndarray):
        dt = arrs[0].dtype
        for a in arrs:
            if a.dtype != dt or not a.flags.c_contiguous:
                raise ValueError("oracle arrays must share dtype and be C-contiguous")
        if dt not in _SUF:
            raise ValueError(f"unsupported dtype {dt}")

    # -- operators ---------------------------------------------------------
    def jacobi_constants(self, omega: float, dtype) -> tuple:
        dt = np.dtype(dtype)
        c0, c1 = _CT[dt](), _CT[dt]()
        getattr(self.lib, "mgo_jacobi_constants" + _SUF[dt])(ctypes.c_double(omega),
                                                            ctypes.byref(c0), ctypes.byref(c1))
        return c0.value, c1.value

    def jacobirelaxation(self, v, fh, mu, omega=2.0 / 3.0, nthreads=1):
        """P:125-147; returns the smoothed copy (v itself is not modified here)."""
        out = np.array(v, copy=True)
        self._chk(out, fh)
        self._f("mgo_jacobirelaxation", out)(_ptr(out), _ptr(fh), _side(out), int(mu),
                                             ctypes.c_double(omega), int(nthreads))
        return out

    def rbgs(self, v, fh, mu, nthreads=1):
        out = np.array(v, copy=True)
        self._chk(out, fh)
        self._f("mgo_rbgs", out)(_ptr(out), _ptr(fh), _side(out), int(mu), int(nthreads))
        return out

    def rbgs_half(self, v, fh, colour, nthreads=1):
        """One colour (0 red, 1 black) of one RB-GS sweep."""
        out = np.array(v, copy=True)
        self._chk(out, fh)
        self._f("mgo_rbgs_half", out)(_ptr(out), _ptr(fh), _side(out), int(colour), int(nthreads))
        return out

    def residual(self, v, fh, nthreads=1):
        """P:589-608."""
        self._chk(v, fh)
        r = np.empty_like(v)
        self._f("mgo_residual", v)(_ptr(v), _ptr(fh), _ptr(r), _side(v), int(nthreads))
        return r

    def norm2(self, r) -> float:
        self._chk(r)
        return float(np.sqrt(self._f("mgo_sumsq", r)(_ptr(r), _side(r))))

    def restriction2d(self, vec_h, w=0.25, nthreads=1):
        """P:531-546."""
        self._chk(vec_h)
        nh = _side(vec_h)
        m = (nh - 1) // 2
        out = np.empty(m * m, dtype=vec_h.dtype)
        self._f("mgo_restriction2d", vec_h)(_ptr(vec_h), nh, _ptr(out), ctypes.c_double(w), int(nthreads))
        return out

    def interpolation2d(self, vec_2h, nthreads=1):
        """P:337-425."""
        self._chk(vec_2h)
        m = _side(vec_2h)
        nh = 2 * m + 1
        out = np.empty(nh * nh, dtype=vec_2h.dtype)
        self._f("mgo_interpolation2d", vec_2h)(_ptr(vec_2h), m, _ptr(out), int(nthreads))
        return out

    def prolong_correct(self, vec_2h, vec_h, nthreads=1):
        """P:620-624: vec_h + interpolation2d(vec_2h)."""
        out = np.array(vec_h, copy=True)
        self._chk(out, vec_2h)
        self._f("mgo_prolong_correct", out)(_ptr(vec_2h), _side(vec_2h), _ptr(out), int(nthreads))
        return out

    def coarse_exact(self, f_h):
        """Exact solve A u = f on one level (direct_solver, M:63-72): sine-transform diagonalisation, see the .inc."""
        f_h = np.ascontiguousarray(f_h)
        out = np.zeros_like(f_h)
        self._f("mgo_coarse_exact", out)(_ptr(out), _ptr(f_h), _side(out))
        return out

    def globalforcefunction(self, level, f=4.0, dtype=np.float64):
        """P:283-335: b = f*h^2 at every interior node."""
        n = (1 << level) - 1
        out = np.empty(n * n, dtype=dtype)
        self._f("mgo_globalforcefunction", out)(_ptr(out), int(level), ctypes.c_double(f))
        return out

    # -- cycles ------------------------------------------------------------
    def vcyclemultigrid(self, vec_h, f_h, params: Params | None = None, inplace: bool = False):
        """P:575-627; returns the new iterate (inplace=True overwrites vec_h, as P:581 does)."""
        p = (params or Params()).c()
        out = vec_h if inplace else np.array(vec_h, copy=True)
        self._chk(out, f_h)
        self._f("mgo_vcyclemultigrid", out)(_ptr(out), _ptr(f_h), level_of(_side(out)), ctypes.byref(p))
        return out

    def fullmultigrid(self, f_h, cycles=1, params: Params | None = None):
        """P:629-650; `cycles` V-cycles per level (reference mu0+1 = 31)."""
        p = (params or Params()).c()
        self._chk(f_h)
        out = np.zeros_like(f_h)
        self._f("mgo_fullmultigrid", out)(_ptr(out), _ptr(f_h), level_of(_side(out)), int(cycles), ctypes.byref(p))
        return out

    def solve(self, vec_h, f_h, rtol=1e-8, max_cycles=50, params: Params | None = None):
        """Cycles until ||r||/||r0|| <= rtol.  Returns (u, cycles, history)."""
        p = (params or Params()).c()
        out = np.array(vec_h, copy=True)
        self._chk(out, f_h)
        hist = np.zeros(max_cycles + 1, dtype=np.float64)
        k = self._f("mgo_solve", out)(_ptr(out), _ptr(f_h), level_of(_side(out)), ctypes.c_double(rtol),
                                      int(max_cycles), _ptr(hist), ctypes.byref(p))
        return out, int(k), hist[: k + 1].copy()

    # -- reference-structured CSR path (CPU baseline A) ---------------------
    def csr_build(self, level, dtype=np.float64):
        return ctypes.c_void_p(getattr(self.lib, "mgo_csr_build" + _SUF[np.dtype(dtype)])(int(level)))

    def csr_free(self, handle, dtype=np.float64):
        getattr(self.lib, "mgo_csr_free" + _SUF[np.dtype(dtype)])(handle)

    def csr_jacobirelaxation(self, handle, v, fh, mu, omega=2.0 / 3.0, nthreads=1):
        out = np.array(v, copy=True)
        self._chk(out, fh)
        self._f("mgo_csr_jacobirelaxation", out)(handle, _ptr(out), _ptr(fh), int(mu),
                                                 ctypes.c_double(omega), int(nthreads))
        return out

    def csr_residual(self, handle, v, fh, nthreads=1):
        self._chk(v, fh)
        r = np.empty_like(v)
        self._f("mgo_csr_residual", v)(handle, _ptr(v), _ptr(fh), _ptr(r), int(nthreads))
        return r

    def csr_vcyclemultigrid(self, handles: dict, vec_h, f_h, params: Params | None = None):
        """handles: {level: csr handle} for every level from coarsest to this one."""
        p = (params or Params()).c()
        out = np.array(vec_h, copy=True)
        self._chk(out, f_h)
        level = level_of(_side(out))
        arr = (ctypes.c_void_p * (level + 1))()
        for l, h in handles.items():
            arr[l] = h
        self._f("mgo_csr_vcyclemultigrid", out)(arr, _ptr(out), _ptr(f_h), level, ctypes.byref(p))
        return out


_default: Oracle | None = None


def get() -> Oracle:
    global _default
    if _default is None:
        _default = Oracle()
    return _default
