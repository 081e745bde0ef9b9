"""Host-side mirror of the reference's function surface over the C ABI.

The reference's solver path is a set of C++ free functions taking a `queue&` first
(/root/reference/Poissons_SYCL.cpp: jacobirelaxation P:125, restriction2d P:531,
interpolation2d P:337, vcyclemultigrid P:575, fullmultigrid P:629, globalforcefunction
P:283; level table built by main() P:661-690).  `Multigrid` is that `queue` + level
table; its methods keep the reference's names, argument meaning (host vectors, interior
only, row-major n*n) and return-by-value behaviour.  Resident-data methods (smooth,
residual, cycle, solve, ...) map 1:1 onto include/mgb200.h.

All arithmetic runs in libmgb200.so on the GPU; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import capi

_DT = {np.dtype(np.float64): capi.MG_F64, np.dtype(np.float32): capi.MG_F32}


def _vp(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


class Multigrid:
    """One multigrid context (== reference `queue&` + `jacobi_matrices`, P:24-33)."""

    def __init__(self, finest_level: int, coarsest_level: int = 1, dtype=np.float64,
                 smoother: str = "jacobi", omega: float = 2.0 / 3.0, restrict_weight: float = 0.25,
                 device: int = -1, graph: bool = True, fused: bool = True, coarse_tail: bool = True,
                 rank: int = 0, world: int = 1, agglomerate_level: int = 0,
                 comm_id: Optional[bytes] = None, coarse_solver: str = "sweeps"):
        self._lib = capi.lib()
        self._ctx = ctypes.c_void_p()
        self.dtype = np.dtype(dtype)
        if self.dtype not in _DT:
            raise ValueError("dtype must be float64 or float32")
        cfg = capi.MgConfig()
        self._lib.mg_config_default(ctypes.byref(cfg))
        cfg.finest_level = finest_level
        cfg.coarsest_level = coarsest_level
        cfg.dtype = _DT[self.dtype]
        cfg.smoother = {"jacobi": capi.MG_SMOOTH_JACOBI, "rbgs": capi.MG_SMOOTH_RBGS}[smoother]
        cfg.omega = omega
        cfg.restrict_weight = restrict_weight
        cfg.device = device
        cfg.flags = (capi.MG_GRAPH if graph else 0) | (capi.MG_FUSED if fused else 0) | \
                    (capi.MG_COARSE_TAIL if coarse_tail else 0)
        cfg.rank, cfg.world = rank, world
        cfg.agglomerate_level = agglomerate_level
        # "sweeps": nu1+nu2 sweeps on the coarsest level (P:583-587); "exact": direct solve there (direct_solver, M:63-72)
        cfg.coarse_solver = {"sweeps": capi.MG_COARSE_SWEEPS, "exact": capi.MG_COARSE_EXACT}[coarse_solver]
        self._comm_id = None
        if comm_id is not None:
            self._comm_id = ctypes.create_string_buffer(bytes(comm_id), capi.MG_COMM_ID_BYTES)
            cfg.comm_id = ctypes.cast(self._comm_id, ctypes.c_void_p)
        self.finest_level, self.coarsest_level = finest_level, coarsest_level
        self.rank, self.world = rank, world
        rc = self._lib.mg_create(ctypes.byref(self._ctx), ctypes.byref(cfg))
        if rc != capi.MG_OK:
            msg = self._lib.mg_last_error(None).decode()
            self._ctx = ctypes.c_void_p()
            raise capi.MgError(rc, msg)

    # -- plumbing ----------------------------------------------------------
    def _ck(self, rc: int):
        if rc != capi.MG_OK:
            raise capi.MgError(rc, self._lib.mg_last_error(self._ctx).decode())

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.mg_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def side(self, level: int) -> int:
        return (1 << level) - 1

    def _vec(self, level: int, a) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1)
        n = self.side(level)
        if a.size != n * n:
            raise ValueError(f"level {level} vectors have {n * n} entries, got {a.size}")
        return a

    def level_of(self, vec) -> int:
        """The reference infers the level from the vector length (P:583)."""
        lvl = self._lib.mg_level_of_size(int(np.asarray(vec).size))
        if lvl < 0:
            raise ValueError("vector length is not (2^L-1)^2")
        return lvl

    def info(self, what: int, level: int = 0) -> int:
        out = ctypes.c_int64()
        self._ck(self._lib.mg_get_info(self._ctx, what, level or self.finest_level, ctypes.byref(out)))
        return int(out.value)

    @property
    def launches(self) -> int:
        return self.info(capi.MG_INFO_LAUNCHES)

    def sync(self):
        self._ck(self._lib.mg_sync(self._ctx))

    # -- resident-data API (1:1 with mgb200.h) -------------------------------
    def force_constant(self, f: float = 4.0):
        self._ck(self._lib.mg_force_constant(self._ctx, float(f)))

    def force_synthetic(self, seed: int = 1234):
        """Benchmark right-hand side generated on the device: b = h^2 (2U-1), U from splitmix64(seed, global index)
        (mg_force_synthetic; tests/synth_ref.py restates it in numpy)."""
        self._ck(self._lib.mg_force_synthetic(self._ctx, ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF)))

    def checksum(self, level: int, which: int = 0) -> int:
        """Order-independent 64-bit checksum of this rank's owned values of u (0), f (1) or r (2) (mg_checksum)."""
        out = ctypes.c_uint64(0)
        self._ck(self._lib.mg_checksum(self._ctx, level, which, ctypes.byref(out)))
        return int(out.value)

    def set_rhs(self, level: int, f_h):
        self._ck(self._lib.mg_set_rhs_host(self._ctx, level, _vp(self._vec(level, f_h))))

    def set_u(self, level: int, u):
        self._ck(self._lib.mg_set_u_host(self._ctx, level, _vp(self._vec(level, u))))

    def _get(self, fn, level: int, out=None) -> np.ndarray:
        n = self.side(level)
        if out is None:
            out = np.zeros(n * n, dtype=self.dtype)
        self._ck(fn(self._ctx, level, _vp(out)))
        return out

    def get_u(self, level: int, out=None) -> np.ndarray:
        return self._get(self._lib.mg_get_u_host, level, out)

    def get_rhs(self, level: int, out=None) -> np.ndarray:
        return self._get(self._lib.mg_get_rhs_host, level, out)

    def get_r(self, level: int, out=None) -> np.ndarray:
        return self._get(self._lib.mg_get_r_host, level, out)

    # -- row-slab host buffers (multi-GPU): the C ABI takes FULL-grid host vectors but a rank only
    #    touches the interior rows it stores, [ya, yb).  A caller that holds just those rows passes the
    #    address the full vector would have had: base = slab - (ya-1)*n*itemsize. ------------------
    def slab_rows(self, level: int):
        """1-based interior rows [ya, yb) that set_*/get_* of this rank read or write."""
        n = self.side(level)
        ya = max(self.info(capi.MG_INFO_STORED_ROW_BEGIN, level), 1)
        yb = min(self.info(capi.MG_INFO_STORED_ROW_END, level), n + 1)
        return ya, yb

    def _slab_base(self, level: int, slab: np.ndarray) -> ctypes.c_void_p:
        n = self.side(level)
        ya, yb = self.slab_rows(level)
        if slab.dtype != self.dtype or not slab.flags.c_contiguous or slab.size != (yb - ya) * n:
            raise ValueError(f"slab buffer must be C-contiguous {self.dtype} with {(yb - ya) * n} entries (rows {ya}..{yb - 1})")
        off = (ya - 1) * n * self.dtype.itemsize
        base = slab.ctypes.data - off
        if base <= 0:
            raise ValueError("slab base address underflow")
        return ctypes.c_void_p(base)

    def set_rhs_slab(self, level: int, slab: np.ndarray):
        self._ck(self._lib.mg_set_rhs_host(self._ctx, level, self._slab_base(level, slab)))

    def set_u_slab(self, level: int, slab: np.ndarray):
        self._ck(self._lib.mg_set_u_host(self._ctx, level, self._slab_base(level, slab)))

    def get_u_slab(self, level: int, slab: np.ndarray):
        """Fills the owned rows inside `slab` (halo rows of the buffer are left untouched)."""
        self._ck(self._lib.mg_get_u_host(self._ctx, level, self._slab_base(level, slab)))

    def vcyclemultigrid_slab(self, level: int, u_slab: np.ndarray, f_slab: np.ndarray, nu1=2, nu2=2, gamma=1):
        """mg_host_vcyclemultigrid (P:575) on slab-sized host buffers: u_slab is in/out."""
        self._ck(self._lib.mg_host_vcyclemultigrid(self._ctx, level, self._slab_base(level, u_slab),
                                                   self._slab_base(level, f_slab), nu1, nu2, gamma))

    def zero_u(self, level: int):
        self._ck(self._lib.mg_zero_u(self._ctx, level))

    def smooth(self, level: int, nu: int):
        self._ck(self._lib.mg_smooth(self._ctx, level, nu))

    def residual(self, level: int, norm: bool = False) -> Optional[float]:
        if norm:
            v = ctypes.c_double()
            self._ck(self._lib.mg_residual(self._ctx, level, ctypes.byref(v)))
            return float(v.value)
        self._ck(self._lib.mg_residual(self._ctx, level, None))
        return None

    def restrict(self, fine_level: int):
        self._ck(self._lib.mg_restrict(self._ctx, fine_level))

    def restrict_rhs(self, fine_level: int):
        self._ck(self._lib.mg_restrict_rhs(self._ctx, fine_level))

    def prolong_correct(self, fine_level: int):
        self._ck(self._lib.mg_prolong_correct(self._ctx, fine_level))

    def prolong_set(self, fine_level: int):
        self._ck(self._lib.mg_prolong_set(self._ctx, fine_level))

    def cycle(self, level: Optional[int] = None, nu1: int = 2, nu2: int = 2, gamma: int = 1):
        self._ck(self._lib.mg_cycle(self._ctx, level or self.finest_level, nu1, nu2, gamma))

    def cycles(self, count: int, level: Optional[int] = None, nu1: int = 2, nu2: int = 2, gamma: int = 1):
        """`count` consecutive cycles (the loop P:646-648); same bits as `count` calls of cycle()."""
        self._ck(self._lib.mg_cycles(self._ctx, level or self.finest_level, nu1, nu2, gamma, count))

    def fmg(self, cycles_per_level: int = 1, nu1: int = 2, nu2: int = 2):
        self._ck(self._lib.mg_fmg(self._ctx, cycles_per_level, nu1, nu2))

    def solve(self, rtol: float = 1e-8, max_cycles: int = 50, nu1: int = 2, nu2: int = 2, gamma: int = 1):
        """Cycles until ||r||/||r0|| <= rtol.  Returns (cycles, relres, history)."""
        k = ctypes.c_int()
        rel = ctypes.c_double()
        hist = np.zeros(max_cycles + 1, dtype=np.float64)
        self._ck(self._lib.mg_solve(self._ctx, rtol, max_cycles, nu1, nu2, gamma, ctypes.byref(k),
                                    ctypes.byref(rel), _vp(hist)))
        return int(k.value), float(rel.value), hist[: k.value + 1].copy()

    def time_op(self, op: int, level: int, reps: int) -> float:
        """Device milliseconds for `reps` back-to-back launches of one operator."""
        ms = ctypes.c_float()
        self._ck(self._lib.mg_time_op(self._ctx, op, level, reps, ctypes.byref(ms)))
        return float(ms.value)

    def time_cycle(self, level: int, nu1: int, nu2: int, gamma: int, reps: int) -> float:
        """Device milliseconds for `reps` back-to-back cycles (CUDA events on the context's stream)."""
        ms = ctypes.c_float()
        self._ck(self._lib.mg_time_cycle(self._ctx, level, nu1, nu2, gamma, reps, ctypes.byref(ms)))
        return float(ms.value)

    def time_phases(self, level: int, nu1: int, nu2: int, gamma: int, reps: int) -> dict:
        """Row slabs, communication-avoiding schedule: device ms per cycle of every phase of the plan (mg_time_phases).
        Returns {} when this context does not run the plan, else {phase: {level: ms}} + {"ops_per_cycle": n}."""
        out = (ctypes.c_double * 160)()
        n = ctypes.c_int(0)
        self._ck(self._lib.mg_time_phases(self._ctx, level, nu1, nu2, gamma, reps, out, ctypes.byref(n)))
        if n.value == 0:
            return {}
        names = ["halo_exchange", "pre", "post", "allgather_rhs", "replicated_coarse_cycle"]
        res = {"ops_per_cycle": int(n.value)}
        for k, name in enumerate(names):
            d = {str(l): out[k * 32 + l] for l in range(32) if out[k * 32 + l] > 0.0}
            if d:
                res[name] = d
        return res

    # -- the reference's function surface (host vectors in, host vectors out) --
    def globalforcefunction(self, f: float = 4.0) -> np.ndarray:
        """P:283-335: the finest-level load vector b = f*h^2."""
        self.force_constant(f)
        return self.get_rhs(self.finest_level)

    def jacobirelaxation(self, v, fh, mu: int) -> np.ndarray:
        """P:125-147: `mu` smoothing sweeps; like the reference, `v` is updated in place
        when it is a writable array of the right dtype, and the result is returned."""
        level = self.level_of(v)
        out = np.array(self._vec(level, v), copy=True)
        self._ck(self._lib.mg_host_jacobirelaxation(self._ctx, level, _vp(out), _vp(self._vec(level, fh)), mu))
        if isinstance(v, np.ndarray) and v.dtype == self.dtype and v.flags.writeable:
            v.reshape(-1)[...] = out   # the reference mutates v and returns a copy (P:146)
        return out

    def restriction2d(self, vec_h) -> np.ndarray:
        """P:531-546."""
        fine = self.level_of(vec_h)
        m = self.side(fine - 1)
        out = np.zeros(m * m, dtype=self.dtype)
        self._ck(self._lib.mg_host_restriction2d(self._ctx, fine, _vp(self._vec(fine, vec_h)), _vp(out)))
        return out

    def interpolation2d(self, vec_2h) -> np.ndarray:
        """P:337-425."""
        coarse = self.level_of(vec_2h)
        n = self.side(coarse + 1)
        out = np.zeros(n * n, dtype=self.dtype)
        self._ck(self._lib.mg_host_interpolation2d(self._ctx, coarse + 1, _vp(self._vec(coarse, vec_2h)), _vp(out)))
        return out

    def vcyclemultigrid(self, vec_h, f_h, nu1: int = 2, nu2: int = 2, gamma: int = 1,
                        inplace: bool = False) -> np.ndarray:
        """P:575-627: one cycle on host vectors; returns the new iterate.  inplace=True
        overwrites `vec_h` (the C call is in/out) instead of returning a fresh vector."""
        level = self.level_of(vec_h)
        if inplace:
            out = vec_h.reshape(-1)
            if out.dtype != self.dtype or not out.flags.c_contiguous or not out.flags.writeable:
                raise ValueError("inplace needs a writable C-contiguous array of the context dtype")
        else:
            out = np.array(self._vec(level, vec_h), copy=True)
        self._ck(self._lib.mg_host_vcyclemultigrid(self._ctx, level, _vp(out), _vp(self._vec(level, f_h)),
                                                   nu1, nu2, gamma))
        return out

    def fullmultigrid(self, f_h, cycles_per_level: int = 1, nu1: int = 2, nu2: int = 2, out=None) -> np.ndarray:
        """P:629-650 (reference: cycles_per_level = mu0+1 = 31, nu1 = nu2 = 10).  `out`: optional result buffer."""
        level = self.finest_level
        if out is None:
            out = np.zeros(self.side(level) ** 2, dtype=self.dtype)
        elif out.dtype != self.dtype or out.size != self.side(level) ** 2 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous array of the context dtype and finest-level size")
        self._ck(self._lib.mg_host_fullmultigrid(self._ctx, _vp(self._vec(level, f_h)), _vp(out),
                                                 cycles_per_level, nu1, nu2))
        return out


def comm_id() -> bytes:
    """128-byte communicator id; create on rank 0 and broadcast to every rank."""
    buf = ctypes.create_string_buffer(capi.MG_COMM_ID_BYTES)
    rc = capi.lib().mg_comm_id(ctypes.cast(buf, ctypes.c_void_p))
    if rc != capi.MG_OK:
        raise capi.MgError(rc, capi.lib().mg_last_error(None).decode())
    return buf.raw

