"""Problem set-up around the solver path: host-side, like the reference's own set-up code.

* `load_vector`  generalises `globalforcefunction` (P:283-335: constant f = 4, zero Dirichlet ring) to a sampled
  f(x, y) and non-zero Dirichlet data g(x, y).  The reference lumps the P1 load to b_i = f h^2 per interior node
  (P:175-186 summed over the six triangles of a node); with a sampled f that is b_i = f(x_i, y_i) h^2.  Boundary
  values are eliminated into the right-hand side: a boundary neighbour of an interior node contributes +g to its
  row (A = [-1; -1 4 -1; -1]), so the device keeps its zero ring and every kernel stays as it is.
* `ProblemVar` / `multigrid_solver`  the call shape of the reference's second sketch (Multigrid_functions.cpp,
  "M:line"): one problem object carrying a per-level right-hand-side dictionary `b_dict` (M:16-26), solved by
  `multigrid_solver(obj)` = full multigrid from the coarsest level up (M:175-197), here over the structured-grid
  operators of libmgb200 (the unstructured-mesh transfer operators M:98-130 are out of scope).
  Like the sketch, the coarsest level is solved directly (`direct_solver`, M:63-72 / M:136-139: coarse_solver="exact");
  coarse_solver="sweeps" gives the first version's nu1+nu2 sweeps there (P:583-587).

This is set-up code: it builds host vectors and sequences C-ABI calls.  The solver arithmetic (smoothing, residual,
transfers, cycles) runs in libmgb200.so on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np

from .solver import Multigrid


def node_coordinates(level: int):
    """Coordinates of the interior nodes of the unit square on `level`: (x, y) with shape (n, n), row <-> y
    (P:227-228: index = (row-1) n + (col-1))."""
    N = 1 << level
    t = np.arange(1, N, dtype=np.float64) / N
    return np.meshgrid(t, t, indexing="xy")


def load_vector(level: int, f=4.0, g: Optional[Callable] = None, dtype=np.float64) -> np.ndarray:
    """b for -Laplace(u) = f on the unit square, u = g on the boundary (g=None: zero ring, the reference's case).

    f: constant or callable f(x, y) on arrays; g: callable g(x, y) evaluated on the boundary nodes."""
    N = 1 << level
    n = N - 1
    h = 1.0 / N
    x, y = node_coordinates(level)
    fv = f(x, y) if callable(f) else np.full((n, n), float(f))
    b = np.asarray(fv, dtype=np.float64) * h * h          # P:182-184: f * (h^2/2)/3 over six triangles
    if g is not None:
        t = np.arange(1, N, dtype=np.float64) / N
        zero, one = np.zeros(n), np.ones(n)
        b[0, :] += g(t, zero)        # bottom ring row y = 0 feeds interior row 1
        b[-1, :] += g(t, one)        # top ring row y = 1
        b[:, 0] += g(zero, t)        # left ring column x = 0
        b[:, -1] += g(one, t)        # right ring column x = 1
    return np.ascontiguousarray(b.reshape(-1), dtype=dtype)


@dataclass
class ProblemVar:
    """The problem object of the v2 sketch (M:16-26): levels, cycle parameters and the per-level load vectors.

    `b_dict[level]` is the load vector of that level (interior only, row-major).  Levels missing from the dictionary
    get the full-weighting restriction of the next finer one (what P:641 does for every level)."""
    finest_level: int = 5            # M:45
    coarsest_level: int = 1          # M:44 has 0 (no unknowns on a structured grid); 1 = one unknown
    mu0: int = 2                     # M:46: fullmultigrid runs mu0+1 cycles per level (M:187)
    mu1: int = 1                     # M:47
    mu2: int = 1                     # M:48
    omega: float = 2.0 / 3.0         # M:49 reads `4 / 5`, an integer division that evaluates to 0 (erratum); P:127 value
    smoother: str = "jacobi"
    coarse_solver: str = "exact"     # M:136-139: direct solve on the coarsest level ("sweeps": P:583-587)
    dtype: type = np.float64
    b_dict: Dict[int, np.ndarray] = field(default_factory=dict)


def fullmultigrid(mg: Multigrid, obj: ProblemVar, level: Optional[int] = None) -> np.ndarray:
    """M:175-191 on resident data: solve the coarser problem first, interpolate it as the initial guess (M:185),
    then mu0+1 V(mu1, mu2) cycles (M:186-188).  Returns the iterate of `level` (default: finest)."""
    top = obj.finest_level if level is None else level
    if top not in obj.b_dict:
        raise ValueError(f"b_dict has no load vector for level {top}")
    for l in range(top, obj.coarsest_level - 1, -1):      # right-hand sides: given, else restricted from above
        if l in obj.b_dict:
            mg.set_rhs(l, obj.b_dict[l])
        else:
            mg.restrict_rhs(l + 1)
    mg.zero_u(obj.coarsest_level)                          # M:176
    # mu0+1 consecutive cycles per level as ONE mg_cycles call: the same bits as mu0+1 mg_cycle calls, and the library
    # may fuse the post-smoothing of one cycle with the pre-smoothing of the next (visit chains)
    mg.cycles(obj.mu0 + 1, obj.coarsest_level, obj.mu1, obj.mu2, 1)   # coarsest level: direct solve (M:137) or nu1 + nu2 sweeps (P:583-587)
    for l in range(obj.coarsest_level + 1, top + 1):
        mg.prolong_set(l)                                  # M:185
        mg.cycles(obj.mu0 + 1, l, obj.mu1, obj.mu2, 1)     # M:186-188
    return mg.get_u(top)


def multigrid_solver(obj: ProblemVar, **ctx_kw) -> np.ndarray:
    """M:193-197: create the execution context, run full multigrid on the finest level's load vector."""
    with Multigrid(obj.finest_level, coarsest_level=obj.coarsest_level, dtype=obj.dtype, smoother=obj.smoother,
                   omega=obj.omega, coarse_solver=obj.coarse_solver, **ctx_kw) as mg:
        return fullmultigrid(mg, obj)
