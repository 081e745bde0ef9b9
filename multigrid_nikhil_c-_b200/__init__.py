"""multigrid_nikhil_c-_b200 — B200-native (sm_100a) drop-in for the geometric-multigrid
Poisson path of nikhilTkur/Multigrid_Nikhil_C- (smoother, residual, restriction,
prolongation+correction, V/W/FMG cycles).

The directory name contains a hyphen; import it with
    importlib.import_module("multigrid_nikhil_c-_b200")
or through the repo-root alias module `mgb200`.

  capi    ctypes binding of lib/libmgb200.so (C ABI: include/mgb200.h)
  solver  `Multigrid`: the reference's function surface over that ABI
  dist    torch.distributed plumbing for one-process-per-GPU row slabs
  problem set-up helpers: sampled load vectors / Dirichlet data, the v2 `ProblemVar` + `multigrid_solver` call shape
"""
from . import capi
from .problem import ProblemVar, load_vector, multigrid_solver
from .solver import Multigrid, comm_id

__all__ = ["capi", "Multigrid", "comm_id", "ProblemVar", "load_vector", "multigrid_solver"]
