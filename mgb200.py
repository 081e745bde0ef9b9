"""Importable alias of the package directory `multigrid_nikhil_c-_b200/` (hyphenated
names cannot be written in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("multigrid_nikhil_c-_b200")
capi = _pkg.capi
Multigrid = _pkg.Multigrid
comm_id = _pkg.comm_id
ProblemVar = _pkg.ProblemVar
load_vector = _pkg.load_vector
multigrid_solver = _pkg.multigrid_solver
problem = _pkg.problem
package = _pkg
