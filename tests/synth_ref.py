"""tests/synth_ref.py -- TEST INFRASTRUCTURE: numpy restatements of two device-side helpers of libmgb200
(csrc/kernels.cuh: k_fill_synthetic, k_checksum), used to check them on small grids.  Never imported by the product."""
import numpy as np

_G = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(state, idx):
    """splitmix64 finaliser of state + (idx+1)*golden, vectorised over uint64 arrays (wraps mod 2^64)."""
    with np.errstate(over="ignore"):
        z = np.asarray(state, dtype=np.uint64) + (np.asarray(idx, dtype=np.uint64) + np.uint64(1)) * _G
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _idx(level, rows):
    n = (1 << level) - 1
    ya, yb = (1, n + 1) if rows is None else rows
    return np.arange((ya - 1) * n, (yb - 1) * n, dtype=np.uint64)


def synthetic_rhs(level, seed=1234, dtype=np.float64, rows=None):
    """mg_force_synthetic: interior vector in the reference layout, or only the node rows [rows[0], rows[1])."""
    z = splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF), _idx(level, rows))
    u01 = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    h = 1.0 / (1 << level)
    return ((h * h) * (2.0 * u01 - 1.0)).astype(dtype)


def checksum(level, values, rows=None):
    """mg_checksum of the node rows [rows[0], rows[1]) held row-major in `values` (float64 or float32)."""
    v = np.ascontiguousarray(values).reshape(-1)
    bits = v.view(np.uint64) if v.dtype == np.float64 else v.view(np.uint32).astype(np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum(splitmix64(bits, _idx(level, rows)), dtype=np.uint64))
