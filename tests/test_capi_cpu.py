"""CPU-only checks of the C-ABI boundary: the library loads, exports every symbol
include/mgb200.h declares, the pure host-logic entry points work, and creating a context
without a GPU fails loudly (no CPU fallback)."""
import ctypes

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(mgb):
    lib = mgb.capi.lib()
    syms = mgb.capi.declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_level_queries(mgb):
    lib = mgb.capi.lib()
    for level in range(1, 16):
        n = (1 << level) - 1
        assert lib.mg_level_side(level) == n
        assert lib.mg_level_of_size(n * n) == level        # the reference's int(log2(sqrt(size)+1)), P:583
    assert lib.mg_level_of_size(10) == -1
    assert lib.mg_level_side(0) == -1


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_slab_partition_covers_interior_and_nests(mgb, world):
    """Row slabs tile the interior rows 1..N-1 exactly, and coarse row I has the owner of
    fine row 2I on every level (so coarse rows never straddle ranks)."""
    lib = mgb.capi.lib()
    a, b = ctypes.c_int(), ctypes.c_int()
    for level in range(4, 15):
        N = 1 << level
        owner = np.full(N + 1, -1)
        for r in range(world):
            assert lib.mg_slab_rows(level, r, world, ctypes.byref(a), ctypes.byref(b)) == 0
            assert a.value < b.value
            assert np.all(owner[a.value:b.value] == -1)
            owner[a.value:b.value] = r
        assert np.all(owner[1:N] >= 0) and owner[0] == -1 and owner[N] == -1
        if level > 4:
            assert np.array_equal(owner[2:N:2], prev[1:N // 2])
        prev = owner
    assert lib.mg_slab_rows(3, 0, 3, ctypes.byref(a), ctypes.byref(b)) != 0


def test_config_default(mgb):
    cfg = mgb.capi.MgConfig()
    mgb.capi.lib().mg_config_default(ctypes.byref(cfg))
    assert cfg.omega == 2.0 / 3.0 and cfg.restrict_weight == 0.25 and cfg.world == 1 and cfg.coarsest_level == 1


def test_no_cpu_fallback(mgb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(mgb.capi.MgError) as ei:
        mgb.Multigrid(5)
    assert ei.value.code == mgb.capi.MG_ERR_CUDA
    assert "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_reported_not_crashed(mgb):
    lib = mgb.capi.lib()
    assert lib.mg_create(None, None) == mgb.capi.MG_ERR_ARG
    assert lib.mg_destroy(None) == mgb.capi.MG_ERR_ARG
    assert lib.mg_sync(None) == mgb.capi.MG_ERR_ARG
    assert lib.mg_cycle(None, 3, 2, 2, 1) == mgb.capi.MG_ERR_ARG
