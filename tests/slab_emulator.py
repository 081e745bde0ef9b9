"""Host-logic emulation of the row-slab schedule of libmgb200 (csrc/ctx.cu + csrc/comm.cu)
on the CPU: same partition (mg_slab_rows from the C ABI), same exchange points, the level
operators supplied by the CPU oracle.  Every rank keeps FULL-size arrays whose rows outside
its stored slab (owned rows +- 1 halo row) are NaN-poisoned, so any dependence on data a
real rank would not have turns into NaN in its owned rows.  Test infrastructure only."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

import oracle


class SlabEmu:
    def __init__(self, lib, orc, level, aggl, rank, world, params: oracle.Params, dtype=np.float64):
        self.lib, self.o, self.level, self.aggl, self.rank, self.world, self.p = lib, orc, level, aggl, rank, world, params
        self.dtype = np.dtype(dtype)
        self.u, self.f, self.r, self.own = {}, {}, {}, {}
        a, b = ctypes.c_int(), ctypes.c_int()
        for l in range(params.coarsest_level, level + 1):
            n = (1 << l) - 1
            if world > 1 and l > aggl:
                assert lib.mg_slab_rows(l, rank, world, ctypes.byref(a), ctypes.byref(b)) == 0
                self.own[l] = (a.value, b.value)
            else:
                self.own[l] = (1, n + 1)
            self.u[l] = self._poison(l)
            self.f[l] = self._poison(l)
            self.r[l] = self._poison(l)

    # --- helpers (1-based interior row i <-> array row i-1) ---
    def n(self, l):
        return (1 << l) - 1

    def dist_level(self, l):
        return self.world > 1 and l > self.aggl

    def _poison(self, l):
        return np.full((self.n(l), self.n(l)), np.nan, dtype=self.dtype)

    def stored(self, l):
        a, b = self.own[l]
        return max(a - 1, 1), min(b + 1, self.n(l) + 1)

    def keep(self, l, full, rows):
        out = self._poison(l)
        a, b = rows
        out[a - 1:b - 1] = full.reshape(self.n(l), self.n(l))[a - 1:b - 1]
        return out

    def set(self, which, l, full):
        getattr(self, which)[l] = self.keep(l, np.asarray(full, dtype=self.dtype), self.stored(l))

    def exchange(self, l, arr):
        """halo rows: the neighbour's first / last owned row (comm_halo_exchange, depth 1)"""
        if not self.dist_level(l):
            return
        a, b = self.own[l]
        reqs, bufs = [], []
        if self.rank > 0:
            reqs.append(dist.isend(torch.from_numpy(arr[a - 1].copy()), self.rank - 1))
            t = torch.empty(self.n(l), dtype=torch.from_numpy(arr[0:1]).dtype)
            reqs.append(dist.irecv(t, self.rank - 1))
            bufs.append((a - 2, t))
        if self.rank < self.world - 1:
            reqs.append(dist.isend(torch.from_numpy(arr[b - 2].copy()), self.rank + 1))
            t = torch.empty(self.n(l), dtype=torch.from_numpy(arr[0:1]).dtype)
            reqs.append(dist.irecv(t, self.rank + 1))
            bufs.append((b - 1, t))
        for r in reqs:
            r.wait()
        for row, t in bufs:
            arr[row] = t.numpy()

    # --- operators, mirroring Ctx::smooth_t / residual_t / restrict_t / prolong_t ---
    def smooth(self, l, nu):
        for _ in range(nu):
            if self.p.smoother == 0:
                new = self.o.jacobirelaxation(self.u[l].reshape(-1), self.f[l].reshape(-1), 1, self.p.omega)
                self.u[l] = self.keep(l, new, self.own[l])
                self.exchange(l, self.u[l])
            else:
                for colour in (0, 1):
                    new = self.o.rbgs_half(self.u[l].reshape(-1), self.f[l].reshape(-1), colour)
                    self.u[l] = self.keep(l, new, self.own[l])
                    self.exchange(l, self.u[l])

    def residual(self, l):
        self.r[l] = self.keep(l, self.o.residual(self.u[l].reshape(-1), self.f[l].reshape(-1)), self.own[l])

    def restrict(self, l, from_rhs=False):
        src = self.f[l] if from_rhs else self.r[l]
        self.exchange(l, src)
        coarse = self.o.restriction2d(src.reshape(-1), self.p.restrict_weight)
        lc = l - 1
        if self.dist_level(l) and not self.dist_level(lc):
            a, b = ctypes.c_int(), ctypes.c_int()
            self.lib.mg_slab_rows(lc, self.rank, self.world, ctypes.byref(a), ctypes.byref(b))
            mine = coarse.reshape(self.n(lc), self.n(lc))[a.value - 1:b.value - 1].copy()
            parts = [None] * self.world
            dist.all_gather_object(parts, mine)          # comm_allgather_rows
            self.f[lc] = np.concatenate(parts, axis=0)
            assert self.f[lc].shape == (self.n(lc), self.n(lc))
        else:
            self.f[lc] = self.keep(lc, coarse, self.own[lc])
        if not from_rhs:
            self.u[lc] = self.keep(lc, np.zeros(self.n(lc) ** 2, dtype=self.dtype), self.stored(lc))

    def prolong(self, l, add=True):
        lc = l - 1
        e = self.u[lc].reshape(-1)
        if add:
            new = self.o.prolong_correct(e, np.nan_to_num(self.u[l], nan=0.0).reshape(-1))
            # rows that were not stored stay poisoned
            new = np.where(np.isnan(self.u[l]).reshape(-1), np.nan, new)
        else:
            new = self.o.interpolation2d(e)
        self.u[l] = self.keep(l, new, self.stored(l))   # owned + halo rows, no exchange needed

    def cycle(self, l):
        p = self.p
        self.smooth(l, p.nu1)
        if l <= p.coarsest_level:
            self.smooth(l, p.nu2)
            return
        self.residual(l)
        self.restrict(l)
        reps = 1 if l - 1 <= p.coarsest_level else max(1, p.gamma)
        for _ in range(reps):
            self.cycle(l - 1)
        self.prolong(l, True)
        self.smooth(l, p.nu2)

    def fmg(self, cycles):
        p = self.p
        for l in range(self.level, p.coarsest_level, -1):
            self.restrict(l, from_rhs=True)
        lc = p.coarsest_level
        self.u[lc] = self.keep(lc, np.zeros(self.n(lc) ** 2, dtype=self.dtype), self.stored(lc))
        for _ in range(cycles):
            self.cycle(lc)
        for l in range(lc + 1, self.level + 1):
            self.prolong(l, add=False)
            for _ in range(cycles):
                self.cycle(l)

    def owned(self, l, arr):
        a, b = self.own[l]
        return np.asarray(arr).reshape(self.n(l), self.n(l))[a - 1:b - 1]
