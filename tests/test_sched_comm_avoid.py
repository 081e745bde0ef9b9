"""CPU check (gloo, world 2 and 4) of the communication-avoiding V-cycle schedule of csrc/sched.h.

The schedule is data: mg_plan_vcycle (C ABI, no GPU needed) returns the op list that fused.cu
executes on the GPUs.  Here the same ops are executed with the CPU oracle on full-size arrays whose
rows outside the slab a rank stores are NaN-poisoned.  If the plan fails to provide any row a kernel
needs -- wrong extent arithmetic, a missing exchange -- NaNs reach the owned rows and the comparison
with the single-domain oracle fails.  One deep exchange of u + one all-gather per cycle must suffice."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    """a TCP port nobody listens on right now (rendezvous of the gloo process group)"""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]
EXCH, PRE, POST, GATHER_F, REPL_CYCLE = range(5)


class TraceEmu:
    def __init__(self, lib, orc, top, aggl, rank, world, ns1, ns2, halos, rbgs=False):
        self.rbgs = rbgs
        self.lib, self.o, self.top, self.aggl, self.rank, self.world = lib, orc, top, aggl, rank, world
        self.ns1, self.ns2, self.halo = ns1, ns2, halos
        self.u, self.f, self.own = {}, {}, {}
        a, b = ctypes.c_int(), ctypes.c_int()
        for l in range(1, top + 1):
            if l > aggl:
                assert lib.mg_slab_rows(l, rank, world, ctypes.byref(a), ctypes.byref(b)) == 0
                self.own[l] = (a.value, b.value)
            else:
                self.own[l] = (1, self.n(l) + 1)
            self.u[l] = self.poison(l)
            self.f[l] = self.poison(l)
        self.exchanges = 0

    def n(self, l):
        return (1 << l) - 1

    def poison(self, l):
        return np.full((self.n(l), self.n(l)), np.nan)

    def stored(self, l):
        a, b = self.own[l]
        h = self.halo[l] if l > self.aggl else 0
        return max(a - h, 1), min(b + h, self.n(l) + 1)

    def keep(self, l, full, rows):
        out = self.poison(l)
        a, b = rows
        out[a - 1:b - 1] = np.asarray(full).reshape(self.n(l), self.n(l))[a - 1:b - 1]
        return out

    def exchange(self, l, arr, depth):
        a, b = self.own[l]
        reqs, bufs = [], []
        if self.rank > 0:
            reqs.append(dist.isend(torch.from_numpy(arr[a - 1:a - 1 + depth].copy()), self.rank - 1))
            t = torch.empty((depth, self.n(l)), dtype=torch.float64)
            reqs.append(dist.irecv(t, self.rank - 1))
            bufs.append((a - 1 - depth, t))
        if self.rank < self.world - 1:
            reqs.append(dist.isend(torch.from_numpy(arr[b - 1 - depth:b - 1].copy()), self.rank + 1))
            t = torch.empty((depth, self.n(l)), dtype=torch.float64)
            reqs.append(dist.irecv(t, self.rank + 1))
            bufs.append((b - 1, t))
        for r in reqs:
            r.wait()
        for row, t in bufs:
            arr[row:row + t.shape[0]] = t.numpy()
        self.exchanges += 1

    def smooth(self, u, f, ns):
        """ns pipeline stages: ns Jacobi sweeps, or ns/2 red-black sweeps (one stage per colour)"""
        return self.o.rbgs(u, f, ns // 2) if self.rbgs else self.o.jacobirelaxation(u, f, ns)

    def run(self, ops):
        o = self.o
        for kind, l, a, b in ops:
            if kind == EXCH:
                self.exchange(l, self.u[l] if a == 0 else self.f[l], b)
            elif kind == PRE:
                u1 = self.smooth(self.u[l].reshape(-1), self.f[l].reshape(-1), self.ns1)
                r = o.residual(u1, self.f[l].reshape(-1))
                coarse = o.restriction2d(r)
                self.u[l] = self.keep(l, u1, (a, b))
                lc = l - 1
                ca, cb = (a + 1) // 2, (b - 1) // 2 + 1          # coarse rows with centre 2I in [a, b)
                ca, cb = max(ca, 1), min(cb, self.n(lc) + 1)
                self.f[lc] = self.keep(lc, coarse, (ca, cb))
                sa, sb = self.stored(lc)
                self.u[lc] = self.keep(lc, np.zeros(self.n(lc) ** 2), (sa, sb))
            elif kind == GATHER_F:
                aa, bb = ctypes.c_int(), ctypes.c_int()
                self.lib.mg_slab_rows(l, self.rank, self.world, ctypes.byref(aa), ctypes.byref(bb))
                mine = self.f[l][aa.value - 1:bb.value - 1].copy()
                parts = [None] * self.world
                dist.all_gather_object(parts, mine)
                self.f[l] = np.concatenate(parts, axis=0)
                self.u[l] = np.zeros_like(self.f[l])
            elif kind == REPL_CYCLE:
                import oracle
                p = oracle.Params(nu1=self.ns1 // 2, nu2=self.ns2 // 2, smoother=1) if self.rbgs else oracle.Params(nu1=self.ns1, nu2=self.ns2)
                self.u[l] = o.vcyclemultigrid(self.u[l].reshape(-1), self.f[l].reshape(-1), p).reshape(self.n(l), self.n(l))
            elif kind == POST:
                e = self.u[l - 1].reshape(-1)
                u0 = o.prolong_correct(e, self.u[l].reshape(-1))
                u2 = self.smooth(u0, self.f[l].reshape(-1), self.ns2)
                self.u[l] = self.keep(l, u2, (a, b))


def _worker(rank, world, port, top, aggl, ns1, ns2, q, rbgs=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mgb200
        import oracle
        from conftest import rand_vec
        lib = mgb200.capi.lib()
        o = oracle.get()
        ops = (ctypes.c_int * 400)()
        halo = (ctypes.c_int * 32)()
        nops = lib.mg_plan_vcycle(top, aggl, world, rank, ns1, ns2, 0, 10 ** 6, ops, 100, halo)
        assert nops > 0, "plan not applicable"
        plan = [tuple(ops[4 * i:4 * i + 4]) for i in range(nops)]
        assert sum(1 for k in plan if k[0] == EXCH) == 1 and sum(1 for k in plan if k[0] == GATHER_F) == 1
        emu = TraceEmu(lib, o, top, aggl, rank, world, ns1, ns2, list(halo), rbgs)
        x, b = rand_vec(top, np.float64, 71), rand_vec(top, np.float64, 72, 1e-3)
        n = emu.n(top)
        # the rank holds only its owned rows of u (halo invalid: hv_u = 0) and f on all stored rows
        emu.u[top] = emu.keep(top, x, emu.own[top])
        emu.f[top] = emu.keep(top, b, emu.stored(top))
        emu.run(plan)
        pp = oracle.Params(nu1=ns1 // 2, nu2=ns2 // 2, smoother=1) if rbgs else oracle.Params(nu1=ns1, nu2=ns2)
        want = o.vcyclemultigrid(x, b, pp).reshape(n, n)
        a, bb = emu.own[top]
        got = emu.u[top][a - 1:bb - 1]
        assert not np.isnan(got).any(), "owned rows depend on rows the schedule did not provide"
        assert np.array_equal(got, want[a - 1:bb - 1]), "schedule result differs from the single-domain V-cycle"
        assert emu.exchanges == 1
        q.put((rank, "ok"))
    except Exception as ex:  # noqa: BLE001
        q.put((rank, f"FAIL: {type(ex).__name__}: {ex}"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,top,aggl,ns1,ns2,rbgs", [(2, 8, 5, 2, 2, False), (2, 9, 6, 2, 2, False), (4, 9, 6, 2, 2, False),
                                                        (2, 8, 6, 1, 1, False), (4, 9, 7, 2, 1, False), (2, 9, 5, 1, 2, False),
                                                        (2, 9, 7, 4, 4, True), (2, 9, 6, 2, 2, True)])
def test_comm_avoiding_schedule_is_sufficient(world, top, aggl, ns1, ns2, rbgs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, top, aggl, ns1, ns2, q, rbgs)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
    assert all(msg == "ok" for _, msg in res), res


def test_plan_extents_and_tightness(mgb):
    """Extents follow the recurrences of sched.h; shrinking any of them by one row must break sufficiency
    (checked analytically here: the recurrences are equalities)."""
    lib = mgb.capi.lib()
    ops = (ctypes.c_int * 400)()
    halo = (ctypes.c_int * 32)()
    n = lib.mg_plan_vcycle(14, 10, 8, 3, 2, 2, 0, 0, ops, 100, halo)
    plan = [tuple(ops[4 * i:4 * i + 4]) for i in range(n)]
    assert [k for k, *_ in plan] == [EXCH, EXCH, PRE, PRE, PRE, PRE, GATHER_F, REPL_CYCLE, POST, POST, POST, POST]
    assert list(halo)[11:15] == [10, 22, 46, 94]
    assert plan[0][3] == 94 and plan[1][3] == 93                 # u: e+NS1+2, f: e+NS1+1
    assert lib.mg_plan_vcycle(14, 10, 1, 0, 2, 2, 0, 0, ops, 100, halo) == -1      # single GPU: not applicable
    assert lib.mg_plan_vcycle(8, 4, 8, 0, 2, 2, 0, 0, ops, 100, halo) == -1        # slabs thinner than the halo


# ------------------------------------------------------------------------------------------------
# In-process simulation of ALL ranks (no process group): cheap enough for the real hierarchy shape of the
# 16385^2 / 8-GPU configuration (four distributed levels above the agglomeration level), scaled down.
# ------------------------------------------------------------------------------------------------
class SimRanks:
    def __init__(self, lib, orc, top, aggl, world, ns1, ns2, rbgs=False):
        self.lib, self.o, self.top, self.aggl, self.world = lib, orc, top, aggl, world
        self.emus, self.plans = [], []
        for r in range(world):
            ops = (ctypes.c_int * 400)()
            halo = (ctypes.c_int * 32)()
            nops = lib.mg_plan_vcycle(top, aggl, world, r, ns1, ns2, 0, 10 ** 6, ops, 100, halo)
            assert nops > 0
            self.plans.append([tuple(ops[4 * i:4 * i + 4]) for i in range(nops)])
            self.emus.append(TraceEmu(lib, orc, top, aggl, r, world, ns1, ns2, list(halo), rbgs))
        assert len({tuple(k for k, *_ in p) for p in self.plans}) == 1, "ranks disagree on the op sequence"

    def run(self):
        nops = len(self.plans[0])
        for i in range(nops):
            kind, l = self.plans[0][i][0], self.plans[0][i][1]
            if kind == EXCH:
                which, depth = self.plans[0][i][2], self.plans[0][i][3]
                arrs = [(e.u[l] if which == 0 else e.f[l]) for e in self.emus]
                snap = [a.copy() for a in arrs]
                for r, e in enumerate(self.emus):
                    a, b = e.own[l]
                    if r > 0:
                        pa, pb = self.emus[r - 1].own[l]
                        arrs[r][a - 1 - depth:a - 1] = snap[r - 1][pb - 1 - depth:pb - 1]
                    if r < self.world - 1:
                        na, nb = self.emus[r + 1].own[l]
                        arrs[r][b - 1:b - 1 + depth] = snap[r + 1][na - 1:na - 1 + depth]
            elif kind == GATHER_F:
                parts = []
                aa, bb = ctypes.c_int(), ctypes.c_int()
                for r, e in enumerate(self.emus):
                    self.lib.mg_slab_rows(l, r, self.world, ctypes.byref(aa), ctypes.byref(bb))
                    parts.append(e.f[l][aa.value - 1:bb.value - 1].copy())
                full = np.concatenate(parts, axis=0)
                for e in self.emus:
                    e.f[l] = full.copy()
                    e.u[l] = np.zeros_like(full)
            else:
                for r, e in enumerate(self.emus):
                    e.run([self.plans[r][i]])


@pytest.mark.parametrize("world,top,aggl,ns1,ns2,rbgs", [(8, 10, 6, 2, 2, False), (4, 9, 5, 2, 2, False), (8, 10, 7, 1, 2, False),
                                                        (4, 10, 6, 4, 4, True), (2, 7, 4, 2, 2, False)])
def test_comm_avoiding_schedule_all_ranks_in_process(mgb, orc, world, top, aggl, ns1, ns2, rbgs):
    import oracle
    from conftest import rand_vec
    lib = mgb.capi.lib()
    sim = SimRanks(lib, orc, top, aggl, world, ns1, ns2, rbgs)
    x, b = rand_vec(top, np.float64, 73), rand_vec(top, np.float64, 74, 1e-3)
    for e in sim.emus:
        e.u[top] = e.keep(top, x, e.own[top])
        e.f[top] = e.keep(top, b, e.stored(top))
    sim.run()
    pp = oracle.Params(nu1=ns1 // 2, nu2=ns2 // 2, smoother=1, nthreads=4) if rbgs else oracle.Params(nu1=ns1, nu2=ns2, nthreads=4)
    n = (1 << top) - 1
    want = orc.vcyclemultigrid(x, b, pp).reshape(n, n)
    for r, e in enumerate(sim.emus):
        a, bb = e.own[top]
        got = e.u[top][a - 1:bb - 1]
        assert not np.isnan(got).any(), f"rank {r}: owned rows depend on rows the schedule did not provide"
        assert np.array_equal(got, want[a - 1:bb - 1]), f"rank {r}"
