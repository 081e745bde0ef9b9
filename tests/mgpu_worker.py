"""Multi-GPU parity worker: run under torchrun (one rank per GPU).  Every rank checks the
rows it owns against the single-domain CPU oracle, bit for bit (results must not depend on
the decomposition).  Exit code != 0 on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mgb200  # noqa: E402
import oracle  # noqa: E402
import conftest  # noqa: E402
from conftest import rand_vec  # noqa: E402

# MGB200_TEST_EMU=1 (tests/test_emulated_library.py): the same checks on CPU ranks -- gloo instead of nccl for the
# bootstrap, the CUDA-on-CPU emulation build of the library (tests/host_emul), its file-based NCCL stand-in
EMU = conftest.EMULATED
if EMU:
    conftest.use_emulated_library()
DEV = "cpu" if EMU else "cuda"

pkg = mgb200.package
mgdist = __import__("importlib").import_module("multigrid_nikhil_c-_b200.dist")


def check(mg, level, got, want, what):
    a = mgdist.owned_slice(mg, level, got)
    b = mgdist.owned_slice(mg, level, want)
    if not np.array_equal(a, b):
        d = np.abs(a - b)
        raise AssertionError(f"rank {mg.rank}: {what}: max diff {d.max():.3e}, {int((a != b).sum())} mismatches")


def main():
    rank, world, local = mgdist.env_ranks()
    if EMU:
        dist.init_process_group("gloo")
        os.environ["LOCAL_RANK"] = "0"   # the emulation has one device
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    o = oracle.get()
    n_checks = 0
    cases = [(np.float64, 7, 4, "jacobi", True), (np.float64, 7, 4, "rbgs", True), (np.float64, 9, 6, "jacobi", True),
             (np.float64, 9, 5, "rbgs", True), (np.float64, 8, 5, "jacobi", False), (np.float64, 8, 6, "rbgs", False),
             (np.float32, 8, 6, "jacobi", True), (np.float32, 8, 5, "rbgs", True), (np.float64, 10, 7, "jacobi", True)]
    if os.environ.get("MGB200_WORKER_QUICK") == "1":   # bounded run for the CPU suite (tests/test_emulated_library.py)
        cases = [cases[0], cases[1], cases[7]]
    for dtype, level, aggl, smoother, fused in cases:
            if (1 << (aggl + 1)) // world < 8:
                continue
            x, b = rand_vec(level, dtype, 41), rand_vec(level, dtype, 42, 1e-3)
            sid = 0 if smoother == "jacobi" else 1
            if True:
                mg = mgdist.create(level, dtype=dtype, smoother=smoother, agglomerate_level=aggl, fused=fused)
                assert mg.info(mgb200.capi.MG_INFO_DISTRIBUTED, level) == 1
                assert mg.info(mgb200.capi.MG_INFO_DISTRIBUTED, aggl) == 0
                # smoother
                mg.set_u(level, x); mg.set_rhs(level, b); mg.smooth(level, 3)
                want = o.jacobirelaxation(x, b, 3) if sid == 0 else o.rbgs(x, b, 3)
                check(mg, level, mg.get_u(level), want, f"smooth {smoother} L{level} {dtype.__name__}")
                # residual + norm (identical on every rank)
                mg.set_u(level, x)
                nrm = mg.residual(level, norm=True)
                r = o.residual(x, b)
                check(mg, level, mg.get_r(level), r, "residual")
                assert abs(nrm - o.norm2(r)) <= 1e-12 * o.norm2(r)
                t = torch.tensor([nrm], dtype=torch.float64, device=DEV)
                lst = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(lst, t)
                assert all(float(v.item()) == nrm for v in lst), "norm differs between ranks"
                # cycles: V, W; FMG; solve
                for gamma in (1, 2):
                    p = oracle.Params(smoother=sid, gamma=gamma, nthreads=2)
                    mg.set_u(level, x); mg.set_rhs(level, b)
                    want = x
                    for k in range(2):
                        mg.cycle(level, 2, 2, gamma)
                        want = o.vcyclemultigrid(want, b, p)
                        check(mg, level, mg.get_u(level), want, f"cycle {k} gamma={gamma} {smoother} L{level} aggl{aggl}")
                        n_checks += 1
                # consecutive cycles in one call (mg_cycles; with MGB200_CHAIN=1 POST+PRE visit chains on the slabs)
                for nu, cnt in ((2, 3), (1, 2)):
                    p = oracle.Params(nu1=nu, nu2=nu, smoother=sid, nthreads=2)
                    mg.set_u(level, x); mg.set_rhs(level, b)
                    mg.cycles(cnt, level, nu, nu, 1)
                    want = x
                    for _ in range(cnt):
                        want = o.vcyclemultigrid(want, b, p)
                    check(mg, level, mg.get_u(level), want, f"{cnt} chained V({nu},{nu}) {smoother} L{level} aggl{aggl}")
                    n_checks += 1
                p = oracle.Params(smoother=sid, nthreads=2)
                check(mg, level, mg.fullmultigrid(b, 1, 2, 2), o.fullmultigrid(b, 1, p), "fmg")
                check(mg, level, mg.fullmultigrid(b, 2, 1, 1), o.fullmultigrid(b, 2, oracle.Params(nu1=1, nu2=1, smoother=sid, nthreads=2)), "fmg 2x V(1,1)")
                mg.set_rhs(level, b); mg.zero_u(level)
                k, rel, hist = mg.solve(1e-8, 30)
                u, ko, ho = o.solve(np.zeros_like(b), b, 1e-8, 30, p)
                assert k == ko, (k, ko)
                assert np.allclose(hist, ho, rtol=1e-10, atol=0)
                check(mg, level, mg.get_u(level), u, "solve")
                mg.close()
    dist.barrier()
    if rank == 0:
        extra = ""
        if EMU:   # message counts of the emulated NCCL (lets the caller see which schedule really ran)
            import ctypes
            L = mgb200.capi.lib()
            L.cuda_emu_nccl_sends.restype = L.cuda_emu_nccl_allgathers.restype = ctypes.c_longlong
            extra = f" sends={L.cuda_emu_nccl_sends()} allgathers={L.cuda_emu_nccl_allgathers()}"
        print(f"MGPU OK world={world} checks={n_checks}{extra}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
