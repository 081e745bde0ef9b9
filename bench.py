#!/usr/bin/env python
"""bench.py — V-cycle throughput of the B200-native multigrid path (BASELINE.json metric:
"V-cycle ms & grid-pt updates/s at 4097^2 fp64; smoother HBM GB/s vs B200 peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one V(2,2) cycle (full weighting, bilinear prolongation, coarsened to 3x3) over one synthetic
right-hand side.
  N = 1 : BASELINE configs[1]: 4097^2 fp64, weighted Jacobi (w = 2/3), one B200.
  N > 1 : BASELINE configs[2]: 16385^2 fp64, red-black Gauss-Seidel, row slabs over the N GPUs of one box (one process
          per GPU under torchrun, halo exchange over NVLink, coarse levels agglomerated).  The line carries a
          `strong_scaling` block: rank 0 alone runs the SAME workload on one GPU (same device-generated right-hand
          side), so speed-up and efficiency are same-workload numbers, and `mgpu_parity` says whether the N-rank
          iterate equals the 1-rank iterate bit for bit (64-bit checksums of the owned rows, mg_checksum).  The Jacobi
          cycle on the same grid is reported under `extra`.
`value` = grid-point updates per second of the whole job, inputs resident in HBM, timed with CUDA events on the
library's stream (max over ranks); K consecutive cycles per timed region through mg_cycles (the loop P:646-648).
`e2e` = the same metric through the reference-shaped host call mg_host_vcyclemultigrid (host vectors in, host vector
out: H2D of vec_h and f_h and D2H of the result inside the timed region), for pinned and for pageable caller buffers.
`--impl reference` times the CPU oracle port of the reference's algorithm on the same config with all host cores
(the reference itself needs DPC++/oneMKL and cannot be built here).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_point_updates_per_s"
UNIT = "updates/s"


def updates_per_cycle(level: int, coarsest: int, nu1: int, nu2: int, gamma: int = 1) -> int:
    """Sum over level visits of (nu1+nu2) * n_l^2 (89 413 008 for L=12..1 V(2,2))."""
    total, visits = 0, 1
    for l in range(level, coarsest - 1, -1):
        n = (1 << l) - 1
        total += visits * (nu1 + nu2) * n * n
        if l - 1 > coarsest:
            visits *= gamma
    return total


def measured_peak_gbs() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock + throttle reasons through NVML while the timed regions run."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device: int, period: float = 0.02):
        self.device, self.period = device, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self.h).gpu
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(mhz), int(util)))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t is not None:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        busy = [m for m, u in self.samples if u >= 50] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "samples_under_load": len(busy)}


class LegGuard:
    """Keeps the one JSON line safe from a hang in an informational leg of an N > 1 run.

    The core measurement (value, roofline, e2e) is in `line` before any optional leg starts.  Every optional leg of a
    multi-rank run is collective (tolerance solve, parity run, per-phase timing, the extra Jacobi timing): if one rank
    stalls inside NCCL, all do, the driver's timeout ends the job and the whole record is lost.  Each leg therefore runs
    under a deadline; when it passes, rank 0 prints the line as it stands (with `aborted_leg`) and every rank leaves with
    os._exit(0) -- every rank arms the same deadline at the same barrier.  The hung call holds no GIL (ctypes / torch
    release it), so the timer thread runs."""

    def __init__(self, rank: int, line: dict, enabled: bool, budget_s: float):
        self.rank, self.line, self.enabled, self.budget_s = rank, line, enabled, budget_s
        self._lock = threading.Lock()
        self._printed = False
        self._timer = None
        self._name = None

    def _fire(self):
        with self._lock:
            if not self._printed:
                self._printed = True
                self.line["aborted_leg"] = {"leg": self._name, "after_s": self.budget_s,
                                            "note": "informational leg did not finish; the core measurement above is complete"}
                if self.rank == 0:
                    print(json.dumps(self.line, default=str), flush=True)
        os._exit(0)

    def __call__(self, name: str):
        self._name = name
        return self

    def __enter__(self):
        if self.enabled:
            self._timer = threading.Timer(self.budget_s, self._fire)
            self._timer.daemon = True
            self._timer.start()
        return self

    def __exit__(self, etype, exc, tb):
        if self._timer is not None:
            self._timer.cancel()
            self._timer = None
        if exc is not None and isinstance(exc, Exception):
            # an informational leg must not take the core measurement down with it: note the error, go on (if the other
            # ranks are now waiting for this one inside a collective, the next leg's deadline ends the job cleanly)
            self.line.setdefault("leg_errors", {})[self._name] = f"{etype.__name__}: {exc}"
            return True
        return False

    def print_final(self):
        with self._lock:
            if self._printed:
                return
            self._printed = True
            if self.rank == 0:
                print(json.dumps(self.line), flush=True)


def synthetic_rhs(level: int, dtype) -> np.ndarray:
    """SURVEY 8d input (ii): b = h^2 * U(-1,1), numpy default_rng(1234), row-major interior order."""
    n = (1 << level) - 1
    h = 1.0 / (1 << level)
    return (h * h * np.random.default_rng(1234).uniform(-1.0, 1.0, n * n)).astype(dtype)


def default_workload(args, world):
    """(level, smoother) of the BASELINE config this N measures."""
    level = args.level or (12 if world == 1 else 14)
    smoother = args.smoother or ("jacobi" if world == 1 else "rbgs")
    return level, smoother


def workload_name(level, dtype_name, nu1, nu2, gamma, smoother):
    """config.workload: identical in both arms (the driver compares the strings)."""
    n = (1 << level) + 1
    sm = "weighted Jacobi (w=2/3)" if smoother == "jacobi" else "red-black Gauss-Seidel"
    return (f"{n}^2 {'fp64' if dtype_name == 'f64' else 'fp32'} V({nu1},{nu2}) gamma={gamma} {sm}, full weighting / bilinear, "
            f"coarsened to 3x3")


def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs must not inherit that)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def device_synthetic_rows(level: int, ya: int, yb: int, dtype, out: np.ndarray, seed: int = 1234):
    """Host copy of the device-generated right-hand side (mg_force_synthetic: b = h^2 (2U-1), U from splitmix64 of the
    global interior index) for node rows [ya, yb), written into `out` in row chunks (bounded temporaries)."""
    n = (1 << level) - 1
    h2 = (1.0 / (1 << level)) ** 2
    G = np.uint64(0x9E3779B97F4A7C15)
    step = max(1, (1 << 22) // n)
    o = out.reshape(yb - ya, n)
    with np.errstate(over="ignore"):
        for r0 in range(ya, yb, step):
            r1 = min(yb, r0 + step)
            idx = np.arange((r0 - 1) * n, (r1 - 1) * n, dtype=np.uint64)
            z = np.uint64(seed) + (idx + np.uint64(1)) * G
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z ^= z >> np.uint64(31)
            u01 = (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
            o[r0 - ya:r1 - ya] = (h2 * (2.0 * u01 - 1.0)).astype(dtype).reshape(r1 - r0, n)


class GpuLocalAffinity:
    """Bind the calling thread to the CPUs that are local to a GPU (NVML's CPU affinity) while the pinned host buffers
    are allocated and first touched, so that they live on the GPU's NUMA node; the previous affinity is restored on
    exit (the CPU baseline and OpenMP must see all cores).  Best effort: any failure leaves everything as it was."""

    def __init__(self, device: int):
        self.device, self.saved, self.cpus = device, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
            allowed = os.sched_getaffinity(0)
            cpus &= allowed
            if cpus and cpus != allowed:
                self.saved = allowed
                os.sched_setaffinity(0, cpus)
                self.cpus = len(cpus)
        except Exception:
            self.saved = None
        return self

    def __exit__(self, *a):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass


NUMA_CPUS = None   # CPUs the last pinned allocation was bound to (None: no binding happened)


def pinned(nelem: int, dtype, device=None):
    """Pinned host buffer; with `device` given it is allocated while the thread is bound to that GPU's local CPUs, so
    the pages land on the GPU's NUMA node.  The binding covers only the allocation call."""
    global NUMA_CPUS
    import torch
    tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
    if device is None:
        t = torch.empty(nelem, dtype=tdt, pin_memory=torch.cuda.is_available())
    else:
        with GpuLocalAffinity(device) as aff:
            t = torch.empty(nelem, dtype=tdt, pin_memory=torch.cuda.is_available())
            NUMA_CPUS = aff.cpus
    return t, t.numpy()


# ------------------------------------------------------------------------------------
# CPU legs (oracle): cpu_baseline of our arm and the whole --impl reference arm
# ------------------------------------------------------------------------------------
def load_oracle():
    import oracle
    try:
        return oracle.Oracle(native=True), "native"      # -march=native, built on this box
    except Exception:
        return oracle.get(), "portable"


def cpu_vcycle_rate(level, nu1, nu2, steps, warmup, smoother=0, gamma=1, with_csr=False, with_as_written=False,
                    device_rhs=False):
    import oracle
    o, build = load_oracle()
    nt = host_threads()
    p = oracle.Params(nu1=nu1, nu2=nu2, smoother=smoother, gamma=gamma, nthreads=nt)
    if device_rhs:
        n = (1 << level) - 1
        b = np.empty(n * n, dtype=np.float64)
        device_synthetic_rows(level, 1, n + 1, np.float64, b)
    else:
        b = synthetic_rhs(level, np.float64)
    u = np.zeros_like(b)
    for _ in range(warmup):
        u = o.vcyclemultigrid(u, b, p, inplace=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = o.vcyclemultigrid(u, b, p, inplace=True)
    dt = (time.perf_counter() - t0) / steps
    upd = updates_per_cycle(level, 1, nu1, nu2, gamma)
    sm = "weighted Jacobi" if smoother == 0 else "red-black Gauss-Seidel"
    out = {"value": upd / dt, "unit": UNIT, "cores": nt, "kind": "port", "ms_per_step": dt * 1e3,
           "sample": f"{steps} V({nu1},{nu2}) gamma={gamma} {sm} cycles at {(1 << level) + 1}^2 fp64 after {warmup} warm-up, "
                     f"matrix-free OpenMP oracle ({build} build), all {nt} host threads"}
    if with_csr:
        # CPU baseline A: the reference's own structure (assembled CSR SpMV + scal/add passes, P:138-144)
        lv = min(level, 11)
        hd = {l: o.csr_build(l) for l in range(1, lv + 1)}
        bb = synthetic_rhs(lv, np.float64)
        uu = o.csr_vcyclemultigrid(hd, np.zeros_like(bb), bb, p)
        t0 = time.perf_counter()
        uu = o.csr_vcyclemultigrid(hd, uu, bb, p)
        dta = time.perf_counter() - t0
        for h in hd.values():
            o.csr_free(h)
        out["reference_structured"] = {"value": updates_per_cycle(lv, 1, nu1, nu2) / dta, "unit": UNIT, "cores": nt,
                                       "ms_per_step": dta * 1e3,
                                       "sample": f"1 V({nu1},{nu2}) cycle at {(1 << lv) + 1}^2 fp64, CSR SpMV + scal/add "
                                                 f"passes as P:138-144 / P:604-607, {nt} threads"}
    if with_as_written:
        out["reference_as_written"] = reference_as_written()
    return out


def reference_as_written():
    """The reference's OWN source (oracle/_ref: Poissons_SYCL.cpp compiled where it lies against stub oneMKL / SYCL
    headers, single host thread) timed as written: fp32, 1025^2, levels 10..7, one vcyclemultigrid call with its own
    mu1 = mu2 = 10 (P:575, P:20-22).  As written it does not solve Poisson, so this is a time, not a
    parity claim."""
    import ctypes
    try:
        lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_poisson.so"))
    except OSError as ex:
        return {"unavailable": f"oracle/_ref not built ({ex})"}
    try:
        lib.ref_finest_level.restype = ctypes.c_int
        lib.ref_coarsest_level.restype = ctypes.c_int
        fin, coa = lib.ref_finest_level(), lib.ref_coarsest_level()
        mu = (ctypes.c_int * 3)()
        lib.ref_params(mu)
        n = (1 << fin) - 1
        f = np.full(n * n, 4.0 * (1.0 / (1 << fin)) ** 2, dtype=np.float32)
        v = np.zeros(n * n, dtype=np.float32)
        vp, fp = v.ctypes.data_as(ctypes.c_void_p), f.ctypes.data_as(ctypes.c_void_p)
        lib.ref_vcyclemultigrid.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        t0 = time.perf_counter()
        lib.ref_vcyclemultigrid(fin, vp, fp)        # first call: includes the FEM assembly of every level (P:661-690)
        first_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            lib.ref_vcyclemultigrid(fin, vp, fp)
        sec = (time.perf_counter() - t0) / reps
    except Exception as ex:  # noqa: BLE001 - informational leg only
        return {"unavailable": f"oracle/_ref failed: {ex}"}
    upd = (mu[1] + mu[2]) * sum(((1 << l) - 1) ** 2 for l in range(coa, fin + 1))
    return {"value": upd / sec, "unit": UNIT, "cores": 1, "kind": "reference", "ms_per_step": sec * 1e3,
            "first_call_incl_assembly_s": first_s,
            "sample": f"vcyclemultigrid (P:575) AS WRITTEN from the reference's own source: fp32 {n + 2}^2, levels {fin}..{coa}, "
                      f"mu1={mu[1]} + mu2={mu[2]} Jacobi sweeps per level (oneMKL calls served by the stub header, 1 thread), "
                      f"mean of {reps} calls after the assembling first call"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    level, smoother = default_workload(args, world)
    sid = 0 if smoother == "jacobi" else 1
    steps = args.steps
    r = cpu_vcycle_rate(level, args.nu1, args.nu2, steps, args.warmup, smoother=sid, gamma=args.gamma, device_rhs=world > 1)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(level, "f64", args.nu1, args.nu2, args.gamma, smoother), "level": level,
                       "arm": "CPU oracle port of the reference's algorithm (oracle/), all host cores"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import mgb200
    from mgb200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    comm = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # keep stdout to the single JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_comm_id():
        blob = [mgb200.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        return blob[0]

    if world > 1:
        comm = new_comm_id()

    level, smoother = default_workload(args, world)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    nu1, nu2, gamma = args.nu1, args.nu2, args.gamma
    n = (1 << level) - 1
    esize = np.dtype(dtype).itemsize
    SEED = 1234

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_u64_over_ranks(x: int) -> int:
        """Sum mod 2^64 (the checksum of a grid is the sum of the checksums of its row slabs)."""
        if dist is None:
            return x
        lst = [None] * world
        dist.all_gather_object(lst, int(x))
        return sum(lst) & 0xFFFFFFFFFFFFFFFF

    flags = dict(graph=not args.no_graph, fused=not args.no_fused, coarse_tail=not args.no_tail)
    mg = mgb200.Multigrid(level, dtype=dtype, smoother=smoother, device=local_rank, rank=rank, world=world,
                          comm_id=comm, agglomerate_level=args.aggl, **flags)
    slab = world > 1 and not args.full_host_vectors
    if world > 1:
        # the right-hand side is generated on the device from the GLOBAL index (mg_force_synthetic): identical on every
        # rank, for any number of ranks and for the single-GPU run of the same grid below; nothing crosses PCIe
        mg.force_synthetic(SEED)
        rhs_desc = f"h^2*(2U-1), U = splitmix64(seed {SEED}, global index), generated on the device"
    else:
        f_t, f_host = pinned(n * n, dtype, local_rank)           # on the GPU's NUMA node
        u_t, u_host = pinned(n * n, dtype, local_rank)
        f_host[:] = synthetic_rhs(level, dtype)
        u_host[:] = 0
        mg.set_rhs(level, f_host)
        rhs_desc = "h^2*U(-1,1) numpy default_rng(1234), row-major interior order (SURVEY 8d input ii)"
    mg.zero_u(level)
    upd = updates_per_cycle(level, 1, nu1, nu2, gamma)

    # ---- resident-data timing: exactly K cycles per region, several regions so that the
    #      clock sampler sees the load; the median region is reported ----
    K, W = args.steps, max(args.warmup, 3)
    mg.time_cycle(level, nu1, nu2, gamma, W)
    barrier()
    mg.time_cycle(level, nu1, nu2, gamma, K)      # builds (and caches) the K-cycle graph: not timed
    barrier()
    probe = mg.time_cycle(level, nu1, nu2, gamma, K)
    regions = int(min(200, max(5, 1500.0 / max(probe, 1e-3))))
    sampler = ClockSampler(local_rank)
    sampler.start()
    region_ms = []
    l0 = mg.launches
    for _ in range(regions):
        barrier()
        ms = mg.time_cycle(level, nu1, nu2, gamma, K)
        barrier()
        region_ms.append(max_over_ranks(ms))
    launches = (mg.launches - l0) // regions
    clocks = sampler.stop()
    ms_region = statistics.median(region_ms)
    ms_step = ms_region / K
    value = upd / (ms_step * 1e-3)
    # one cycle at a time (no POST+PRE fusion across cycle boundaries): what a single mg_cycle call costs
    mg.time_cycle(level, nu1, nu2, gamma, 1)
    iso = []
    for _ in range(20):
        barrier()
        iso.append(max_over_ranks(mg.time_cycle(level, nu1, nu2, gamma, 1)))
    isolated_ms = statistics.median(iso)

    # ---- per-kernel rooflines on the finest level (CUDA events on the library's stream) ----
    peak, peak_src = measured_peak_gbs()
    own_rows = n if world == 1 else (mg.info(capi.MG_INFO_ROW_END, level) - mg.info(capi.MG_INFO_ROW_BEGIN, level))
    pts = own_rows * n
    rb = smoother == "rbgs"
    kernels = {}
    reps = 20
    algo = {"smoother_sweep": (capi.MG_OP_SMOOTH1, 3.0), "two_sweeps_one_launch": (capi.MG_OP_SMOOTH2, 3.0), "residual": (capi.MG_OP_RESIDUAL, 3.0),
            "residual_norm_only": (capi.MG_OP_RESIDUAL_NORM, 2.0), "restrict": (capi.MG_OP_RESTRICT, 1.25),
            "prolong_correct": (capi.MG_OP_PROLONG, 2.25)}

    def timed_op(op):
        mg.time_op(op, level, 2)
        barrier()
        return max_over_ranks(mg.time_op(op, level, reps) / reps)

    for name, (op, s_per_pt) in algo.items():
        if world > 1 and name in ("restrict", "prolong_correct"):
            continue
        try:
            t = timed_op(op)
        except capi.MgError:
            continue
        gbs = s_per_pt * esize * pts / (t * 1e-3) / 1e9
        kernels[name] = {"ms": t, "algorithmic_bytes": s_per_pt * esize * pts, "GBps": gbs, "frac_of_peak": gbs / peak}
    pre_key, post_key, chain_key = "pre_fused(2 sweeps+residual+restrict)", "post_fused(prolong+correct+2 sweeps)", \
        "postpre_chain(prolong+correct+sweeps+residual+restrict)"
    for name, op, s_per_pt in ((pre_key, capi.MG_OP_PRE_FUSED, 3.25), (post_key, capi.MG_OP_POST_FUSED, 3.25),
                               (chain_key, capi.MG_OP_POSTPRE_FUSED, 3.5)):
        try:
            t = timed_op(op)
        except capi.MgError:
            continue
        kernels[name] = {"ms": t, "algorithmic_bytes": s_per_pt * esize * pts,
                         "GBps": s_per_pt * esize * pts / (t * 1e-3) / 1e9,
                         "frac_of_peak": s_per_pt * esize * pts / (t * 1e-3) / 1e9 / peak}
    # cumulative cost of the cycle from each level down (level_ms[l] - level_ms[l-1] = cost of level l's visit);
    # single cycles, so that every level is timed the same way
    level_ms = {}
    if world == 1:
        for l in range(max(2, min(6, level)), level + 1):
            try:
                mg.time_cycle(l, nu1, nu2, gamma, 1)
                level_ms[str(l)] = statistics.median([mg.time_cycle(l, nu1, nu2, gamma, 1) for _ in range(15)])
            except capi.MgError:
                pass
    # dominant kernel of the timed region: the fused PRE kernel on the finest level (largest single share of the
    # cycle, profiles/*launches*), else the plain sweep.  The plain smoother numbers are kept under "smoother".
    dom = pre_key if (pre_key in kernels and not args.no_fused) else "smoother_sweep"
    ns = (4 if rb else 2)
    tname = 'double' if esize == 8 else 'float'
    kname = (f"k_stream<{tname},{ns},PRE,{'rbgs' if rb else 'jacobi'}> (2 {'red-black GS' if rb else 'Jacobi'} "
             f"sweeps + residual + full weighting, finest level)") if dom == pre_key else "k_jacobi / k_rbgs (one sweep, finest level)"
    # K consecutive cycles with visit chains: the one big launch per cycle on the finest level is the POST+PRE chain
    # kernel (PRE and POST run once per K cycles), provided the cycle's sweep counts are the ones the chain fuses
    chain_used = (chain_key in kernels and K > 1 and os.environ.get("MGB200_CHAIN") != "0" and not args.no_fused and
                  ((not rb and nu1 == 2 and nu2 == 2) or (rb and nu1 == 1 and nu2 == 1)))
    if chain_used:
        dom = chain_key
        kname = (f"k_stream_chain<{tname},4,{'rbgs' if rb else 'jacobi'}> (prolongation + correction + nu2+nu1 sweeps + residual + "
                 f"full weighting in one launch, finest level; replaces POST + PRE = 6.5 S bytes per point by 3.5 S)")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{'chain' if dom == chain_key else ('pre' if dom == pre_key else 'sweep')}_{smoother}_L{level}_{args.dtype}_n{world}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["GBps"] / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                "ms_per_launch": kernels[dom]["ms"], "rows_per_rank": own_rows,
                "share_of_step": kernels[dom]["ms"] / ms_step,
                "two_launch_equivalent": ({"bytes": 6.5 * esize * pts, "GBps": 6.5 * esize * pts / (kernels[dom]["ms"] * 1e-3) / 1e9,
                                           "frac": 6.5 * esize * pts / (kernels[dom]["ms"] * 1e-3) / 1e9 / peak,
                                           "note": "the chain kernel does the work of POST + PRE (6.5 S algorithmic bytes in two launches)"}
                                          if chain_used else None),
                "smoother": {k: kernels[k] for k in ("smoother_sweep", "two_sweeps_one_launch") if k in kernels},
                "kernels": kernels, "cycle_ms_from_level_down": level_ms}

    # ---- end to end through the reference-shaped host call (P:575 on host vectors) ----
    e2e_steps = max(3, min(K, 10)) if not args.no_e2e else 1
    if slab:
        ya, yb = mg.slab_rows(level)
        f_t, f_host = pinned((yb - ya) * n, dtype, local_rank)
        u_t, u_host = pinned((yb - ya) * n, dtype, local_rank)
        device_synthetic_rows(level, ya, yb, dtype, f_host, SEED)
        h2d = 2 * (yb - ya) * n * esize
    else:
        if world > 1:
            f_t, f_host = pinned(n * n, dtype, local_rank)
            u_t, u_host = pinned(n * n, dtype, local_rank)
            device_synthetic_rows(level, 1, n + 1, dtype, f_host, SEED)
        h2d = 2 * n * n * esize

    def e2e_time(uh, fh):
        def call():
            if slab:
                mg.vcyclemultigrid_slab(level, uh, fh, nu1, nu2, gamma)
            else:
                mg.vcyclemultigrid(uh, fh, nu1, nu2, gamma, inplace=True)
        uh[:] = 0
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            call()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps

    e2e_ms = e2e_time(u_host, f_host)
    e2e = {"value": upd / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": own_rows * n * esize,
           "call": "mg_host_vcyclemultigrid (vcyclemultigrid P:575 on PINNED host vectors)",
           "host_buffers_on_gpu_numa_node_cpus": NUMA_CPUS}
    if world == 1 and not args.no_e2e:
        # the same call on ordinary (pageable) memory, which is what the std::vector surface of
        # include/mgb200_driver.hpp hands over
        try:
            up, fp = np.zeros(n * n, dtype=dtype), np.array(f_host, copy=True)
            pg_ms = e2e_time(up, fp)
            e2e["pageable"] = {"value": upd / (pg_ms * 1e-3), "unit": UNIT, "ms_per_step": pg_ms,
                               "call": "mg_host_vcyclemultigrid on PAGEABLE host vectors (the std::vector call shape)"}
            del up, fp
        except Exception as ex:  # noqa: BLE001 - informational leg only
            e2e["pageable"] = {"error": str(ex)}

    # informational: the reference's top-level call shape, fullmultigrid(f_h) -> u (P:629 / main P:727): one H2D of f,
    # one V(2,2) per level on the way up, one D2H of u.  Transfers are amortised over ~4/3 cycles' worth of work.
    if world == 1 and not args.no_e2e:
        try:
            fmg_upd = sum(updates_per_cycle(l, 1, nu1, nu2) for l in range(1, level + 1))
            mg.fullmultigrid(f_host, 1, nu1, nu2, out=u_host)
            t0 = time.perf_counter()
            for _ in range(3):
                mg.fullmultigrid(f_host, 1, nu1, nu2, out=u_host)
            fmg_ms = (time.perf_counter() - t0) * 1e3 / 3
            e2e["fullmultigrid_call"] = {"ms": fmg_ms, "value": fmg_upd / (fmg_ms * 1e-3), "unit": UNIT,
                                         "call": "mg_host_fullmultigrid, 1 V(2,2) per level, pinned host f in / pinned host u out"}
            # the same full-multigrid pass on resident data (mg_fmg; wall clock around a stream sync)
            mg.fmg(1, nu1, nu2)
            mg.sync()
            t0 = time.perf_counter()
            for _ in range(5):
                mg.fmg(1, nu1, nu2)
            mg.sync()
            fmg_res_ms = (time.perf_counter() - t0) * 1e3 / 5
            e2e["fullmultigrid_call"]["resident_ms"] = fmg_res_ms
            e2e["fullmultigrid_call"]["resident_value"] = fmg_upd / (fmg_res_ms * 1e-3)
        except Exception as ex:  # noqa: BLE001 - informational leg only
            e2e["fullmultigrid_call"] = {"error": str(ex)}

    # ---- the line as far as the core measurement goes; the legs below add to it.  At N > 1 every one of them is
    #      collective, so each runs under a deadline (LegGuard): a stall there costs that leg, not the record ----
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": workload_name(level, args.dtype, nu1, nu2, gamma, smoother),
                       "decomposition": "one GPU" if world == 1 else f"row slabs over {world} GPUs, halo exchange over NVLink (NCCL), coarse levels agglomerated",
                       "level": level, "smoother": smoother, "rhs": rhs_desc, "updates_per_cycle": upd,
                       "l2_policy": "inputs larger than L2 (4 arrays x %.0f MB on the finest level)" % (n * n * esize / 1e6),
                       "regions": regions, "region_stat": "median",
                       "agglomerate_level": mg.info(capi.MG_INFO_AGGLOMERATE_LEVEL, level) if world > 1 else (args.aggl or None),
                       "timed_as": f"{K} consecutive cycles per region through mg_time_cycle (mg_cycles: POST of one cycle and PRE of "
                                   f"the next are one launch on the finest level); isolated_cycle_ms = one mg_cycle at a time",
                       "scaling_note": ("N=1 measures BASELINE configs[1] (4097^2 Jacobi), N>1 measures configs[2] (16385^2 RB-GS): "
                                        "same-workload speed-up / efficiency are in strong_scaling, not in value(N)/value(1)"),
                       "env_knobs": {k: v for k, v in sorted(os.environ.items()) if k.startswith("MGB200_")},
                       "flags": flags},
            "isolated_cycle_ms": isolated_ms,
            "finest_points_per_s": n * n / (ms_step * 1e-3),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    guard = LegGuard(rank, line, enabled=world > 1, budget_s=args.leg_timeout)

    # tolerance-controlled solve on the same right-hand side (the reference runs a fixed number of cycles
    # and prints only the vector length): cycle count, residual history, wall time incl. the per-cycle norm
    with guard("solve"):
        try:
            if world > 1:
                mg.force_synthetic(SEED)
            mg.zero_u(level)
            mg.solve(1e-8, 2, nu1, nu2, gamma)      # warm (graphs of the solve path)
            mg.zero_u(level)
            mg.sync()
            barrier()
            t0 = time.perf_counter()
            k_cyc, relres, hist = mg.solve(1e-8, 40, nu1, nu2, gamma)
            mg.sync()
            solve_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
            line["solve"] = {"rtol": 1e-8, "cycles": k_cyc, "relres": relres, "ms": solve_ms, "ms_per_cycle": solve_ms / max(k_cyc, 1),
                             "overhead_vs_isolated_cycle": solve_ms / max(k_cyc, 1) / isolated_ms - 1.0,
                             "residual_history": [float(h) for h in hist],
                             "factors": [float(hist[i + 1] / hist[i]) for i in range(len(hist) - 1) if hist[i] > 0]}
        except Exception as ex:  # noqa: BLE001 - informational leg only
            line["solve"] = {"error": str(ex)}

    # ---- N > 1: multi-GPU parity record + the same workload on ONE GPU (rank 0 alone), the strong-scaling denominator.
    #      Parity: from u = 0 on the same device-generated right-hand side, two single cycles and one 3-cycle run
    #      (mg_cycles: visit chains) on N ranks and on 1 rank; the iterates agree bit for bit iff the checksums do. ----
    PAR = "u=0; 2 x mg_cycle; mg_cycles(3); 64-bit checksum of the owned rows (mg_checksum), summed over the ranks"

    def parity_run(m):
        m.force_synthetic(SEED)
        m.zero_u(level)
        m.cycle(level, nu1, nu2, gamma)
        m.cycle(level, nu1, nu2, gamma)
        m.cycles(3, level, nu1, nu2, gamma)
        return m.checksum(level, 0)

    if world > 1 and not args.no_n1:
        with guard("strong_scaling"):
            csum_n = sum_u64_over_ranks(parity_run(mg))
            strong = None
            if rank == 0:
                try:
                    mg1 = mgb200.Multigrid(level, dtype=dtype, smoother=smoother, device=local_rank, **flags)
                    csum_1 = parity_run(mg1)
                    mg1.time_cycle(level, nu1, nu2, gamma, W)
                    mg1.time_cycle(level, nu1, nu2, gamma, K)
                    ms1 = statistics.median([mg1.time_cycle(level, nu1, nu2, gamma, K) for _ in range(5)]) / K
                    mg1.time_cycle(level, nu1, nu2, gamma, 1)
                    iso1 = statistics.median([mg1.time_cycle(level, nu1, nu2, gamma, 1) for _ in range(9)])
                    mg1.close()
                    strong = {"workload": workload_name(level, args.dtype, nu1, nu2, gamma, smoother) + ", 1 GPU (rank 0 alone), same right-hand side",
                              "n1_ms_per_step": ms1, "n1_value": upd / (ms1 * 1e-3), "n_gpus": world, "ms_per_step": ms_step,
                              "speedup": ms1 / ms_step, "efficiency": ms1 / ms_step / world,
                              "n1_isolated_cycle_ms": iso1, "isolated_cycle_ms": isolated_ms, "isolated_efficiency": iso1 / isolated_ms / world,
                              "mgpu_parity": bool(csum_1 == csum_n), "parity_check": PAR,
                              "checksum_1gpu": f"{csum_1:016x}", "checksum_ngpu": f"{csum_n:016x}"}
                except Exception as ex:  # noqa: BLE001 - informational leg only
                    strong = {"error": str(ex)}
                line["strong_scaling"] = strong
                line["n1_same_workload"] = ({"ms_per_step": strong["n1_ms_per_step"], "value": strong["n1_value"], "unit": UNIT,
                                             "workload": strong["workload"]} if "n1_ms_per_step" in strong else strong)
            barrier()

    # ---- N > 1: where the distributed cycle goes.  Device time per phase of the communication-avoiding plan from eagerly
    #      launched cycles with CUDA events around every op (mg_time_phases); per phase: max and mean over the ranks.
    #      COLLECTIVE: every rank runs it (the cycles exchange halos), rank 0 reports (eager launches were confirmed on the
    #      hardware at 2 GPUs only, profiles/r02_scaling.md: hence the deadline) ----
    if world > 1 and not args.no_phases:
        with guard("phases_ms"):
            try:
                ph = mg.time_phases(level, nu1, nu2, gamma, 5)
            except Exception as ex:  # noqa: BLE001 - informational leg only
                ph = {"error": str(ex)}
            allp = [None] * world
            dist.all_gather_object(allp, ph)
            if all(p and "error" not in p for p in allp):
                names = [k for k in allp[0] if k != "ops_per_cycle"]
                tot = {k: [sum(p.get(k, {}).values()) for p in allp] for k in names}
                phases = {"how": "5 eager (un-captured) cycles, CUDA events around every op of the communication-avoiding plan; "
                                 "includes the launch gaps of eager launches",
                          "per_phase_max_over_ranks": {k: max(v) for k, v in tot.items()},
                          "per_phase_mean_over_ranks": {k: sum(v) / world for k, v in tot.items()},
                          "sum_of_phases_max_rank": max(sum(tot[k][r] for k in names) for r in range(world)),
                          "rank0_by_level": {k: allp[0][k] for k in names},
                          "captured_cycle_ms": isolated_ms}
            else:
                errs = [p.get("error") for p in allp if p and "error" in p]
                phases = {"unavailable": errs[0] if errs else "this schedule does not run the communication-avoiding plan"}
            line["phases_ms"] = phases

    # ---- N > 1: the weighted-Jacobi cycle on the same grid (round 1's scaling workload), timing only ----
    if world > 1 and smoother != "jacobi" and not args.no_extra:
        with guard("extra.jacobi_same_grid"):
            try:
                mg.close()
                mg = None
                mgj = mgb200.Multigrid(level, dtype=dtype, smoother="jacobi", device=local_rank, rank=rank, world=world,
                                       comm_id=new_comm_id(), agglomerate_level=args.aggl, **flags)
                mgj.force_synthetic(SEED)
                mgj.zero_u(level)
                mgj.time_cycle(level, nu1, nu2, gamma, W)
                mgj.time_cycle(level, nu1, nu2, gamma, K)
                tj = []
                for _ in range(7):
                    barrier()
                    tj.append(max_over_ranks(mgj.time_cycle(level, nu1, nu2, gamma, K)))
                msj = statistics.median(tj) / K
                extra = {"jacobi_same_grid": {"workload": workload_name(level, args.dtype, nu1, nu2, gamma, "jacobi") + f", row slabs over {world} GPUs",
                                              "ms_per_step": msj, "value": upd / (msj * 1e-3), "unit": UNIT}}
                mgj.close()
                if rank == 0 and not args.no_n1:
                    mg1 = mgb200.Multigrid(level, dtype=dtype, smoother="jacobi", device=local_rank, **flags)
                    mg1.force_synthetic(SEED)
                    mg1.zero_u(level)
                    mg1.time_cycle(level, nu1, nu2, gamma, W)
                    mg1.time_cycle(level, nu1, nu2, gamma, K)
                    ms1 = statistics.median([mg1.time_cycle(level, nu1, nu2, gamma, K) for _ in range(5)]) / K
                    mg1.close()
                    extra["jacobi_same_grid"].update({"n1_ms_per_step": ms1, "speedup": ms1 / msj, "efficiency": ms1 / msj / world})
                barrier()
            except Exception as ex:  # noqa: BLE001 - informational leg only
                extra = {"jacobi_same_grid": {"error": str(ex)}}
            line["extra"] = extra

    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_vcycle_rate(level, nu1, nu2, steps=3, warmup=1, smoother=1 if rb else 0, gamma=gamma,
                                               with_csr=True, with_as_written=True)
    if mg is not None:
        mg.close()
    guard.print_final()
    if dist is not None:
        with guard("teardown"):
            dist.barrier()
            dist.destroy_process_group()


def run_micro(args, rank, world, local_rank):
    """BASELINE.json configs[4]: smoother + residual sweep alone (HBM-roofline micro-benchmark) with temporal
    blocking depths k = 1, 2, 3, 4, on one GPU or on row slabs.  Resident data only (b = f h^2, u = 0), no host
    vectors; prints one JSON line: bytes are the algorithmic 3S per point per LAUNCH, so the
    "effective" rate of a k-sweep launch is k times its GB/s."""
    import torch
    import mgb200
    from mgb200 import capi
    torch.cuda.set_device(local_rank)
    dist, comm = None, None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        blob = [mgb200.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        comm = blob[0]
    level = args.level or 15
    dtype = np.float32 if args.dtype == "f32" else np.float64
    esize = np.dtype(dtype).itemsize
    n = (1 << level) - 1
    smoother = args.smoother or "jacobi"
    mg = mgb200.Multigrid(level, coarsest_level=max(1, level - 1), dtype=dtype, smoother=smoother, device=local_rank,
                          rank=rank, world=world, comm_id=comm, agglomerate_level=args.aggl or max(1, level - 1))
    mg.force_constant(4.0)
    mg.zero_u(level)
    rows = n if world == 1 else mg.info(capi.MG_INFO_ROW_END, level) - mg.info(capi.MG_INFO_ROW_BEGIN, level)
    pts_total = n * n
    peak, peak_src = measured_peak_gbs()
    reps = max(3, args.steps)
    out = {}
    ops = [("jacobi_k1", capi.MG_OP_SMOOTH1, 1, 3.0), ("jacobi_k2_one_launch", capi.MG_OP_SMOOTH2, 2, 3.0),
           ("jacobi_k3_one_launch", capi.MG_OP_SMOOTH3, 3, 3.0), ("jacobi_k4_one_launch", capi.MG_OP_SMOOTH4, 4, 3.0),
           ("residual", capi.MG_OP_RESIDUAL, 1, 3.0), ("residual_norm_only", capi.MG_OP_RESIDUAL_NORM, 1, 2.0)]
    for name, op, k, s_per_pt in ops:
        try:
            mg.time_op(op, level, max(1, args.warmup))
            if dist is not None:
                dist.barrier()
            ms = mg.time_op(op, level, reps) / reps
        except capi.MgError:
            continue
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        gbs = s_per_pt * esize * pts_total / (ms * 1e-3) / 1e9          # whole job (all ranks)
        out[name] = {"ms": ms, "sweeps_per_launch": k, "GBps": gbs, "frac_of_peak_per_gpu": gbs / world / peak,
                     "effective_GBps_per_sweep": gbs * k, "point_updates_per_s": k * pts_total / (ms * 1e-3)}
    best = max(out.items(), key=lambda kv: kv[1].get("point_updates_per_s", 0))
    line = {"metric": "smoother_point_updates_per_s", "value": best[1]["point_updates_per_s"], "unit": UNIT, "n_gpus": world,
            "steps": reps, "warmup": args.warmup, "ms_per_step": best[1]["ms"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"{n + 2}^2 {'fp32' if esize == 4 else 'fp64'} smoother + residual sweeps alone "
                                   f"(temporal blocking k=1..4), {smoother}" + ("" if world == 1 else f", row slabs over {world} GPUs"),
                       "level": level, "rows_per_rank": rows, "best": best[0]},
            "roofline": {"bound": "hbm", "achieved": best[1]["GBps"] / world, "peak": peak, "unit": "GB/s",
                         "frac": best[1]["GBps"] / world / peak, "traffic": None, "peak_source": peak_src, "kernels": out},
            "gpu_launches": int(mg.launches)}
    mg.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=0, help="finest level (default 12 at 1 GPU, 14 at N>1)")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--smoother", default=None, choices=["jacobi", "rbgs"],
                    help="default: jacobi at 1 GPU (BASELINE configs[1]), rbgs at N > 1 (configs[2])")
    ap.add_argument("--nu1", type=int, default=2)
    ap.add_argument("--nu2", type=int, default=2)
    ap.add_argument("--gamma", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-tail", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--aggl", type=int, default=0, help="agglomeration level for N>1 (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (tuning runs)")
    ap.add_argument("--no-n1", action="store_true", help="N>1: skip the single-GPU run of the same workload on rank 0 (strong_scaling, mgpu_parity)")
    ap.add_argument("--no-phases", action="store_true", help="N>1: skip the per-phase timing of the distributed cycle")
    ap.add_argument("--no-extra", action="store_true", help="N>1: skip the extra weighted-Jacobi timing on the same grid")
    ap.add_argument("--leg-timeout", type=float, default=150.0,
                    help="N>1: seconds an informational (collective) leg may take before the line is printed without it")
    ap.add_argument("--micro", action="store_true",
                    help="BASELINE configs[4]: smoother/residual micro-benchmark (default 32769^2; use --dtype f32)")
    ap.add_argument("--full-host-vectors", action="store_true",
                    help="N>1: every rank holds the full-grid host vectors (default: only the rows of its slab)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.micro:
        run_micro(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
