#!/usr/bin/env python
"""bench.py — V-cycle throughput of the B200-native multigrid path (BASELINE.json metric:
"V-cycle ms & grid-pt updates/s at 4097^2 fp64; smoother HBM GB/s vs B200 peak").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one V(2,2) cycle (weighted Jacobi w=2/3, full weighting, bilinear
prolongation, coarsened to 3x3) over one synthetic right-hand side.
  N = 1 : configs[1], 4097^2 fp64 on one B200.
  N > 1 : the same cycle on the row-slab decomposed 16385^2 grid (configs[2] geometry,
          Jacobi smoother so the metric stays comparable across N), one process per GPU
          under torchrun, halo exchange over NVLink; coarse levels agglomerated.
`value` = grid-point updates per second of the whole job, inputs resident in HBM, timed
with CUDA events on the library's stream (max over ranks).  `e2e` = the same metric
through the reference-shaped host call mg_host_vcyclemultigrid (host vectors in, host
vector out: H2D of vec_h and f_h and D2H of the result inside the timed region).
`--impl reference` times the CPU oracle port of the reference's algorithm with all host
threads on the same config (the reference itself needs DPC++/oneMKL and cannot be built
here).
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


def synthetic_rhs(level: int, dtype) -> np.ndarray:
    """SURVEY 8d input (ii): b = h^2 * U(-1,1), numpy default_rng(1234), row-major interior order."""
    n = (1 << level) - 1
    h = 1.0 / (1 << level)
    return (h * h * np.random.default_rng(1234).uniform(-1.0, 1.0, n * n)).astype(dtype)


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


def cpu_vcycle_rate(level, nu1, nu2, steps, warmup, smoother=0, with_csr=False):
    import oracle
    o, build = load_oracle()
    nt = o.max_threads()
    p = oracle.Params(nu1=nu1, nu2=nu2, smoother=smoother, nthreads=nt)
    b = synthetic_rhs(level, np.float64)
    u = np.zeros_like(b)
    for _ in range(warmup):
        u = o.vcyclemultigrid(u, b, p, inplace=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        u = o.vcyclemultigrid(u, b, p, inplace=True)
    dt = (time.perf_counter() - t0) / steps
    upd = updates_per_cycle(level, 1, nu1, nu2)
    out = {"value": upd / dt, "unit": UNIT, "cores": nt, "kind": "port", "ms_per_step": dt * 1e3,
           "sample": f"{steps} V({nu1},{nu2}) cycles at {(1 << level) + 1}^2 fp64 after {warmup} warm-up, "
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
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    level = args.level or (12 if world == 1 else 14)
    if world > 1:
        level = min(level, 13)  # bounded sample: a 16385^2 oracle cycle takes ~4x longer per step
    r = cpu_vcycle_rate(level, args.nu1, args.nu2, args.steps, args.warmup)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{(1 << level) + 1}^2 fp64 V({args.nu1},{args.nu2}) weighted Jacobi, FW/bilinear, CPU oracle port"},
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
        blob = [mgb200.comm_id() if rank == 0 else None]
        dist.broadcast_object_list(blob, src=0)
        comm = blob[0]

    level = args.level or (12 if world == 1 else 14)
    dtype = np.float64 if args.dtype == "f64" else np.float32
    nu1, nu2, gamma = args.nu1, args.nu2, args.gamma
    n = (1 << level) - 1
    esize = np.dtype(dtype).itemsize

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

    mg = mgb200.Multigrid(level, dtype=dtype, smoother=args.smoother, device=local_rank, rank=rank, world=world,
                          comm_id=comm, graph=not args.no_graph, fused=not args.no_fused,
                          coarse_tail=not args.no_tail, agglomerate_level=args.aggl)
    slab = world > 1 and not args.full_host_vectors
    if slab:
        # a rank only ever touches the interior rows it stores: keep just those on the host (Multigrid.set_rhs_slab)
        ya, yb = mg.slab_rows(level)
        f_t, f_host = pinned((yb - ya) * n, dtype, local_rank)   # on the GPU's NUMA node
        u_t, u_host = pinned((yb - ya) * n, dtype, local_rank)
        # one generator per GLOBAL row: every rank holds the same values for a row it stores (its halo rows are the
        # neighbour's owned rows), and the right-hand side does not depend on the number of ranks
        h = 1.0 / (1 << level)
        fv = f_host.reshape(yb - ya, n)
        for i, row in enumerate(range(ya, yb)):
            fv[i] = (h * h * np.random.default_rng([1234, row]).uniform(-1.0, 1.0, n)).astype(dtype)
        u_host[:] = 0
        mg.set_rhs_slab(level, f_host)
    else:
        f_t, f_host = pinned(n * n, dtype, local_rank)           # on the GPU's NUMA node
        u_t, u_host = pinned(n * n, dtype, local_rank)
        f_host[:] = synthetic_rhs(level, dtype)
        u_host[:] = 0
        mg.set_rhs(level, f_host)
    mg.zero_u(level)
    upd = updates_per_cycle(level, 1, nu1, nu2, gamma)

    # ---- resident-data timing: exactly K cycles per region, several regions so that the
    #      clock sampler sees the load; the median region is reported ----
    K, W = args.steps, max(args.warmup, 3)
    mg.time_cycle(level, nu1, nu2, gamma, W)
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

    # ---- per-kernel rooflines on the finest level (CUDA events on the library's stream) ----
    peak, peak_src = measured_peak_gbs()
    pts = n * n if world == 1 else (mg.info(capi.MG_INFO_ROW_END, level) - mg.info(capi.MG_INFO_ROW_BEGIN, level)) * n
    kernels = {}
    reps = 20
    algo = {"jacobi_sweep": (capi.MG_OP_SMOOTH1, 3.0), "two_sweeps_one_launch": (capi.MG_OP_SMOOTH2, 3.0), "residual": (capi.MG_OP_RESIDUAL, 3.0),
            "residual_norm_only": (capi.MG_OP_RESIDUAL_NORM, 2.0), "restrict": (capi.MG_OP_RESTRICT, 1.25),
            "prolong_correct": (capi.MG_OP_PROLONG, 2.25)}
    for name, (op, s_per_pt) in algo.items():
        if world > 1 and name in ("restrict", "prolong_correct"):
            continue
        try:
            t = mg.time_op(op, level, reps) / reps
        except capi.MgError:
            continue
        gbs = s_per_pt * esize * pts / (t * 1e-3) / 1e9
        kernels[name] = {"ms": t, "algorithmic_bytes": s_per_pt * esize * pts, "GBps": gbs, "frac_of_peak": gbs / peak}
    for name, op, s_per_pt in (("pre_fused(2 sweeps+residual+restrict)", capi.MG_OP_PRE_FUSED, 3.25),
                               ("post_fused(prolong+correct+2 sweeps)", capi.MG_OP_POST_FUSED, 3.25),
                               ("postpre_chain(prolong+correct+4 sweeps+residual+restrict)", capi.MG_OP_POSTPRE_FUSED, 3.5)):
        try:
            t = mg.time_op(op, level, reps) / reps
        except capi.MgError:
            continue
        unfused = (2.25 + 4 * 3 + 3 + 1.25) if "chain" in name else ((2 * 3 + 3 + 1.25) if "pre" in name else (2.25 + 2 * 3))
        kernels[name] = {"ms": t, "algorithmic_bytes": s_per_pt * esize * pts,
                         "GBps": s_per_pt * esize * pts / (t * 1e-3) / 1e9,
                         "frac_of_peak": s_per_pt * esize * pts / (t * 1e-3) / 1e9 / peak,
                         "effective_unfused_GBps": unfused * esize * pts / (t * 1e-3) / 1e9}
    # cumulative cost of the cycle from each level down (level_ms[l] - level_ms[l-1] = cost of level l's visit)
    level_ms = {}
    if world == 1:
        for l in range(max(2, min(6, level)), level + 1):
            try:
                mg.time_cycle(l, nu1, nu2, gamma, 3)
                level_ms[str(l)] = mg.time_cycle(l, nu1, nu2, gamma, 20) / 20
            except capi.MgError:
                pass
    # dominant kernel of the timed region: the fused PRE kernel (2 sweeps + residual + restriction) on the finest
    # level when MG_FUSED is on (largest single share of the cycle, profiles/*_launches_one_vcycle.txt), else the
    # Jacobi sweep.  The plain smoother numbers the BASELINE metric asks for are kept under "smoother".
    pre_key = "pre_fused(2 sweeps+residual+restrict)"
    dom = pre_key if (pre_key in kernels and not args.no_fused) else "jacobi_sweep"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and level == 12 and dtype == np.float64:
        try:
            traffic = json.load(open(tpath)).get("pre_fused" if dom == pre_key else "jacobi_sweep")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm",
                "kernel": ("k_stream<T,2,PRE> (2 Jacobi sweeps + residual + full weighting, finest level)" if dom == pre_key
                           else "k_jacobi (one weighted-Jacobi sweep, finest level)"),
                "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["GBps"] / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes"],
                "ms_per_launch": kernels[dom]["ms"],
                "smoother": {k: kernels[k] for k in ("jacobi_sweep", "two_sweeps_one_launch") if k in kernels},
                "kernels": kernels, "cycle_ms_from_level_down": level_ms}

    # ---- end to end through the reference-shaped host call (P:575 on host vectors) ----
    def e2e_call():
        if slab:
            mg.vcyclemultigrid_slab(level, u_host, f_host, nu1, nu2, gamma)
        else:
            mg.vcyclemultigrid(u_host, f_host, nu1, nu2, gamma, inplace=True)

    e2e_steps = max(3, min(K, 10)) if not args.no_e2e else 1
    u_host[:] = 0
    for _ in range(2):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    rows = n if world == 1 else (mg.info(capi.MG_INFO_ROW_END, level) - mg.info(capi.MG_INFO_ROW_BEGIN, level))
    e2e = {"value": upd / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
           "h2d_bytes_per_step": 2 * (n if world == 1 else mg.slab_rows(level)[1] - mg.slab_rows(level)[0]) * n * esize,
           "d2h_bytes_per_step": rows * n * esize,
           "call": "mg_host_vcyclemultigrid (vcyclemultigrid P:575 on pinned host vectors)",
           "host_buffers_on_gpu_numa_node_cpus": NUMA_CPUS}

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

    # tolerance-controlled solve on the same right-hand side (the reference runs a fixed number of cycles
    # and prints only the vector length): cycle count, residual history, wall time incl. the per-cycle norm read-back
    solve_info = None
    try:
        mg.zero_u(level)
        mg.sync()
        barrier()
        t0 = time.perf_counter()
        k_cyc, relres, hist = mg.solve(1e-8, 40, nu1, nu2, gamma)
        mg.sync()
        solve_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        solve_info = {"rtol": 1e-8, "cycles": k_cyc, "relres": relres, "ms": solve_ms,
                      "residual_history": [float(h) for h in hist],
                      "factors": [float(hist[i + 1] / hist[i]) for i in range(len(hist) - 1) if hist[i] > 0]}
    except Exception as ex:  # noqa: BLE001 - informational leg only
        solve_info = {"error": str(ex)}

    # N > 1: the same workload on ONE GPU (rank 0 alone, resident data, same flags), so that the strong-scaling
    # denominator for this grid size is in the same line (bench.py --gpus 1 measures BASELINE configs[1], 4097^2)
    n1 = None
    if world > 1 and not args.no_n1:
        if rank == 0:
            try:
                mg1 = mgb200.Multigrid(level, dtype=dtype, smoother=args.smoother, device=local_rank, graph=not args.no_graph,
                                       fused=not args.no_fused, coarse_tail=not args.no_tail)
                mg1.force_constant(4.0)
                mg1.zero_u(level)
                mg1.time_cycle(level, nu1, nu2, gamma, W)
                ms1 = mg1.time_cycle(level, nu1, nu2, gamma, K) / K
                mg1.close()
                n1 = {"workload": f"{n + 2}^2, same cycle on 1 GPU (rank 0 alone, b = 4h^2)", "ms_per_step": ms1,
                      "value": upd / (ms1 * 1e-3), "unit": UNIT}
            except Exception as ex:  # noqa: BLE001 - informational leg only
                n1 = {"error": str(ex)}
        barrier()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if dtype == np.float64 else "f32", "data": "synthetic",
            "config": {"workload": f"{n + 2}^2 {'fp64' if esize == 8 else 'fp32'} V({nu1},{nu2}) gamma={gamma} "
                                   f"{args.smoother}, full weighting / bilinear, coarsened to 3x3"
                                   + ("" if world == 1 else f", row slabs over {world} GPUs"),
                       "level": level, "rhs": "h^2*U(-1,1) rng(1234)" if not slab else "h^2*U(-1,1), rng seeded per global row", "updates_per_cycle": upd,
                       "l2_policy": "inputs larger than L2 (4 arrays x %.0f MB on the finest level)" % (n * n * esize / 1e6),
                       "regions": regions, "region_stat": "median", "agglomerate_level": mg.info(capi.MG_INFO_AGGLOMERATE_LEVEL, level) if world > 1 else None,
                       "timed_as": f"{K} consecutive cycles per region through mg_time_cycle (mg_cycles)",
                       "env_knobs": {k: v for k, v in sorted(os.environ.items()) if k.startswith("MGB200_")},
                       "flags": {"visit_chain": os.environ.get("MGB200_CHAIN") == "1", "graph": not args.no_graph,
                                                                              "fused": not args.no_fused,
                                                                              "coarse_tail": not args.no_tail}},
            "finest_points_per_s": n * n / (ms_step * 1e-3),
            "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches * 1), "clocks": clocks}
    if os.environ.get("MGB200_CHAIN") == "1":
        # with visit chains the K timed cycles share POST+PRE launches on the finest level; also report one isolated cycle
        line["isolated_cycle_ms"] = statistics.median([mg.time_cycle(level, nu1, nu2, gamma, 1) for _ in range(20)])
    line["solve"] = solve_info
    if n1 is not None:
        line["n1_same_workload"] = n1
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_vcycle_rate(level, nu1, nu2, steps=3, warmup=1, with_csr=True)
    mg.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
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
    mg = mgb200.Multigrid(level, coarsest_level=max(1, level - 1), dtype=dtype, smoother=args.smoother, device=local_rank,
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
                                   f"(temporal blocking k=1..4), {args.smoother}" + ("" if world == 1 else f", row slabs over {world} GPUs"),
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
    ap.add_argument("--smoother", default="jacobi", choices=["jacobi", "rbgs"])
    ap.add_argument("--nu1", type=int, default=2)
    ap.add_argument("--nu2", type=int, default=2)
    ap.add_argument("--gamma", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--no-tail", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--aggl", type=int, default=0, help="agglomeration level for N>1 (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (tuning runs)")
    ap.add_argument("--no-n1", action="store_true", help="N>1: skip the single-GPU run of the same workload on rank 0")
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
