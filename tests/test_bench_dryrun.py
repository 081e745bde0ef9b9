"""CPU dry run of bench.py's control flow with a stub context (no GPU here): catches Python
errors in the harness and checks the JSON contract (keys the driver reads).  The numbers are
fake; the real ones come from the B200 run."""
import argparse
import io
import json
import contextlib

import numpy as np
import pytest

import bench


class StubMG:
    """Mimics the slice of multigrid_nikhil_c-_b200.Multigrid that bench.py uses."""

    def __init__(self, level, dtype=np.float64, rank=0, world=1, **kw):
        self.level, self.dtype, self.rank, self.world = level, np.dtype(dtype), rank, world
        self._launches = 0
        n = (1 << level) - 1
        rows = (n + 1) // world
        self.own = (max(1, rank * rows), (rank + 1) * rows if rank < world - 1 else n + 1)

    def side(self, level):
        return (1 << level) - 1

    def slab_rows(self, level):
        return max(self.own[0] - 6, 1), min(self.own[1] + 6, self.side(level) + 1)

    def info(self, what, level=0):
        from mgb200 import capi
        return {capi.MG_INFO_ROW_BEGIN: self.own[0], capi.MG_INFO_ROW_END: self.own[1],
                capi.MG_INFO_AGGLOMERATE_LEVEL: 11}.get(what, 0)

    @property
    def launches(self):
        return self._launches

    def set_rhs(self, level, f):
        assert f.size == self.side(level) ** 2

    def set_rhs_slab(self, level, slab):
        ya, yb = self.slab_rows(level)
        assert slab.size == (yb - ya) * self.side(level)

    def zero_u(self, level):
        pass

    def force_constant(self, f):
        pass

    def force_synthetic(self, seed=1234):
        pass

    def checksum(self, level, which=0):
        return 0x1234 if self.world == 1 else 0x1234 // self.world + (0x1234 % self.world if self.rank == 0 else 0)

    def cycle(self, level=None, nu1=2, nu2=2, gamma=1):
        self._launches += 13

    def time_phases(self, level, nu1, nu2, gamma, reps):
        return {"ops_per_cycle": 3, "halo_exchange": {str(level): 0.03}, "pre": {str(level): 0.2}, "post": {str(level): 0.2}} if self.world > 1 else {}

    def cycles(self, count, level=None, nu1=2, nu2=2, gamma=1):
        self._launches += 13 * count

    def time_cycle(self, level, nu1, nu2, gamma, reps):
        self._launches += 13 * reps
        return 0.3 * reps * 4.0 ** (level - 12)

    def time_op(self, op, level, reps):
        from mgb200 import capi
        if op == capi.MG_OP_POST_FUSED:
            raise capi.MgError(3, "not available")
        return 0.08 * reps

    def vcyclemultigrid(self, u, f, nu1, nu2, gamma, inplace=False):
        assert inplace and u.size == f.size
        return u

    def fullmultigrid(self, f, cycles, nu1, nu2, out=None):
        return out if out is not None else np.zeros_like(f)

    def fmg(self, cycles, nu1, nu2):
        pass

    def solve(self, rtol, max_cycles, nu1, nu2, gamma):
        return 3, 1e-9, np.array([1.0, 1e-3, 1e-6, 1e-9])

    def sync(self):
        pass

    def vcyclemultigrid_slab(self, level, u, f, nu1, nu2, gamma):
        assert u.size == f.size

    def close(self):
        pass


def _args(**kw):
    d = dict(gpus=1, steps=3, warmup=1, impl="ours", level=6, dtype="f64", smoother="jacobi", nu1=2, nu2=2, gamma=1,
             no_graph=False, no_fused=False, no_tail=False, no_cpu=True, aggl=0, no_e2e=False, full_host_vectors=False, no_n1=False,
             no_extra=False, no_phases=False, micro=False, leg_timeout=120.0)
    d.update(kw)
    return argparse.Namespace(**d)


REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"]


@pytest.mark.parametrize("world", [1, 2])
def test_bench_control_flow_and_json_contract(monkeypatch, world):
    import torch
    import mgb200
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(mgb200, "Multigrid", StubMG)
    monkeypatch.setattr(mgb200, "comm_id", lambda: bytes(128))
    monkeypatch.setattr(bench, "pinned", lambda nelem, dtype, device=None: (None, np.empty(nelem, dtype=dtype)))
    monkeypatch.setattr(bench, "ClockSampler", lambda dev: type("S", (), {"start": lambda s: None, "stop": lambda s: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 1}})())
    if world > 1:
        import torch.distributed as dist
        monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
        monkeypatch.setattr(dist, "broadcast_object_list", lambda *a, **k: None)
        monkeypatch.setattr(dist, "all_gather_object", lambda lst, obj: lst.__setitem__(
            slice(None), [obj] * len(lst) if isinstance(obj, dict) else [obj] + [0] * (len(lst) - 1)))
        monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
        monkeypatch.setattr(dist, "all_reduce", lambda *a, **k: None)
        monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
        monkeypatch.setattr(torch, "tensor", lambda data, **k: torch.as_tensor(np.asarray(data, dtype=np.float64)))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_ours(_args(gpus=world, level=7 if world > 1 else 6), rank=0, world=world, local_rank=0)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["higher_is_better"] is True and d["n_gpus"] == world
    assert set(["bound", "achieved", "peak", "unit", "frac", "traffic"]) <= set(d["roofline"])
    assert set(["value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"]) <= set(d["e2e"])
    assert d["gpu_launches"] > 0 and "workload" in d["config"] and d["isolated_cycle_ms"] > 0
    if world > 1:
        ss = d["strong_scaling"]
        assert set(["n1_ms_per_step", "speedup", "efficiency", "mgpu_parity", "checksum_1gpu", "checksum_ngpu"]) <= set(ss)
        assert d["phases_ms"]["per_phase_max_over_ranks"]["pre"] == pytest.approx(0.2)
        assert d["config"]["smoother"] == "jacobi"     # explicit --smoother wins; the default at N > 1 is rbgs
        assert bench.default_workload(_args(level=0, smoother=None), 2) == (14, "rbgs")
        assert bench.default_workload(_args(level=0, smoother=None), 1) == (12, "jacobi")


def test_updates_per_cycle_matches_survey():
    assert bench.updates_per_cycle(12, 1, 2, 2) == 89413008
    assert bench.updates_per_cycle(8, 1, 2, 2, gamma=2) > bench.updates_per_cycle(8, 1, 2, 2)


def test_reference_arm_line(capsys):
    bench.run_reference(_args(impl="reference", level=6, steps=2, warmup=1), rank=0, world=1)
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    bench.run_reference(_args(impl="reference", level=6), rank=1, world=2)    # other ranks print nothing
    assert capsys.readouterr().out == ""
    # N > 1: the SAME workload string as our arm (the driver compares them), RB-GS, all host cores despite torchrun's
    # OMP_NUM_THREADS=1, right-hand side = the host restatement of the device generator
    import os
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        bench.run_reference(_args(impl="reference", level=6, smoother=None, steps=2, warmup=1), rank=0, world=2)
    finally:
        os.environ.pop("OMP_NUM_THREADS", None)
    d2 = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d2["config"]["workload"] == bench.workload_name(6, "f64", 2, 2, 1, "rbgs")
    assert d2["cpu_baseline"]["cores"] == bench.host_threads()


def test_device_synthetic_rows_match_the_test_restatement():
    import synth_ref
    for level, dtype in ((5, np.float64), (7, np.float32)):
        n = (1 << level) - 1
        out = np.empty(n * n, dtype=dtype)
        bench.device_synthetic_rows(level, 1, n + 1, dtype, out)
        assert np.array_equal(out, synth_ref.synthetic_rhs(level, 1234, dtype))
        part = np.empty(5 * n, dtype=dtype)
        bench.device_synthetic_rows(level, 4, 9, dtype, part)
        assert np.array_equal(part, out.reshape(n, n)[3:8].reshape(-1))


def test_micro_benchmark_line(monkeypatch, capsys):
    import torch
    import mgb200
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(mgb200, "Multigrid", lambda level, **kw: StubMG(level, dtype=kw.get("dtype", np.float64)))
    bench.run_micro(_args(micro=True, level=9, dtype="f32", steps=3), rank=0, world=1, local_rank=0)
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["metric"] == "smoother_point_updates_per_s" and d["dtype"] == "f32"
    assert "jacobi_k2_one_launch" in d["roofline"]["kernels"] and d["roofline"]["frac"] > 0


def test_gpu_local_affinity_binds_and_restores(monkeypatch):
    """bench.GpuLocalAffinity: bound to the GPU's CPUs inside, previous mask restored outside, silent without NVML."""
    import os
    import sys
    import types
    allowed = sorted(os.sched_getaffinity(0))
    if len(allowed) < 2:
        pytest.skip("needs two CPUs")
    fake = types.ModuleType("pynvml")
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByIndex = lambda i: i
    fake.nvmlDeviceGetCpuAffinity = lambda h, n: [1 << (allowed[0] % 64) if i == allowed[0] // 64 else 0 for i in range(n)]
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    before = os.sched_getaffinity(0)
    with bench.GpuLocalAffinity(0) as a:
        assert os.sched_getaffinity(0) == {allowed[0]} and a.cpus == 1
    assert os.sched_getaffinity(0) == before
    fake.nvmlDeviceGetCpuAffinity = lambda h, n: (_ for _ in ()).throw(RuntimeError("no NVML"))
    with bench.GpuLocalAffinity(0) as a:
        assert os.sched_getaffinity(0) == before and a.cpus is None


def test_leg_guard_prints_the_core_line_when_a_collective_leg_stalls():
    """bench.LegGuard: a leg that does not come back within its budget costs that leg, not the record -- rank 0 prints
    the line as it stands with `aborted_leg`, every rank exits 0; a finished leg disarms its deadline."""
    import subprocess
    import sys
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = ("import sys, time, json; sys.path.insert(0, %r); import bench\n"
            "line = {'metric': bench.METRIC, 'value': 1.0}\n"
            "g = bench.LegGuard(int(sys.argv[1]), line, enabled=True, budget_s=0.3)\n"
            "with g('quick'):\n    line['quick'] = True\n"
            "time.sleep(0.5)\n"                     # the disarmed deadline of 'quick' must not fire here
            "with g('stalled'):\n    time.sleep(30)\n"
            "print('NOT REACHED')\n") % root
    for rank in (0, 1):
        p = subprocess.run([sys.executable, "-c", prog, str(rank)], capture_output=True, text=True, timeout=25)
        assert p.returncode == 0, p.stderr
        assert "NOT REACHED" not in p.stdout
        if rank == 0:
            d = json.loads(p.stdout.strip().splitlines()[-1])
            assert d["value"] == 1.0 and d["quick"] is True and d["aborted_leg"]["leg"] == "stalled"
        else:
            assert p.stdout.strip() == ""
    # not enabled (N = 1): no deadline, one final line, print_final only once
    line = {"metric": bench.METRIC}
    g = bench.LegGuard(0, line, enabled=False, budget_s=0.01)
    with g("leg"):
        import time
        time.sleep(0.05)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        g.print_final()
        g.print_final()
    assert len(buf.getvalue().strip().splitlines()) == 1 and "aborted_leg" not in line


def test_leg_guard_records_an_exception_and_goes_on():
    line = {}
    g = bench.LegGuard(0, line, enabled=True, budget_s=30.0)
    with g("broken"):
        raise RuntimeError("boom")
    with g("fine"):
        line["fine"] = 1
    assert line["leg_errors"] == {"broken": "RuntimeError: boom"} and line["fine"] == 1
    with pytest.raises(KeyboardInterrupt):
        with g("interrupted"):
            raise KeyboardInterrupt()
