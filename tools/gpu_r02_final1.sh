#!/bin/bash
# FINAL 1-GPU call of round 2: the whole GPU suite, the bench lines of BASELINE configs 2-5, ncu evidence per kernel and per
# level (reports are summarised to text ON the box and deleted: gpurun_out/ is limited to 64 MiB).
set -u
mkdir -p gpurun_out; O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; tail -1 $O/r02_final_smoke.log
timeout 900 python -m pytest tests -m gpu -q --durations=6 > $O/r02_final_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02_final_pytest_gpu.log; tail -12 $O/r02_final_pytest_gpu.log
timeout 400 python bench.py > $O/r02_final_bench_cfg2_4097_jacobi.json 2> $O/r02_final_bench_cfg2.err; tail -c 400 $O/r02_final_bench_cfg2.err
timeout 300 python bench.py --level 14 --smoother rbgs --no-cpu --no-e2e > $O/r02_final_bench_cfg3_16385_rbgs_n1.json 2> $O/r02_final_bench_cfg3.err
timeout 300 python bench.py --level 13 --gamma 2 --no-cpu --no-e2e > $O/r02_final_bench_cfg4_8193_W.json 2> $O/r02_final_bench_cfg4.err
timeout 300 python bench.py --micro --dtype f32 --level 15 > $O/r02_final_bench_cfg5_32769_f32_micro.json 2> $O/r02_final_bench_cfg5.err
python - <<'PY' > gpurun_out/r02_final_fmg_8193.txt 2>&1
import sys, time, statistics
sys.path.insert(0, '.')
import mgb200
with mgb200.Multigrid(13) as mg:      # BASELINE config 4, second half: full multigrid at 8193^2 (1 V(2,2) per level), resident data
    mg.force_synthetic(1234)
    mg.fmg(1, 2, 2); mg.sync()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); mg.fmg(1, 2, 2); mg.sync(); ts.append((time.perf_counter() - t0) * 1e3)
    r0 = None
    print(f"fullmultigrid 8193^2 fp64, 1 V(2,2) per level, resident: {statistics.median(ts):.3f} ms (min {min(ts):.3f})")
    mg.zero_u(13); k, rel, h = mg.solve(1e-8, 40, 2, 2, 2)
    print(f"W(2,2) solve to 1e-8: {k} cycles, relres {rel:.3e}, factors {[round(h[i+1]/h[i], 4) for i in range(k)]}")
PY
cat gpurun_out/r02_final_fmg_8193.txt
python - <<'PY'
import json
for tag in ("cfg2_4097_jacobi", "cfg3_16385_rbgs_n1", "cfg4_8193_W", "cfg5_32769_f32_micro"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r02_final_bench_{tag}.json") if l.startswith("{")][-1])
        k = d["roofline"]["kernels"]
        print(f"{tag:24s} ms {d['ms_per_step']:.4f} iso {d.get('isolated_cycle_ms', 0):.4f} launches {d['gpu_launches']} e2e {d.get('e2e', {}).get('ms_per_step')} pageable {d.get('e2e', {}).get('pageable', {}).get('ms_per_step')}")
        print("    ", " ".join(f"{n[:10]}:{v['ms']*1e3:.0f}us/{v.get('frac_of_peak', v.get('frac_of_peak_per_gpu', 0)):.2f}" for n, v in k.items()))
        if d.get("solve"): print("     solve", d["solve"].get("cycles"), d["solve"].get("ms"), d["solve"].get("overhead_vs_isolated_cycle"))
        if d.get("cpu_baseline"): print("     cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], (d["cpu_baseline"].get("reference_as_written") or {}).get("ms_per_step"))
    except Exception as ex:
        print(tag, "FAILED", ex)
PY
# ---- ncu: launch list of the default bench command (times only), then --set full per kernel and per level ----
CMD="python bench.py --no-cpu --no-e2e --steps 2 --warmup 3"
$CMD > $O/r02_final_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -c 700 --csv --log-file $O/r02_final_launches.csv $CMD > $O/r02_final_ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --kernel-name-base demangled"
cap() { tag=$1; kre=$2; n=$3; shift 3; python tools/profile_ops.py "$@" > $O/r02_final_$tag.plain.log 2>&1 && timeout 600 $NCU -k regex:"$kre" -c $n -o /tmp/r02_$tag python tools/profile_ops.py "$@" > $O/r02_final_$tag.ncu.log 2>&1 && python profiles/summarize.py full /tmp/r02_$tag.ncu-rep $O/r02_final_ncu_$tag.txt > /dev/null 2>&1; rm -f /tmp/r02_$tag.ncu-rep; ls -la $O/r02_final_ncu_$tag.txt; }
cap jacobi_L12 "k_stream" 9 12 jacobi f64 pre post chain
cap jacobi_L11 "k_stream" 9 11 jacobi f64 pre post chain
cap jacobi_L10 "k_stream" 9 10 jacobi f64 pre post chain
cap jacobi_L8 "k_stream" 9 8 jacobi f64 pre post chain
cap rbgs_L12 "k_stream" 6 12 rbgs f64 pre post
cap rbgs_L14 "k_stream" 6 14 rbgs f64 pre post
cap unfused_L12 "k_jacobi|k_residual|k_restrict|k_prolong" 15 12 jacobi f64 sweep residual norm restrict prolong
cap tail_and_cycle "k_tail|k_stream" 30 12 jacobi f64 cycle
du -sh $O
