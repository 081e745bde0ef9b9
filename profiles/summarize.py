"""Summarise ncu captures brought back in gpurun_out/ into small tracked text files.
usage: python profiles/summarize.py launches <launches.csv> <out.txt> [first_id count]
       python profiles/summarize.py full <prof.ncu-rep> <out.txt>
"""
import collections
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'dram__bytes_write.sum.per_second', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg.per_second',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def launches(path, out, first=None, count=None):
    lines = [l for l in open(path) if not l.startswith('==')]
    data = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        t = float(row['Metric Value'].replace(',', ''))
        t = {'ns': t / 1000, 'us': t, 'ms': t * 1000, 'usecond': t, 'nsecond': t / 1000, 'msecond': t * 1000}[row['Metric Unit']]
        data.append((int(row['ID']), row['Kernel Name'].split('(')[0].replace('void ', ''), row['Grid Size'], row['Block Size'], t))
    if first is not None:
        data = [d for d in data if first <= d[0] < first + count]
    agg = collections.OrderedDict()
    for d in data:
        k = (d[1], d[2])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += d[4]
    tot = sum(d[4] for d in data)
    with open(out, 'w') as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {len(data)} launches, total {tot:.1f} us\n")
        f.write("# (per-launch times under ncu are cold-cache and serialised: compare SHARES)\n")
        f.write(f"{'kernel':34s} {'grid':>16s} {'n':>4s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}\n")
        for (name, grid), (n, t) in agg.items():
            f.write(f"{name[:34]:34s} {grid:>16s} {n:4d} {t:10.2f} {t / n:8.2f} {100 * t / tot:5.1f}%\n")
    print(open(out).read())


def full(path, out):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units = r[0], r[1]
    with open(out, 'w') as f:
        f.write(f"# ncu --set full --clock-control none: {path}\n")
        for row in r[2:]:
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"{w} = {row[i]} {units[i]}\n")
            f.write('---\n')
    print(open(out).read())


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3], *(int(a) for a in sys.argv[4:6]))
    else:
        full(sys.argv[2], sys.argv[3])
