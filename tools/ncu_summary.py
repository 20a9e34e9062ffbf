#!/usr/bin/env python
"""Summarise gpurun_out/launches.csv (+ an optional .ncu-rep) into profiles/<tag>_*.{csv,txt}.

usage: python tools/ncu_summary.py <tag> [gpurun_out/prof.ncu-rep]
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def launches(tag):
    src = ROOT / 'gpurun_out' / 'launches.csv'
    if not src.exists():
        return
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get('Metric Name') == 'gpu__time_duration.sum':
                a = agg.setdefault(d['Kernel Name'][:110], [0, 0.0, d['Grid Size'], d['Block Size']])
                a[0] += 1
                a[1] += float(d['Metric Value'].replace(',', ''))
    tot = sum(a[1] for a in agg.values()) or 1.0
    out = ROOT / 'profiles' / f'{tag}_launches.csv'
    with open(out, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n')
        f.write('kernel,launches,total_ns,avg_ns,share,grid,block\n')
        for k, a in agg.items():
            f.write(f'"{k}",{a[0]},{a[1]:.0f},{a[1] / a[0]:.0f},{a[1] / tot:.4f},"{a[2]}","{a[3]}"\n')
    print(out.read_text())


def full(tag, rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    out = ROOT / 'profiles' / f'{tag}_ncu_full.txt'
    with open(out, 'w') as f:
        f.write(f'# ncu --set full --clock-control none --import-source on; source: {Path(rep).name}\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d.get('Kernel Name', '')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f'{k:90s} {d[k]:>18s} {units[hdr.index(k)]}\n')
    print(out.read_text())


if __name__ == '__main__':
    tag = sys.argv[1]
    (ROOT / 'profiles').mkdir(exist_ok=True)
    launches(tag)
    if len(sys.argv) > 2:
        full(tag, sys.argv[2])
