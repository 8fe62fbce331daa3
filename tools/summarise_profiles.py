"""Turn ncu CSV exports into the small tables committed under profiles/.

  python tools/summarise_profiles.py launches <launch-list.csv>        > profiles/..._launch_table.md
  python tools/summarise_profiles.py full <ncu --page raw --csv file>  > profiles/..._ncu_full_summary.md
"""
import collections
import csv
import re
import sys

FULL_METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs), CTAs/SM'),
    ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem), CTAs/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_scoreboard / issue'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier / issue'),
    ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg_throttle / issue'),
    ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio_throttle / issue'),
    ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_selected / issue'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait / issue'),
]


def launches(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').strip()
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1000 if u in ('ns', 'nsecond') else (v * 1000 if u in ('ms', 'msecond') else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    print(f'{n} launches, {tot / 1e3:.2f} ms of kernel time (ncu: serialised, caches flushed between kernels)\n')
    print('| kernel | launches | total us | share | avg us |')
    print('|---|---:|---:|---:|---:|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k[:80]}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f} % | {v[1] / v[0]:.1f} |')


def full(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print('| kernel | grid | block | ' + ' | '.join(lbl for _, lbl in FULL_METRICS) + ' |')
    print('|---|---|---|' + '---:|' * len(FULL_METRICS))
    for r in rows[2:]:
        name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').strip()
        cells = []
        for m, _ in FULL_METRICS:
            if m in col:
                try:
                    cells.append(f'{float(r[col[m]].replace(",", "")):.4g} {units[col[m]]}'.strip())
                except ValueError:
                    cells.append(r[col[m]])
            else:
                cells.append('-')
        print(f'| `{name}` | {r[col["Grid Size"]]} | {r[col["Block Size"]]} | ' + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2])
