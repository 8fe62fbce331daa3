"""Summarise an `ncu --set full` report: per kernel (mean over its captured launches) duration, DRAM bytes read + written,
tensor-pipe and DRAM utilisation, occupancy, registers.  Writes a markdown table and, with --json, the per-launch DRAM
traffic map bench.py reads (profiles/r02_ncu_traffic.json).

  python tools/ncu_traffic.py gpurun_out/prof.ncu-rep --md profiles/r02_ncu_full_fused.md --json profiles/r02_ncu_traffic.json
"""
import argparse
import collections
import csv
import io
import json
import subprocess

M = {
    'dur_us': 'gpu__time_duration.sum',
    'dram_rd': 'dram__bytes_read.sum',
    'dram_wr': 'dram__bytes_write.sum',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'warps_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'smem_dyn': 'launch__shared_mem_per_block_dynamic',
    'l2_hit': 'lts__t_sector_hit_rate.pct',
}


def to_float(v, unit):
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return None
    u = (unit or '').lower().split('/')[0]
    scale = {'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'byte': 1.0, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}
    return x * scale.get(u, 1.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--md', default='')
    ap.add_argument('--json', default='')
    a = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', a.rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = r[ix['Kernel Name']].split('(')[0]
        d = agg.setdefault(name, collections.defaultdict(list))
        for k, m in M.items():
            if m in ix:
                v = to_float(r[ix[m]], units[ix[m]])
                if v is not None:
                    d[k].append(v)
    lines = ['| kernel | launches | duration us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | warps active % | regs | grid x block | dyn smem KB |',
             '|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|']
    traffic = {}
    for name, d in agg.items():
        mean = lambda k: sum(d[k]) / len(d[k]) if d[k] else float('nan')
        traffic[name] = mean('dram_rd') + mean('dram_wr')
        lines.append(f"| `{name}` | {len(d['dur_us'])} | {mean('dur_us'):.1f} | {mean('dram_rd') / 1e6:.2f} | {mean('dram_wr') / 1e6:.2f} | "
                     f"{mean('dram_pct'):.1f} | {mean('tensor_pct'):.2f} | {mean('warps_pct'):.1f} | {mean('regs'):.0f} | "
                     f"{mean('grid'):.0f} x {mean('block'):.0f} | {mean('smem_dyn') / 1e3:.1f} |")
    text = '\n'.join(lines)
    print(text)
    if a.md:
        with open(a.md, 'w') as f:
            f.write(f'ncu --set full --clock-control none (cold caches: ncu flushes between replays), report `{a.rep}`\n\n' + text + '\n')
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(traffic, f, indent=1)


if __name__ == '__main__':
    main()
