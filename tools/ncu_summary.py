"""Per-kernel launch counts / average durations from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(list)
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    try:
        agg[d['Kernel Name'][:64]].append(float(d['Metric Value'].replace(',', '')))
    except Exception:
        pass
tot = sum(sum(v) for v in agg.values())
print(f'{"kernel":64s} {"n":>5s} {"avg us":>9s} {"share":>6s}')
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f'{k:64s} {len(v):5d} {sum(v) / len(v) / 1e3:9.1f} {sum(v) / tot:6.3f}')
print(f'total {tot / 1e6:.2f} ms over {sum(len(v) for v in agg.values())} launches')
