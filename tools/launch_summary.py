"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean and minimum duration per kernel.
usage: python tools/launch_summary.py gpurun_out/launches.csv [--md]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
d = collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        rec = dict(zip(hdr, r))
        if rec.get('Metric Name') == 'gpu__time_duration.sum':
            name = re.sub(r'\(.*', '', rec['Kernel Name'])
            unit = rec['Metric Unit']
            v = float(rec['Metric Value'].replace(',', ''))
            v *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1., 'usecond': 1., 'ms': 1e3, 'msecond': 1e3}.get(unit, 1.)
            d.setdefault(name, []).append(v)
tot = sum(sum(v) for v in d.values())
md = '--md' in sys.argv
if md:
    print("| kernel | launches | mean us | min us | share |\n|---|---|---|---|---|")
for k, v in d.items():
    if md:
        print(f"| `{k[:80]}` | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    else:
        print(f"{k[:80]:80s} n={len(v):5d} mean={sum(v) / len(v):9.1f} us  min={min(v):9.1f}  share={100 * sum(v) / tot:5.1f}%")
