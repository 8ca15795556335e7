"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: launches, total time, share.
usage: python tools/launch_shares.py gpurun_out/launches_step.csv [skip_launches] [take_launches]"""
import csv, re, sys
from collections import OrderedDict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    v = float(r[vi].replace(',', ''))
    u = r[ui]
    us = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    rows.append((r[ki], us))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
take = int(sys.argv[3]) if len(sys.argv) > 3 else len(rows)
rows = rows[skip:skip + take]
agg = OrderedDict()
for k, us in rows:
    k = re.sub(r'\(.*', '', k)
    k = re.sub(r'^void ', '', k).replace('<unnamed>::', '')
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print('launches %d, total %.3f ms (cold-cache, serialised under ncu: shares only)' % (len(rows), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print('%6.2f%%  %9.3f ms  %4d x  %s' % (100 * a[1] / tot, a[1] / 1e3, a[0], k[:110]))
