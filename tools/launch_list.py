#!/usr/bin/env python
"""Per-launch table of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
d = collections.OrderedDict()
for r in rows[1:]:
    k = (int(r[ix['ID']]), r[ix['Kernel Name']].split('(')[0][-44:], r[ix['Grid Size']])
    d.setdefault(k, {})[r[ix['Metric Name']]] = (float(r[ix['Metric Value']].replace(',', '')), r[ix['Metric Unit']])
ms = lambda t: t[0] / 1e6 if t[1] == 'ns' else (t[0] / 1e3 if t[1] == 'us' else t[0])
gb = lambda x: x[0] * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1}[x[1]]
tot = 0.0
for k, v in d.items():
    t = ms(v['gpu__time_duration.sum'])
    b = gb(v['dram__bytes_read.sum']) + gb(v['dram__bytes_write.sum']) if 'dram__bytes_read.sum' in v else 0.0
    tot += t
    print("%4d %-44s %-14s %9.3f ms %7.2f GB %6.0f GB/s" % (k[0], k[1], k[2], t, b, b / t * 1e3 if t else 0))
print("total %.3f ms" % tot)
