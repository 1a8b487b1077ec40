"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/profile_step.py: per-kernel totals of
the LAST step and, with the op list profile_step prints, a per-op table.
   python tools/launch_table.py launches.csv [profile_step.log] [launches_per_step]"""
import csv
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
  lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
  if r.get('Metric Name') != 'gpu__time_duration.sum':
    continue
  v = float(r['Metric Value'].replace(',', ''))
  unit = r['Metric Unit']
  us = v / 1000.0 if unit in ('ns', 'nsecond') else v if unit in ('us', 'usecond') else v * 1000.0
  rows.append((r['Kernel Name'].split('(')[0], r.get('Grid Size', ''), us))
ops = []
if len(sys.argv) > 2:
  for l in open(sys.argv[2]):
    p = l.split()
    if len(p) >= 3 and p[0].isdigit():
      ops.append((p[1], ' '.join(p[2:])))
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else None
if per_step is None:
  # the step ends with decode_kernel
  idx = [i for i, r in enumerate(rows) if 'decode_kernel' in r[0]]
  per_step = idx[-1] - idx[-2]
  last = rows[idx[-2] + 1: idx[-1] + 1]
else:
  last = rows[-per_step:]
tot = sum(r[2] for r in last)
agg = defaultdict(lambda: [0, 0.0])
for n, g, us in last:
  agg[n][0] += 1
  agg[n][1] += us
print('launches in the last step: %d, total %.1f us' % (len(last), tot))
print('| kernel | launches | total us | share |\n|---|---|---|---|')
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
  print('| `%s` | %d | %.1f | %.1f %% |' % (n, c, us, 100 * us / tot))
print()
for i, (n, g, us) in enumerate(last):
  print('%3d %-48s grid %-18s %8.1f us' % (i, n[-48:], g, us))
if ops:
  print('\nops (host plan order):')
  for i, o in enumerate(ops):
    print(i, o)
