"""Per-kernel device time of scripts/bench_configs.py under `ncu --metrics gpu__time_duration.sum --csv` (serialised launches:
compare shares).  usage: python scripts/summarize_cfg_launches.py launches.csv"""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    seq.append((re.sub(r'\(.*', '', row['Kernel Name']), float(row['Metric Value'].replace(',', '')) / 1e6, row['Grid Size']))
groups, cur = [], collections.OrderedDict()
for name, val, g in seq:
    if 'configure' in name or 'boot' in name:
        if cur:
            groups.append(cur); cur = collections.OrderedDict()
        continue
    k = (name, g); cur.setdefault(k, [0, 0.0]); cur[k][0] += 1; cur[k][1] += val
if cur:
    groups.append(cur)
for gi, g in enumerate(groups):
    tot = sum(v[1] for v in g.values())
    print('--- plan', gi, 'total ms %.2f' % tot)
    for k, v in g.items():
        print('   %-40s grid=%-12s n=%3d  total=%8.3f ms  avg=%7.3f ms  %4.1f%%' % (k[0][:40], k[1], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
