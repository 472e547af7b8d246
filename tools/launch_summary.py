"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections, csv, re, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')
    v = float(row['Metric Value'].replace(',', ''))
    v = v / 1e3 if row['Metric Unit'] == 'ns' else (v * 1e3 if row['Metric Unit'] == 'ms' else v)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print("launches %d   total %.1f ms" % (sum(cnt.values()), T / 1e3))
for k, v in tot.most_common(top):
    print("%-64s %6d %9.0f us %5.1f%%" % (k[:64], cnt[k], v, 100 * v / T))
