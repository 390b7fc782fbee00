"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
total = 0.0
n = 0
for row in r:
    name = re.sub(r'\(.*', '', row[ki])
    name = re.sub(r'^void ', '', name)
    ns = float(row[vi].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns; total += ns; n += 1
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f'{n} launches, {total/1e3/div:.1f} us per step (sum of kernel durations, /{div:g} steps)')
print(f'{"kernel":70s} {"launches/step":>13s} {"us/step":>10s} {"share":>7s}')
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:70]:70s} {c/div:13.1f} {ns/1e3/div:10.1f} {100*ns/total:6.1f}%')
