"""Summarise an `ncu --set full` report (one CUDA-graph replay of the bench step) into a
markdown table: per kernel launch duration, DRAM traffic, lane efficiency, issue / tensor
activity.   python tools/ncu_summary.py gpurun_out/top_full.ncu-rep > profiles/rNN_ncu_top_kernels.md"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
col = {n: i for i, n in enumerate(h)}


units = rows[1]
SCALE = {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'second': 1e6,   # -> microseconds
         'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6,
         'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}            # -> megabytes


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(',', '')) * SCALE.get(units[col[name]], 1.0)
    except (KeyError, ValueError):
        return default


agg = collections.OrderedDict()
for r in rows[2:]:
    name = r[col['Kernel Name']].split('(')[0].replace('void ', '').replace('mpa::', '')
    grid = r[col['launch__grid_size']]
    key = (name, grid)
    a = agg.setdefault(key, collections.defaultdict(float))
    a['n'] += 1
    a['us'] += f(r, 'gpu__time_duration.sum')
    a['rd'] += f(r, 'dram__bytes_read.sum')
    a['wr'] += f(r, 'dram__bytes_write.sum')
    a['thr'] += f(r, 'smsp__thread_inst_executed_per_inst_executed.ratio')
    a['issue'] += f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')
    a['tensor'] += f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')
    a['occ'] += f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active')
    a['regs'] = f(r, 'launch__registers_per_thread')
print('| kernel | grid | launches | us / launch | DRAM read MB | DRAM write MB | threads / warp-inst | issue active % | '
      'tensor pipe % | warps active % | regs |')
print('|---|---|---|---|---|---|---|---|---|---|---|')
for (name, grid), a in agg.items():
    n = a['n']
    print(f"| `{name}` | {grid} | {int(n)} | {a['us'] / n:.1f} | {a['rd'] / n:.2f} | {a['wr'] / n:.2f} | "
          f"{a['thr'] / n:.1f} | {a['issue'] / n:.0f} | {a['tensor'] / n:.1f} | {a['occ'] / n:.0f} | {int(a['regs'])} |")
