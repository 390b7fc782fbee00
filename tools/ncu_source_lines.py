"""Per-CUDA-source-line counters of an `ncu --set full --import-source on` report:
    python tools/ncu_source_lines.py REPORT.ncu-rep KERNEL_REGEX [launch_index] [top_n]
Columns: file:line, share of stall samples, share of warp instructions, active threads per warp
instruction, top stall reasons, source text."""
import csv
import io
import os
import re
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                      '--kernel-name', f'regex:{kre}'], capture_output=True, text=True).stdout
# one group of "File Path" blocks per launch; a launch starts again at the kernel's own file
blocks = re.split(r'(?m)^(?="File Path",)', out)
blocks = [b for b in blocks if b.startswith('"File Path"')]
launches, seen = [], set()
for b in blocks:
    path = next(csv.reader(io.StringIO(b.splitlines()[0])))[1]
    if path in seen:
        launches.append([])
        seen = set()
    if not launches:
        launches.append([])
    seen.add(path)
    launches[-1].append((path, b))
lines = []
for path, b in launches[which]:
    rows = list(csv.reader(io.StringIO(b)))
    h = rows[2]
    col = {}
    for i, n in enumerate(h):
        col.setdefault(n, i)
    stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
    for r in rows[3:]:
        if len(r) < len(h) or not r[0].isdigit():
            continue
        def f(n):
            try:
                return float(r[col[n]])
            except ValueError:
                return 0.0
        lines.append(dict(loc=f'{os.path.basename(path)}:{r[0]}', src=r[1].strip(), smp=f('# Samples'),
                          ins=f('Instructions Executed'), thr=f('Thread Instructions Executed'),
                          stalls=sorted(((s, f(s)) for s in stalls), key=lambda t: -t[1])[:2]))
tot_s = sum(l['smp'] for l in lines) or 1
tot_i = sum(l['ins'] for l in lines) or 1
print(f"# launch {which}: samples {int(tot_s)}, warp instr {int(tot_i)}, thread instr {int(sum(l['thr'] for l in lines))}, "
      f"threads / warp instr {sum(l['thr'] for l in lines) / tot_i:.1f}")
for l in sorted(lines, key=lambda l: -l['smp'])[:top_n]:
    print(f"{l['loc']:>20s} {100 * l['smp'] / tot_s:5.1f}% smp {100 * l['ins'] / tot_i:5.1f}% ins thr/ins "
          f"{l['thr'] / max(l['ins'], 1):5.1f} {[(s[6:], int(v)) for s, v in l['stalls']]} | {l['src'][:100]}")
