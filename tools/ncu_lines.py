"""Join an ncu SASS source page (per-instruction execution counts, program order) with
nvdisasm line info of the same cubin -> executed warp instructions per CUDA source line.

    python tools/ncu_lines.py report.ncu-rep <mangled kernel name> [top N] [column]

`column` defaults to "Instructions Executed"; "# Samples" gives the warp-sampling (time)
attribution per source line instead.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, mangled = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
column = sys.argv[4] if len(sys.argv) > 4 else 'Instructions Executed'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'hrfuser_b200', 'libhrfuser_b200.so')
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
sass = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for ln in sass.splitlines():
    if ln.startswith('.text.'):
        inside = ln.startswith('.text.' + mangled + ':')
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln):
        lines.append(cur)
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h)}
counts = []
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] == 'Address':
        break
    try:
        counts.append(int(r[col[column]]))
    except ValueError:
        break
print(len(lines), 'SASS instructions with line info;', len(counts), 'in the profile')
n = min(len(lines), len(counts))
agg = collections.Counter()
for i in range(n):
    agg[lines[i]] += counts[i]
tot = sum(agg.values())
src_cache = {}
for (f, l), c in agg.most_common(top):
    path = os.path.join(ROOT, 'hrfuser_b200', 'csrc', f) if f else None
    text = ''
    if path and os.path.isfile(path):
        src_cache.setdefault(path, open(path).read().splitlines())
        text = src_cache[path][l - 1].strip()[:90] if l - 1 < len(src_cache[path]) else ''
    print(f'{c:9d} {c / tot:6.1%}  {f}:{l}  {text}')
