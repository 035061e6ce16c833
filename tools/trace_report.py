"""Concurrency statistics of a kernel timeline written by tools/trace_step.py.

    python tools/trace_report.py gpurun_out/trace.json

Takes the last of the three profiled replays, prints the step span, the time
with 0 / 1 / 2 / 3+ kernels in flight, per-kernel-family busy time and the families that run
ALONE longest (the serial bottlenecks of the graph).
"""
import collections
import json
import re
import sys

recs = json.load(open(sys.argv[1]))
recs = [r for r in recs if r['dur'] > 0 and 'Memcpy' not in r['name'] and 'Memset' not in r['name']]
recs.sort(key=lambda r: r['start'])
# tools/trace_step.py profiles three identical replays: the last third of the kernel records
N_REPLAYS = 3
last = recs[len(recs) - len(recs) // N_REPLAYS:]
t0 = last[0]['start']
t1 = max(r['start'] + r['dur'] for r in last)
print(f'{len(last)} kernels, span {t1 - t0:.1f} us, sum of kernel durations {sum(r["dur"] for r in last):.1f} us')


def fam(n):
    n = re.sub(r'^void ', '', n)
    m = re.match(r'(hrf::\w+)(<[^>]*>)?', n)
    if m:
        return m.group(1) + (m.group(2) or '')
    return n.split('<')[0].split('(')[0][:60]


ev = []
for i, r in enumerate(last):
    ev.append((r['start'], 1, i))
    ev.append((r['start'] + r['dur'], -1, i))
ev.sort()
active, prev = set(), t0
level = collections.Counter()
alone = collections.Counter()
for t, d, i in ev:
    dt = t - prev
    if dt > 0:
        level[min(len(active), 4)] += dt
        if len(active) == 1:
            alone[fam(last[next(iter(active))]['name'])] += dt
    prev = t
    if d > 0:
        active.add(i)
    else:
        active.discard(i)
print('kernels in flight -> us:', {k: round(v, 1) for k, v in sorted(level.items())})
busy = collections.Counter()
cnt = collections.Counter()
for r in last:
    busy[fam(r['name'])] += r['dur']
    cnt[fam(r['name'])] += 1
print('\nfamily: n, total us (in-graph durations), alone us')
for k, v in busy.most_common(25):
    print(f'  {k[:70]:70s} {cnt[k]:4d} {v:9.1f} {alone.get(k, 0):8.1f}')
streams = collections.Counter(r['stream'] for r in last)
print('\nstreams:', dict(streams))
# coarse timeline: 20 bins, dominant family per bin
nb = 24
print('\ntimeline (dominant family by busy time per bin):')
for b in range(nb):
    lo, hi = t0 + (t1 - t0) * b / nb, t0 + (t1 - t0) * (b + 1) / nb
    c = collections.Counter()
    for r in last:
        s, e = max(r['start'], lo), min(r['start'] + r['dur'], hi)
        if e > s:
            c[fam(r['name'])] += e - s
    top = ', '.join(f'{k[:38]} {v / (hi - lo):.2f}' for k, v in c.most_common(3))
    print(f'  {lo - t0:7.0f} us  load {sum(c.values()) / (hi - lo):4.2f}  {top}')
