"""Aggregate warp-stall samples of the first kernel in an ncu report (SASS source page)."""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr_idx = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
names = [rows[i - 1][1] for i in hdr_idx]
h = rows[hdr_idx[which]]
end = hdr_idx[which + 1] - 1 if which + 1 < len(hdr_idx) else len(rows)
col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot, total, per = {s: 0 for s in stalls}, 0, []
for r in rows[hdr_idx[which] + 1:end]:
    if len(r) < len(h):
        continue
    try:
        ns = int(r[col['# Samples']])
    except ValueError:
        continue
    total += ns
    st = {}
    for s in stalls:
        v = int(r[col[s]] or 0)
        tot[s] += v
        if v:
            st[s] = v
    per.append((ns, r[col['Source']][:80], st))
print(names[which])
print('total samples', total)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
    print(f'  {s:26s} {v:6d} {v / max(total, 1):6.1%}')
print('top instructions:')
for ns, src, st in sorted(per, key=lambda t: -t[0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 16]:
    print(f'  {ns:5d} {src:80s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:2])}')
