"""Summarise an ncu `gpu__time_duration.sum` launch list (csv) by kernel family."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4]
    m = re.search(r'(hrf::\w+|cudnn\w*|cutlass\w*|sm\d+_xmma\w*|nhwc\w*|\w+_kernel\w*|\w+)', name)
    short = re.sub(r'<.*', '', name)[:70]
    fam = short
    agg[fam][0] += 1
    agg[fam][1] += float(r[14].replace(',', '')) / 1e3
tot = sum(v[1] for v in agg.values())
print(f'{len(rows)} launches, {tot / 1e3:.3f} ms total (serialised, cold-cache: compare shares)')
print(f'{"kernel":72s} {"n":>5s} {"us":>10s} {"share":>7s}')
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k:72s} {n:5d} {us:10.1f} {us / tot:7.2%}')
