"""SASS mnemonic counts of the built library (cuobjdump -sass): the opcodes that prove tcgen05 / TMEM /
TMA use, per kernel.   python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, 'hrfuser_b200', 'libhrfuser_b200.so')
KEYS = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'HFMA2',
        'MUFU.TANH', 'MUFU.EX2', 'FFMA2', 'FADD2', 'FMUL2', 'LDGSTS']
sass = subprocess.run(['cuobjdump', '-sass', SO], capture_output=True, text=True, check=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                       text=True).stdout.split('\n')
per, total, cur, n_instr, fi = {}, collections.Counter(), None, 0, 0
for line in sass.split('\n'):
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = re.sub(r'\(.*', '', names[fi])
        fi += 1
        per[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        n_instr += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + '.'):
                per[cur][k] += 1
                total[k] += 1
print('SASS mnemonic counts of hrfuser_b200/libhrfuser_b200.so (cuobjdump -sass, sm_100a), round 2 final (tools/sass_counts.py)')
print('whole library: ' + ' '.join(f'{k}={total[k]}' for k in KEYS) + f'  instructions={n_instr}')
print()
print('kernels that use the tensor cores / TMA (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = '
      'cp.async.bulk.tensor load / store, UBLKCP = cp.async.bulk):')
for name in sorted(per):
    c = per[name]
    if any(c[k] for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UTMASTG', 'UBLKCP')):
        print(f'{name:92s} ' + ' '.join(f'{k}={c[k]}' for k in KEYS if c[k]))
