"""Summarise an `ncu --page source --csv` dump: stall samples by opcode and the
hottest SASS instructions.  Usage: python tools/ncu_source_summary.py file.csv [topN]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, '# Samples') for r in data)
print('kernel', rows[0][1], '| total samples', tot, '| static SASS instructions', len(data))
byop = collections.Counter()
cnt = collections.Counter()
ex = collections.Counter()
for r in data:
    src = r[ix['Source']].strip()
    parts = src.split()
    op = parts[1] if src.startswith('@') else parts[0]
    op = op.split('.')[0]
    byop[op] += I(r, '# Samples')
    cnt[op] += 1
    ex[op] += I(r, 'Instructions Executed')
print('\nby opcode:')
for op, s in byop.most_common(top):
    print(f'  {op:10s} samples {s:8d} {100 * s / max(tot, 1):5.1f}%  static {cnt[op]:5d} executed {ex[op]:12d}')
print('\nstall reasons (all samples):')
for k in hdr:
    if k.startswith('stall_') and 'Not Issued' not in k:
        v = sum(I(r, k) for r in data)
        if v:
            print(f'  {k:24s} {v:8d} {100 * v / max(tot, 1):5.1f}%')
print('\nhottest instructions:')
order = sorted(range(len(data)), key=lambda i: -I(data[i], '# Samples'))[:top]
for i in order:
    r = data[i]
    stalls = {k: I(r, k) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k and I(r, k)}
    main = max(stalls, key=stalls.get) if stalls else ''
    print(f'  #{i:5d} {I(r, "# Samples"):7d}  exec {I(r, "Instructions Executed"):10d}  {main:16s} {r[ix["Source"]].strip()[:90]}')
