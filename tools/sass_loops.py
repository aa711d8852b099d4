"""Static view of a kernel's SASS (cuobjdump -sass of one function): every
backward branch closes a loop; prints each loop's body size and opcode mix.
Usage: cuobjdump -sass lib.so | python tools/sass_loops.py <function-substring>"""
import collections
import re
import sys

want = sys.argv[1] if len(sys.argv) > 1 else ""
ins = []
on = False
for line in sys.stdin:
    if "Function :" in line:
        on = want in line
        if on:
            ins = []
            name = line.strip()
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(name, "instructions:", len(ins))
addr_ix = {a: i for i, (a, _) in enumerate(ins)}
ops = collections.Counter()
for a, t in ins:
    p = t.split()
    op = p[1] if p[0].startswith("@") else p[0]
    ops[op.split(".")[0]] += 1
print("opcode mix:", dict(ops.most_common(14)))
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt <= a and tgt in addr_ix:
        j = addr_ix[tgt]
        body = ins[j:i + 1]
        c = collections.Counter()
        for _, tt in body:
            p = tt.split()
            op = p[1] if p[0].startswith("@") else p[0]
            c[op.split(".")[0]] += 1
        print(f"loop sass#{j}-{i} ({len(body)} instr):", dict(c.most_common(10)))
