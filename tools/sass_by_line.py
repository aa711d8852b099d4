"""Static SASS instructions per source line / region from `nvdisasm -g -c x.cubin`.
Usage: python tools/sass_by_line.py dis.txt <function-substring> [first-instr last-instr]"""
import collections
import re
import sys

txt = open(sys.argv[1]).read().split("\n")
want = sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
on = False
cur = ("?", 0)
n = 0
per = collections.Counter()
first = {}
for line in txt:
    if line.startswith("//-") and ".text." in line:
        on = want in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        if lo <= n <= hi:
            per[cur] += 1
            first.setdefault(cur, n)
        n += 1
print("instructions", n)
tot = sum(per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{v:5d} {100*v/tot:5.1f}%  {k[0]}:{k[1]}  (first at #{first[k]})")
