"""Per-source-line stall samples from `ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass`.  Usage: python tools/ncu_lines.py file.csv [topN]"""
import collections
import csv
import os
import sys

top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(sys.argv[1])))
cur_file, hdr, line_no, line_src = None, None, None, None
samples = collections.Counter()
insts = collections.Counter()
src = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
        ei = hdr.index("Instructions Executed")
        continue
    if hdr is None:
        continue
    if r[0].strip():
        line_no = int(r[0])
        src[(cur_file, line_no)] = r[1].strip()
    if len(r) > si and r[2].strip():  # a SASS row
        try:
            samples[(cur_file, line_no)] += int(r[si] or 0)
            insts[(cur_file, line_no)] += int(r[ei] or 0)
        except ValueError:
            pass
tot = sum(samples.values())
print("total samples", tot, "total warp instructions", sum(insts.values()))
byfile = collections.Counter()
for (f, l), s in samples.items():
    byfile[f] += s
print("by file:", {f: f"{100 * s / tot:.1f}%" for f, s in byfile.most_common()})
for (f, l), s in samples.most_common(top):
    print(f"{100 * s / tot:5.1f}%  inst {insts[(f, l)]:11d}  {f}:{l}  {src.get((f, l), '')[:100]}")
