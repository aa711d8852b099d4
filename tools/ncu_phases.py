"""Per-phase cycles, warp instructions and shared-memory wavefronts of the dense-small
kernel from an ncu capture with source (--import-source on).
Usage: ncu -i X.ncu-rep --page source --csv > sass.csv
       ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python tools/ncu_phases.py sass.csv src.csv <kernel ms> <newton steps in the capture> [dense_small.cu of that build]
The source file must be the one the capture was built from (line numbers)."""
import collections
import csv
import os
import sys

sass_csv, src_csv, ms, newton = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
rows = list(csv.reader(open(src_csv)))
addr2line = {}
cur = line = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = os.path.basename(r[1])
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].strip():
        line = int(r[0])
    if len(r) > 2 and r[2].startswith("0x"):
        addr2line[r[2]] = (cur, line)
s = list(csv.reader(open(sass_csv)))
hdr = s[1]
ix = {h: i for i, h in enumerate(hdr)}
data = s[2:]
here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_path = sys.argv[5] if len(sys.argv) > 5 else os.path.join(here, "fbstab_b200", "csrc", "dense_small.cu")
src = open(src_path).read().split("\n")


def find(st):
    for i, l in enumerate(src):
        if st in l:
            return i + 1
    return 10**9


marks = sorted([
    (find("void publish("), "publish"), (find("double ATv_Hz("), "ATv_Hz"), (find("void Az("), "Az"),
    (find("double GTl("), "GTl"), (find("double Gz("), "Gz"), (find("EvalOut evaluate("), "evaluate"),
    (find("void eliminate("), "eliminate"), (find("void store_row("), "store_row"),
    (find("bool newton_step("), "newton:barrier"),
    (find("// E (lower 8x8 blocks) on the FP64"), "newton:dmma"),
    (find("// fragments -> full symmetric rows"), "newton:transpose"),
    (find("// G block column-wise"), "newton:gload+elimcall"),
    (find("// Schur complement S ="), "newton:schur"),
    (find("// dz = t - Y dl"), "newton:recover"), (find("int feasibility("), "feasibility"),
    (find("// ---- data staging"), "staging"), (find("void v_axpy("), "axpy/select"),
    (find("void solve_one("), "solve_one"), (find("void run_component("), "component"),
    (find("dense_small_kernel(const"), "main")])


def region(l):
    r = "helpers"
    for ln, name in marks:
        if l >= ln:
            r = name
    return r


tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
cps = 8 * 148 * ms * 1e-3 * 1.965e9 / tot
ctx = "?"
samp, inst, wf = collections.Counter(), collections.Counter(), collections.Counter()
wi = ix.get("L1 Wavefronts Shared")
for r in data:
    fl = addr2line.get(r[ix["Address"]])
    if fl and fl[0] == "dense_small.cu" and fl[1] >= marks[0][0]:
        ctx = region(fl[1])
    samp[ctx] += int(r[ix["# Samples"]] or 0)
    inst[ctx] += int(r[ix["Instructions Executed"]] or 0)
    try:
        wf[ctx] += int(r[wi] or 0)
    except (ValueError, TypeError):
        pass
print(f"warp-cycles per sample {cps:.0f} (8 warps x 148 SMs at 1965 MHz)")
print(f"{'phase':24s} {'cycles/newton':>14s} {'instr/newton':>13s} {'smem wavefronts/newton':>23s}")
for k, v in samp.most_common():
    print(f"{k:24s} {v * cps / newton:14.0f} {inst[k] / newton:13.0f} {wf[k] / newton:23.0f}")
print(f"{'total':24s} {sum(samp.values()) * cps / newton:14.0f} {sum(inst.values()) / newton:13.0f} "
      f"{sum(wf.values()) / newton:23.0f}")
