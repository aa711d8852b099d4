"""ncu CSVs of tools/capture_traffic.sh -> profiles/traffic.json."""
import csv
import datetime
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

src, dst = sys.argv[1], sys.argv[2]
out = {"how": "tools/capture_traffic.sh: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum "
              "--clock-control none on the second launch of each config's solve kernel "
              "(tools/prof_config.py, full BASELINE batch, device-resident inputs)",
       "csrc_sha256": bench.csrc_hash(), "captured": datetime.date.today().isoformat(), "configs": {}}
for c in bench.CONFIGS:
    path = os.path.join(src, c + ".csv")
    if not os.path.exists(path):
        continue
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    if len(rows) < 2:
        continue
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[1:]:
        key = (r[ix["ID"]], r[ix["Kernel Name"]])
        val = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
                 "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "nsecond": 1e-9, "usecond": 1e-6,
                 "msecond": 1e-3, "second": 1}.get(unit, 1)
        per.setdefault(key, {})[r[ix["Metric Name"]]] = val * scale
    # the launch that did the work: the longest one (the lane path launches two kernels, one
    # of which returns at once); the second solve is the last such launch
    best = None
    for key, m in per.items():
        if best is None or m.get("gpu__time_duration.sum", 0) >= 0.5 * per[best].get("gpu__time_duration.sum", 0):
            if best is None or int(key[0]) > int(best[0]) or m.get("gpu__time_duration.sum", 0) > 2 * per[best].get("gpu__time_duration.sum", 0):
                best = key
    m = per[best]
    wl = bench.Workload(c)
    tot = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
    out["configs"][c] = {"kernel": best[1][:80], "instances": wl.batch,
                         "dram_bytes_per_instance": tot / wl.batch,
                         "algorithmic_bytes_per_instance": wl.bytes_per_solve(),
                         "ratio": tot / wl.batch / wl.bytes_per_solve(),
                         "kernel_ms_under_ncu": 1e3 * m.get("gpu__time_duration.sum", 0)}
with open(dst, "w") as fh:
    json.dump(out, fh, indent=1)
    fh.write("\n")
