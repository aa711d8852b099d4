"""One BASELINE config, device-resident, two batched solves (for ncu -k regex:<kernel>).
Usage: python tools/prof_config.py <2|3a|3b|4a|4a40|4b|5> [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import fbstab_b200 as fb

wl = bench.Workload(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else wl.batch
dev = torch.device("cuda:0")
d = wl.generate(fb.problems, B, 0, 8)
dd = {k: torch.from_numpy(a).to(dev) for k, a in d.items()}
s = wl.solver(fb, B, 0)
for _ in range(2):
    z, l, v = (torch.zeros(B * n, dtype=torch.float64, device=dev) for n in (wl.nz, wl.nl, wl.nv))
    s.solve_batch(dd, z, l, v)
    torch.cuda.synchronize()
print(wl.name, B, s.path[:60])
